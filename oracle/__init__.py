"""
oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's Lloyd / Voronoi-assignment hot path (jonathansick/tess,
tess/lloyd.pyx:10-90, tess/voronoi.py:75-114,259-305).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; tess_b200/
never does (tests/test_abi.py enforces that).

Parity pin: oracle.py is checked bitwise against the compiled reference itself
(oracle/_ref, built from /root/reference by oracle/build_ref.py) and against the committed
golden fixtures generated from it (tests/golden/make_golden.py) -- see tests/test_oracle.py.
"""
