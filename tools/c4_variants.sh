#!/bin/bash
# tools/c4_variants.sh LIB ... -- C4 (32768^2 x 1e6) one-GPU bench line for several builds of the library
for lib in "$@"; do
    if [ "$lib" = "-" ]; then unset TESS_B200_LIB; else export TESS_B200_LIB=$PWD/$lib; fi
    python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'ms/step %.4f kernel_ms %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
