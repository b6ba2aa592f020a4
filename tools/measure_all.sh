set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r_c3_cvt.json 2> gpurun_out/r_c3_cvt.err
python bench.py --steps 20 --warmup 5 --mode sn --no-cpu-baseline > gpurun_out/r_c3_sn.json 2> gpurun_out/r_c3_sn.err
python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/r_c4.json 2> gpurun_out/r_c4.err
python tools/regime_bench.py > gpurun_out/r_regimes.jsonl 2> gpurun_out/r_regimes.err
python tools/other_configs.py > gpurun_out/r_other.jsonl 2> gpurun_out/r_other.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r_launches_c3.csv python bench.py --steps 2 --warmup 5 --no-e2e --no-cpu-baseline --no-clocks > gpurun_out/r_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_image_main -s 3 -c 1 -f -o gpurun_out/r_full_main python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks > gpurun_out/r_ncu_full.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r_ref.json 2> gpurun_out/r_ref.err
python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest_gpu.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r_smoke.log 2>&1
timeout 300 python tools/dense_probe.py > gpurun_out/r_dense.jsonl 2>&1
timeout 300 python tools/dense_centre_probe.py > gpurun_out/r_dense_centre.jsonl 2>&1
