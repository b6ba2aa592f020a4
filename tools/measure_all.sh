# tools/measure_all.sh -- the single-GPU evidence set behind profiles/ (round 2).  Run under gpurun from
# the repository root; everything lands in gpurun_out/ with an r02_ prefix.  Each program runs plain
# first; the ncu passes repeat the same command lines afterwards (one gpurun call: all ncu runs count
# as one).
set -x
P=gpurun_out/r02
python -m pytest tests -m gpu -x -q > ${P}_pytest_gpu.log 2>&1
python __graft_entry__.py smoke > ${P}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > ${P}_bench_c3_cvt.json 2> ${P}_bench_c3_cvt.err
python bench.py --steps 20 --warmup 5 --mode sn --no-cpu-baseline --no-extras > ${P}_bench_c3_sn.json 2> ${P}_bench_c3_sn.err
python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu-baseline --no-extras > ${P}_bench_c4_1gpu.json 2> ${P}_bench_c4_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err
python tools/regime_bench.py 8192 > ${P}_regimes_8192.jsonl 2> ${P}_regimes.err
TESS_NO_TMA=1 python tools/regime_bench.py 8192 12000,150000 > ${P}_regimes_8192_no_tma_prefetch.jsonl 2>> ${P}_regimes.err
python tools/other_configs.py > ${P}_other_configs.jsonl 2> ${P}_other.err
timeout 200 python tools/dense_probe.py > ${P}_small_bins_1024.jsonl 2>&1
timeout 200 python tools/dense_centre_probe.py > ${P}_dense_centre_4096.jsonl 2>&1
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 80 --csv --log-file ${P}_launches_c3.csv $B > ${P}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_image_main -s 4 -c 1 -f -o ${P}_full_main_c3_cvt $B >> ${P}_ncu.log 2>&1
ncu --set full --clock-control none -k regex:k_image_main -s 4 -c 1 -f -o ${P}_full_main_c3_sn $B --mode sn >> ${P}_ncu.log 2>&1
ncu --set full --clock-control none -k regex:k_image_main -s 3 -c 1 -f -o ${P}_full_main_c4 $B --workload c4 >> ${P}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_build_qlists|k_iteration_phases" -s 8 -c 2 -f -o ${P}_full_builder_phases_c3 $B >> ${P}_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"k_points_main|k_iteration_phases" -s 12 -c 2 -f -o ${P}_full_points_c2 python tools/other_configs.py >> ${P}_ncu.log 2>&1
