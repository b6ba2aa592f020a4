#!/bin/bash
# tools/quick_bench.sh TAG [LIB] -- C3 bench line (no CPU baseline, no e2e) + regime table for one build
# of the library (tuning aid; results under gpurun_out/).
TAG=$1; LIB=$2
if [ -n "$LIB" ]; then export TESS_B200_LIB=$PWD/$LIB; fi
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/b_$TAG.json 2> gpurun_out/b_$TAG.err
python tools/regime_bench.py 8192 80,320,1500,12000,150000 > gpurun_out/reg_$TAG.jsonl 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/b_$TAG.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$TAG', 'ms/step %.4f kernel_ms %.4f frac %.3f step_frac %.3f'%(d['ms_per_step'],r['kernel_ms'],r['frac'],r['step_frac_of_roofline']))
except Exception as e:
    print('$TAG ERR', e)
for l in open('gpurun_out/reg_$TAG.jsonl'):
    try:
        x=json.loads(l); print('   area %6d  main %.4f ms  %.5f ns/px  iter %.4f'%(x['area'],x['ms_main'],x['ns_per_px_main'],x['ms_iter']))
    except Exception: print(l.strip()[:200])
PY
