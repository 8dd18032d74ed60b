#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv or resblock" > gpurun_out/dual_tests.log 2>&1
tail -3 gpurun_out/dual_tests.log
timeout -s KILL 200 python tools/conv_sweep.py > gpurun_out/dual_sweep.log 2>&1
REFVSR_NO_DUAL_MMA=1 timeout -s KILL 200 python tools/conv_sweep.py > gpurun_out/dual_sweep_off.log 2>&1
tail -12 gpurun_out/dual_sweep.log; echo ---; tail -12 gpurun_out/dual_sweep_off.log
timeout -s KILL 300 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/dual_bench.json 2> gpurun_out/dual_bench.err
REFVSR_NO_DUAL_MMA=1 timeout -s KILL 300 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/dual_bench_off.json 2>> gpurun_out/dual_bench.err
python - <<'PY'
import json
for f in ('dual_bench','dual_bench_off'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
