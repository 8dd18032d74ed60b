#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/r8_tests.log 2>&1
tail -3 gpurun_out/r8_tests.log
timeout -s KILL 300 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err
REFVSR_KXFOLD_ALL=1 timeout -s KILL 300 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_foldall.json 2>> gpurun_out/r8_bench.err
python - <<'PY'
import json
for f in ('r8_bench','r8_bench_foldall'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r8_bench.err
