#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/r7_tests.log 2>&1
tail -5 gpurun_out/r7_tests.log
timeout -s KILL 100 python tools/conv_sweep.py > gpurun_out/r7_sweep.log 2>&1
REFVSR_NO_KXFOLD=1 timeout -s KILL 100 python tools/conv_sweep.py > gpurun_out/r7_sweep_nofold.log 2>&1
timeout -s KILL 100 python tools/conv_trace.py 270 480 > gpurun_out/r7_trace_lr.log 2>&1

timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
REFVSR_NO_KXFOLD=1 timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_nofold.json 2>> gpurun_out/r7_bench.err
python - <<'PY'
import json
for f in ('r7_bench','r7_bench_nofold'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)

PY
paste gpurun_out/r7_sweep.log gpurun_out/r7_sweep_nofold.log | awk -F'\t' '{print substr($1,1,75), "| nofold:", substr($2,47,10)}'
