#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv or resblock" > gpurun_out/r6_tests.log 2>&1
tail -3 gpurun_out/r6_tests.log
timeout -s KILL 100 python tools/conv_trace.py 270 480 > gpurun_out/r6_trace_lr.log 2>&1
timeout -s KILL 100 python tools/conv_trace.py 540 960 > gpurun_out/r6_trace_2x.log 2>&1
timeout -s KILL 100 python tools/conv_sweep.py > gpurun_out/r6_sweep.log 2>&1
REFVSR_NO_FAST_EPILOGUE=1 timeout -s KILL 100 python tools/conv_sweep.py > gpurun_out/r6_sweep_nofast.log 2>&1
timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
python - <<'PY'
import json
for f in ('r6_bench',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
paste gpurun_out/r6_sweep.log gpurun_out/r6_sweep_nofast.log | awk -F'\t' '{print substr($1,1,75), "| nofast:", substr($2,47,10)}'
