#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r5_tests.log 2>&1
tail -3 gpurun_out/r5_tests.log
timeout -s KILL 100 python tools/conv_trace.py 270 480 > gpurun_out/r5_trace_lr.log 2>&1
timeout -s KILL 100 python tools/conv_trace.py 540 960 > gpurun_out/r5_trace_2x.log 2>&1
timeout -s KILL 100 python tools/conv_sweep.py > gpurun_out/r5_sweep.log 2>&1
timeout -s KILL 300 python tools/profile_kernels.py > gpurun_out/r5_kern.json 2> gpurun_out/r5_kern.err
timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --fuse > gpurun_out/r5_bench_fuse.json 2>> gpurun_out/r5_bench.err
python - <<'PY'
import json
for f in ('r5_bench','r5_bench_fuse'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -E '"seconds"|"frac"|name' gpurun_out/r5_kern.json | head -30
