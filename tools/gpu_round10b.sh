#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/r10_tests.log 2>&1
tail -2 gpurun_out/r10_tests.log
timeout -s KILL 100 python tools/conv_sweep.py > gpurun_out/r10_sweep.log 2>&1
head -3 gpurun_out/r10_sweep.log; grep "540x960" gpurun_out/r10_sweep.log
timeout -s KILL 300 python bench.py --steps 36 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r10_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
PY
