#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/r11_tests.log 2>&1
tail -2 gpurun_out/r11_tests.log
timeout -s KILL 200 python tools/conv_hist.py > gpurun_out/r11_hist.log 2>&1
head -14 gpurun_out/r11_hist.log
timeout -s KILL 300 python bench.py --steps 36 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r11_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
PY
