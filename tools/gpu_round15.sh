#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x -k "conv or warp or resblock" > gpurun_out/k_conv.log 2>&1; echo "rc=$?" >> gpurun_out/k_conv.log
timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?" >> gpurun_out/bench_final.err
tail -n 2 gpurun_out/k_conv.log; tail -n 2 gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_final.json').read().splitlines() if l.startswith('{')][-1])
print({k: d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], 'conv us', d['roofline']['seconds']*1e6, {k: (round(v['seconds']*1e6,1), round(v['frac'],3)) for k,v in d['roofline_other'].items()})
PY
