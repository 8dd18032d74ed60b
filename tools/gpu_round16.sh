#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x -k "conv or resblock" > gpurun_out/k_conv.log 2>&1; echo "rc=$?" >> gpurun_out/k_conv.log
timeout -s KILL 200 python tools/conv_sweep.py > gpurun_out/sweep_grp.log 2>&1; head -4 gpurun_out/sweep_grp.log; sed -n 13,16p gpurun_out/sweep_grp.log
timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_grp.json 2> gpurun_out/bench_grp.err; echo "rc=$?" >> gpurun_out/bench_grp.err
timeout -s KILL 300 python tools/match_mode_check.py > gpurun_out/match_mode.log 2>&1; cat gpurun_out/match_mode.log | tail -8
tail -n 2 gpurun_out/k_conv.log; tail -n 2 gpurun_out/bench_grp.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_grp.json').read().splitlines() if l.startswith('{')][-1])
print({k: d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], 'conv us', d['roofline']['seconds']*1e6)
PY
