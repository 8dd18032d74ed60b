#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x > gpurun_out/k_all.log 2>&1; echo "rc=$?" >> gpurun_out/k_all.log
timeout -s KILL 300 python tools/conv_sweep.py > gpurun_out/sweep.log 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/model.log 2>&1; echo "rc=$?" >> gpurun_out/model.log
timeout -s KILL 600 python bench.py --steps 36 --warmup 21 --no-cpu-baseline > gpurun_out/bench_mfid.json 2> gpurun_out/bench_mfid.err; echo "rc=$?" >> gpurun_out/bench_mfid.err
tail -n 2 gpurun_out/k_all.log gpurun_out/model.log; cat gpurun_out/sweep.log; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_mfid.json'))
print({k: d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('conv', d['roofline']['seconds']*1e6, 'us', d['roofline']['frac'])
for k,v in d['roofline_other'].items(): print(k, v['seconds']*1e6, 'us', v['achieved'], v['unit'], v['frac'])
PY
