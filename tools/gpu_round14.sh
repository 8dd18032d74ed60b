#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x > gpurun_out/k_all.log 2>&1; echo "rc=$?" >> gpurun_out/k_all.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s --tb=short > gpurun_out/model.log 2>&1; echo "rc=$?" >> gpurun_out/model.log
timeout -s KILL 400 python bench.py --gpus 1 --steps 18 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?" >> gpurun_out/bench_final.err
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_window.py > gpurun_out/ncu_window.log 2>&1; echo "ncu rc=$?"
tail -n 2 gpurun_out/k_all.log gpurun_out/model.log; grep -E "fused window" gpurun_out/model.log; tail -n 2 gpurun_out/bench_final.err; wc -l gpurun_out/launches.csv
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_final.json').read().splitlines() if l.startswith('{')][-1])
print({k: d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], 'conv us', d['roofline']['seconds']*1e6, {k: (round(v['seconds']*1e6,1), round(v['frac'],3)) for k,v in d['roofline_other'].items()})
print(d.get('cpu_baseline'))
PY
