#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x -k "conv" > gpurun_out/k_conv.log 2>&1; echo "rc=$?" >> gpurun_out/k_conv.log
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/model.log 2>&1; echo "rc=$?" >> gpurun_out/model.log
timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "rc=$?" >> gpurun_out/bench_pdl.err
REFVSR_NO_PDL=1 timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
tail -n 2 gpurun_out/k_conv.log gpurun_out/model.log; tail -n 3 gpurun_out/bench_pdl.err
python - <<'PY'
import json
for f in ['gpurun_out/bench_pdl.json','gpurun_out/bench_nopdl.json']:
    d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
    print(f, {k: d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], 'conv us', d['roofline']['seconds']*1e6)
PY
