#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x > gpurun_out/k_all.log 2>&1; echo "rc=$?" >> gpurun_out/k_all.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s --tb=short > gpurun_out/model.log 2>&1; echo "rc=$?" >> gpurun_out/model.log
timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err; echo "rc=$?" >> gpurun_out/bench_fuse.err
timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline --no-fuse > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 2 gpurun_out/k_all.log gpurun_out/model.log gpurun_out/smoke.log; tail -n 3 gpurun_out/bench_fuse.err; grep -E "psnr" gpurun_out/model.log | awk '{print $0}' | sort -t' ' -k5 -n | head -5
python - <<'PY'
import json
for f in ['gpurun_out/bench_fuse.json','gpurun_out/bench_nofuse.json']:
    d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
    print(f, {k: d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', d['e2e']['value'], 'conv us', d['roofline']['seconds']*1e6, {k: round(v['seconds']*1e6,1) for k,v in d['roofline_other'].items()})
PY
