#!/bin/bash
# pointwise-kernel iteration: parity of the gather / warp / resampling kernels, then the bench line (kernel rooflines in roofline_other)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "warp or gather or aligned or reconstruct" 2>&1 | tail -5
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -3
timeout -s KILL 500 python bench.py --no-eager --no-sustained --no-clip > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/g_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"])
for k,v in d["roofline_other"].items(): print(f"{k:20s} frac={v['frac']:.3f} {v['seconds']*1e6:8.1f} us")
PY
tail -3 gpurun_out/g_bench.err
