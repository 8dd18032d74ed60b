#!/bin/bash
# scheduling iteration: model parity (incl. overlap exactness), then the bench with / without the overlap
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -3
run_bench() {
  timeout -s KILL 400 python bench.py --no-eager --no-sustained --no-clip --no-cpu-baseline > gpurun_out/j_bench_$1.json 2> gpurun_out/j_bench_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/j_bench_$1.json").read().strip().splitlines()[-1])
print("$1", {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"],2), "launches", d.get("gpu_launches"))
PY
}
run_bench overlap
REFVSR_NO_OVERLAP=1 run_bench sequential
run_bench overlap_again
