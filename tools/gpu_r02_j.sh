#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run_bench() {
  timeout -s KILL 400 python bench.py --no-eager --no-sustained --no-clip --no-cpu-baseline > gpurun_out/j_bench_$1.json 2> gpurun_out/j_bench_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/j_bench_$1.json").read().strip().splitlines()[-1])
print("$1", {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"],2), "launches", d.get("gpu_launches"))
PY
}
REFVSR_NO_HEAD=1 run_bench nohead_prio
REFVSR_NO_PRIO=1 run_bench head_noprio
REFVSR_NO_HEAD=1 REFVSR_NO_PRIO=1 run_bench nohead_noprio
