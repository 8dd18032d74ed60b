#!/bin/bash
# scheduling iteration: model parity (incl. overlap exactness), then the bench per overlap / cap variant
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -4
run_bench() {
  timeout -s KILL 400 python bench.py --no-eager --no-sustained --no-clip --no-cpu-baseline > gpurun_out/j_bench_$1.json 2> gpurun_out/j_bench_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/j_bench_$1.json").read().strip().splitlines()[-1])
print("$1", {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"],2), "launches", d.get("gpu_launches"))
PY
}
export REFVSR_COLSPLIT=0
REFVSR_NO_OVERLAP=1 run_bench sequential
REFVSR_OVERLAP_CAP=74 REFVSR_OVERLAP_MAIN_CAP=74 run_bench s74_m74
REFVSR_OVERLAP_CAP=74 REFVSR_OVERLAP_MAIN_CAP=0 run_bench s74_m0
REFVSR_OVERLAP_CAP=100 REFVSR_OVERLAP_MAIN_CAP=0 run_bench s100_m0
REFVSR_OVERLAP_CAP=50 REFVSR_OVERLAP_MAIN_CAP=0 run_bench s50_m0
REFVSR_OVERLAP_CAP=0 REFVSR_OVERLAP_MAIN_CAP=0 run_bench s0_m0
REFVSR_OVERLAP_CAP=74 REFVSR_OVERLAP_MAIN_CAP=74 REFVSR_OVERLAP_BW_STEPS=1 run_bench s74_m74_bw1
REFVSR_OVERLAP_CAP=60 REFVSR_OVERLAP_MAIN_CAP=88 REFVSR_OVERLAP_BW_STEPS=1 run_bench s60_m88_bw1
