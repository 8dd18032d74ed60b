#!/bin/bash
# tile-shape sweep of the gather kernels (tools/pointwise_bench.py), then their parity tests at the variants
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 0 1 2; do
  REFVSR_W3_TILE=$v REFVSR_AS_TILE=$v timeout -s KILL 200 python tools/pointwise_bench.py 2>&1 | grep -v Warning | tee -a gpurun_out/h_tiles.log
done
for v in 1 2; do
  REFVSR_W3_TILE=$v REFVSR_AS_TILE=$v timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "warp3 or aligned" 2>&1 | tail -2
done
