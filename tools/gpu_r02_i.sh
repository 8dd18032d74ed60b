#!/bin/bash
# ncu --set full of the gather kernels at the BASELINE shapes (tools/pointwise_bench.py)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in warp3_kernel aligned_sample2_kernel; do
  timeout -s KILL 250 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_${k%_kernel} python tools/pointwise_bench.py > gpurun_out/i_ncu_$k.log 2>&1; tail -1 gpurun_out/i_ncu_$k.log
done
