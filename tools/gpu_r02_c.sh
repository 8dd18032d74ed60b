#!/bin/bash
# round 2, GPU call C: new pointwise kernels (warp3 / aligned_sample2 / reconstruct4), model tests, bench N=1 with the clip leg
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "warp or aligned or reconstruct or chain" > gpurun_out/c_kernels.log 2>&1; tail -4 gpurun_out/c_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -x > gpurun_out/c_model.log 2>&1; tail -3 gpurun_out/c_model.log
timeout -s KILL 600 python bench.py --no-cpu-baseline --no-eager > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; head -c 300 gpurun_out/c_bench.json; echo; tail -3 gpurun_out/c_bench.err
for K in warp3 aligned_sample2 reconstruct4 match_tc; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_$K python tools/profile_kernels.py > gpurun_out/c_ncu_$K.log 2>&1
  tail -1 gpurun_out/c_ncu_$K.log
done
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_window.csv python tools/profile_window.py > gpurun_out/c_ncu_window.log 2>&1
tail -1 gpurun_out/c_ncu_window.log
