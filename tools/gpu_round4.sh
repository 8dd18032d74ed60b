#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x > gpurun_out/k_all.log 2>&1; echo "rc=$?" >> gpurun_out/k_all.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s --tb=short > gpurun_out/model.log 2>&1; echo "rc=$?" >> gpurun_out/model.log
timeout -s KILL 600 python bench.py --steps 36 --warmup 21 --no-cpu-baseline > gpurun_out/bench_mfid.json 2> gpurun_out/bench_mfid.err; echo "rc=$?" >> gpurun_out/bench_mfid.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 1 -o gpurun_out/prof_conv_tc -f python tools/profile_kernels.py > gpurun_out/ncu_conv.log 2>&1
tail -n 3 gpurun_out/k_all.log gpurun_out/model.log; tail -n 5 gpurun_out/bench_mfid.err; cat gpurun_out/bench_mfid.json | cut -c1-900
