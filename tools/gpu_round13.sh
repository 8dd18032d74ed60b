#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/tc_probe.py > gpurun_out/probe2.log 2>&1
grep -E "TC_PROBE|max err|first bad|got|exp|EXC" gpurun_out/probe2.log | head -60
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=line -k "conv_tc" > gpurun_out/k_tc.log 2>&1; tail -n 8 gpurun_out/k_tc.log
REFVSR_TC_LAYOUT=2 timeout -s KILL 200 python tools/conv_sweep.py > gpurun_out/sweep2.log 2>&1; cat gpurun_out/sweep2.log
