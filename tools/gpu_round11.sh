#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python tools/rb_probe.py > gpurun_out/rb_probe.log 2>&1; echo "rc=$?" >> gpurun_out/rb_probe.log
cat gpurun_out/rb_probe.log | head -70
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=line -k "resblock" > gpurun_out/k_rb.log 2>&1; echo "rc=$?" >> gpurun_out/k_rb.log
tail -n 12 gpurun_out/k_rb.log
