#!/bin/bash
# first GPU session: probes + kernel tests + model tests, each under its own timeout, logs in gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 300 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not conv_tc and not large and not (match_argmax and 1) and not split_is" -x --tb=short > gpurun_out/k_simt.log 2>&1; echo "rc=$?" >> gpurun_out/k_simt.log
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv_tc or large" --tb=line > gpurun_out/k_tc.log 2>&1; echo "rc=$?" >> gpurun_out/k_tc.log
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "match_argmax or split_is" --tb=short > gpurun_out/k_match.log 2>&1; echo "rc=$?" >> gpurun_out/k_match.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s --tb=short > gpurun_out/model.log 2>&1; echo "rc=$?" >> gpurun_out/model.log
tail -3 gpurun_out/probe.log gpurun_out/k_simt.log gpurun_out/k_tc.log gpurun_out/k_match.log gpurun_out/model.log
