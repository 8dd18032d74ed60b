#!/bin/bash
# full round-end validation: every GPU test, smoke(), the default bench line (with cpu_baseline) and the reference arm
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short > gpurun_out/final_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/final_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s --tb=short > gpurun_out/final_model.log 2>&1; echo "rc=$?" >> gpurun_out/final_model.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?" >> gpurun_out/final_bench.err
tail -n 3 gpurun_out/final_kernels.log gpurun_out/final_model.log gpurun_out/final_smoke.log gpurun_out/final_bench.err
cut -c1-1800 gpurun_out/final_bench.json
