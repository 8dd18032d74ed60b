#!/bin/bash
# round 2, GPU call A: chain-kernel validation first, then the whole GPU suite, bench (chain on / off), ncu captures
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "chain" > gpurun_out/a_chain.log 2>&1
CH=$?
tail -5 gpurun_out/a_chain.log
if [ $CH -ne 0 ]; then export REFVSR_NO_CHAIN=1; echo "CHAIN TESTS FAILED -> chain disabled for the rest"; fi
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -q -k "not chain" > gpurun_out/a_kernels.log 2>&1; tail -3 gpurun_out/a_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -s > gpurun_out/a_model.log 2>&1; tail -3 gpurun_out/a_model.log
timeout -s KILL 1200 python -m pytest tests/test_gpu_fullsize.py -q -s > gpurun_out/a_fullsize.log 2>&1; tail -12 gpurun_out/a_fullsize.log
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 1500 gpurun_out/a_bench.json
REFVSR_NO_CHAIN=1 timeout -s KILL 400 python bench.py --no-cpu-baseline --no-eager --no-sustained > gpurun_out/a_bench_nochain.json 2> gpurun_out/a_bench_nochain.err; head -c 400 gpurun_out/a_bench_nochain.json
timeout -s KILL 900 python -m pytest tests/test_dropin_runpy.py -q -s -m gpu > gpurun_out/a_dropin.log 2>&1; tail -5 gpurun_out/a_dropin.log
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_window.csv python tools/profile_window.py > gpurun_out/a_ncu_window.log 2>&1
tail -2 gpurun_out/a_ncu_window.log; wc -l gpurun_out/r02_launches_window.csv
for K in warp_vec gather_blocks aligned_sample reconstruct conv_chain; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_$K python tools/profile_kernels.py > gpurun_out/a_ncu_$K.log 2>&1
  tail -1 gpurun_out/a_ncu_$K.log
done
ls -la gpurun_out/*.ncu-rep
