#!/bin/bash
# round 2, GPU call B: chain kernel (both dtypes) -> model tests / bench with the chain on, ncu of the chain kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "chain" > gpurun_out/b_chain.log 2>&1
CH=$?
tail -4 gpurun_out/b_chain.log
if [ $CH -ne 0 ]; then export REFVSR_NO_CHAIN=1; echo "CHAIN TESTS FAILED -> chain disabled for the rest"; fi
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -q -x > gpurun_out/b_model.log 2>&1; tail -3 gpurun_out/b_model.log
timeout -s KILL 900 python -m pytest tests/test_gpu_fullsize.py -q -s -k "bf16 or (small and fp16)" > gpurun_out/b_fullsize.log 2>&1; grep "window\|passed\|failed" gpurun_out/b_fullsize.log | tail -9
timeout -s KILL 600 python bench.py --no-cpu-baseline --no-eager > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; head -c 600 gpurun_out/b_bench.json; echo
timeout -s KILL 600 python bench.py --no-cpu-baseline --no-eager --no-sustained --workload small_mfid > gpurun_out/b_bench_small.json 2> gpurun_out/b_bench_small.err; head -c 400 gpurun_out/b_bench_small.json; echo
timeout -s KILL 900 python -m pytest tests/test_dropin_runpy.py -q -s -m gpu > gpurun_out/b_dropin.log 2>&1; grep "dPSNR\|passed\|failed" gpurun_out/b_dropin.log | tail -5
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_window_chain.csv python tools/profile_window.py > gpurun_out/b_ncu_window.log 2>&1
tail -1 gpurun_out/b_ncu_window.log; wc -l gpurun_out/r02_launches_window_chain.csv
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:conv_chain --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_conv_chain python tools/profile_kernels.py > gpurun_out/b_ncu_chain.log 2>&1
tail -1 gpurun_out/b_ncu_chain.log
ls -la gpurun_out/*.ncu-rep
