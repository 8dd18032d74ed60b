#!/bin/bash
# round 2, last validation: the whole -m gpu suite, smoke, bench (ours, small), refreshed ncu evidence
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/f_gpu_tests.log 2>&1; tail -3 gpurun_out/f_gpu_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; head -c 200 gpurun_out/f_bench.json; echo; tail -2 gpurun_out/f_bench.err
timeout -s KILL 600 python bench.py --workload small_mfid --no-cpu-baseline --no-clip > gpurun_out/f_bench_small.json 2> gpurun_out/f_bench_small.err; head -c 200 gpurun_out/f_bench_small.json; echo
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_conv_tc python tools/ncu_conv.py > gpurun_out/f_ncu_conv.log 2>&1; tail -1 gpurun_out/f_ncu_conv.log
for k in gather_cells_kernel warp_vec_kernel; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_${k%_kernel} python tools/profile_kernels.py > gpurun_out/f_ncu_$k.log 2>&1; tail -1 gpurun_out/f_ncu_$k.log
done
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_window.csv python tools/profile_window.py > gpurun_out/f_ncu_window.log 2>&1; tail -1 gpurun_out/f_ncu_window.log
