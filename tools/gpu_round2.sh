#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 18 --warmup 3 > gpurun_out/bench_mfid.json 2> gpurun_out/bench_mfid.err; echo "rc=$?" >> gpurun_out/bench_mfid.err
timeout -s KILL 600 python bench.py --steps 18 --warmup 3 --workload small_mfid --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?" >> gpurun_out/bench_small.err
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 2 -o gpurun_out/prof_conv_tc -f python tools/profile_kernels.py > gpurun_out/ncu_conv.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:warp_kernel -s 3 -c 2 -o gpurun_out/prof_warp -f python tools/profile_kernels.py > gpurun_out/ncu_warp.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -s 1 -c 1 -o gpurun_out/prof_match -f python tools/profile_kernels.py > gpurun_out/ncu_match.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
cat gpurun_out/bench_mfid.json; tail -n 3 gpurun_out/bench_mfid.err
