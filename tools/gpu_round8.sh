#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 36 --warmup 18 > gpurun_out/bench_mfid.json 2> gpurun_out/bench_mfid.err; echo "rc=$?" >> gpurun_out/bench_mfid.err
timeout -s KILL 600 python bench.py --steps 36 --warmup 18 --workload small_mfid --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?" >> gpurun_out/bench_small.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 12 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -n 4 gpurun_out/bench_mfid.err gpurun_out/bench_small.err; cut -c1-1200 gpurun_out/bench_mfid.json; echo; cut -c1-400 gpurun_out/bench_small.json
