#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 18 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout -s KILL 300 python bench.py --gpus 1 --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_bench.log 2>&1
tail -n 6 gpurun_out/bench_n2.err; cut -c1-600 gpurun_out/bench_n2.json; echo; cut -c1-300 gpurun_out/bench_n1.json
