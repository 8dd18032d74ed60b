#!/bin/bash
# round 2, last validation after the stream-overlap changes: the whole -m gpu suite, smoke, bench (ours, small)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/f_gpu_tests.log 2>&1; tail -3 gpurun_out/f_gpu_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; head -c 200 gpurun_out/f_bench.json; echo; tail -2 gpurun_out/f_bench.err
timeout -s KILL 600 python bench.py --workload small_mfid --no-cpu-baseline --no-clip > gpurun_out/f_bench_small.json 2> gpurun_out/f_bench_small.err; head -c 200 gpurun_out/f_bench_small.json; echo
