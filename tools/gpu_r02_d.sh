#!/bin/bash
# round 2, GPU call D (--gpus 2): frame-sharded clip over NCCL: bit-exactness vs the single stream, then the bench with the clip leg
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export NCCL_MAX_P2P_NCHANNELS=4
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/d_check.log 2>&1; echo "dist_check rc=$?"; grep "dist_check\|Error\|error" gpurun_out/d_check.log | tail -6
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 18 --warmup 3 > gpurun_out/d_bench2.json 2> gpurun_out/d_bench2.err; echo "bench rc=$?"; head -c 300 gpurun_out/d_bench2.json; echo; tail -5 gpurun_out/d_bench2.err
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "warp3" > gpurun_out/d_warp3.log 2>&1; tail -2 gpurun_out/d_warp3.log
timeout -s KILL 300 python tools/profile_kernels.py 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print({k:(round(v['seconds']*1e6,1), round(v['frac'],3)) for k,v in d.items()})"
