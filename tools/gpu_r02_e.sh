#!/bin/bash
# multi-GPU validation of the frame-sharded clip (N = $1): bit-exactness vs the single stream, then the bench with the clip leg
N=${1:-4}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export NCCL_MAX_P2P_NCHANNELS=4
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/e_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep "dist_check\|Error\|error" gpurun_out/e_check_n$N.log | tail -6
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 18 --warmup 3 > gpurun_out/e_bench_n$N.json 2> gpurun_out/e_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/e_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "halo_ms", d["config"]["halo_exchange_ms"])
c=d.get("clip32"); print({k:c[k] for k in c if k not in ("note","rank0_schedule","plan")})
PY
tail -3 gpurun_out/e_bench_n$N.err
