"""world_size-2 gloo run on CPU: frame-sharded processing (refvsr_b200/dist.py) with the input-halo
exchange equals the single-process stream bit for bit."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_segments():
    from refvsr_b200.dist import plan_segments
    assert plan_segments(32, 9, 8) == [(0, 9), (9, 18), (18, 27), (27, 32), (32, 32), (32, 32), (32, 32), (32, 32)]
    assert plan_segments(64, 9, 4) == [(0, 18), (18, 36), (36, 54), (54, 64)]
    assert plan_segments(5, 2, 2) == [(0, 4), (4, 5)]
    with pytest.raises(ValueError):
        plan_segments(10, None, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.oracle_ops import OracleOps
    from refvsr_b200.dist import exchange_halo, gather_frames, plan_segments, run_sharded
    from util import build_case
    spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=OracleOps(), b200_precision='fp32')
    T, n = spec['T'], lrs.shape[0]
    plan = plan_segments(n, cfg.reset_branch, world)
    f0, f1 = plan[rank]
    fl, first = exchange_halo(lrs[f0:f1].clone(), plan, rank, T // 2)      # each rank starts with ITS frames only
    fr, _ = exchange_halo(refs[f0:f1].clone(), plan, rank, T // 2)
    assert torch.equal(fl, lrs[first:first + fl.shape[0]])
    res = list(run_sharded(net, fl, fr, first, (f0, f1), n, T))
    allf = gather_frames(res, plan, rank, (3, 4 * spec['h'], 4 * spec['w']), 'cpu')
    if rank == 0:
        q.put([f.numpy() for f in allf])
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_stream():
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from oracle.oracle_ops import OracleOps
    from refvsr_b200.synth import sliding_windows
    from util import build_case
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=OracleOps(), b200_precision='fp32')
    single = [net(wl, wr, first)['result'][0].numpy() for k, wl, wr, first in sliding_windows(lrs, refs, spec['T'])]
    assert len(sharded) == len(single) == 4
    for a, b in zip(sharded, single):
        assert np.array_equal(a, b), 'segment-sharded output must be bit identical to the continuous stream'
