"""torchrun --nproc-per-node N tools/dist_check.py : frame-sharded clip (refvsr_b200.dist.run_clip_frame_sharded, NCCL) against the
single-stream windowed forward on rank 0 - must be bit-identical (same kernels, same order per tile; the hand-off moves bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from refvsr_b200 import SRNet, get_config
from refvsr_b200.dist import exchange_halo, gather_frames, plan_frames, run_clip_frame_sharded
from refvsr_b200.modules import seeded_test_weights
from refvsr_b200.synth import make_clip, sliding_windows

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
groups = [dist.new_group(list(range(world))), dist.new_group(list(range(world)))]
ok = True
for (h, w, n, nb, prec) in ((96, 128, 23, 4, 'bf16'), (270, 480, 12, 30, 'bf16')):
    cfg = get_config('config_RefVSR_MFID', device='cuda', num_blocks=nb, b200_precision=prec)
    net = SRNet(cfg).eval()
    seeded_test_weights(net, seed=9)
    net = net.to(dev)
    net.Network.chain_max_ctas = 132
    lrs, refs = make_clip(n, h, w, 1, seed=17)
    T = cfg.frame_num
    plan = plan_frames(n, world)
    f0, f1 = plan[rank]
    for rep in range(2):
        fl, first = exchange_halo(lrs[f0:f1].to(dev), plan, rank, T // 2)
        fr, _ = exchange_halo(refs[f0:f1].to(dev), plan, rank, T // 2)
        timing = {}
        res = run_clip_frame_sharded(net, fl, fr, first, plan, rank, groups=groups, timing=timing)
        torch.cuda.synchronize()
    allf = gather_frames(res, plan, rank, (3, 4 * h, 4 * w), dev)
    if rank == 0:
        net.Network.reset_state()
        worst, nbad = 0.0, 0
        for k, wl, wr, first_ in sliding_windows(lrs, refs, T):
            o = net(wl.to(dev), wr.to(dev), first_, False, False)['result'][0]
            d = float((o - allf[k]).abs().max())
            worst = max(worst, d)
            nbad += int(d != 0.0)
        print(f'[dist_check] N={world} {h}x{w} {n} frames {nb} blocks: max |sharded - single stream| = {worst:.3e}, frames differing: {nbad}; '
              f'rank0 schedule: {timing["log"][:6]}', flush=True)
        ok = ok and nbad == 0
    dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
