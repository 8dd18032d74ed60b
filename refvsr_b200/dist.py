"""Frame sharding of one clip over the GPUs of a box.

Where the path shards: with `reset_branch = R` the forward recurrence restarts every R calls
(RefVSR.py:168-170), and a restarted call depends on nothing but its own window; the backward branch is
window-local (RefVSR.py:211-214).  Hence output frames [sR, (s+1)R) form independent *segments*
(verified bit-exact against the reference: SURVEY appendix B) and a clip splits over ranks at segment
boundaries with NO collective on the data path.  The only exchange is an *input halo*: a rank that owns
(decoded / received) frames [f0, f1) needs T//2 neighbour frames on each side to build its first and last
windows.  `exchange_halo` moves exactly those frames with point-to-point send/recv (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Results can be gathered to rank 0 with `gather_frames`.

`reset_branch = None` (the 8K configs) has no segment boundaries: the forward chain is serial, and the
honest multi-GPU mode is independent replicas over different clips.
"""
import torch
import torch.distributed as dist


def plan_segments(num_frames, reset_branch, world_size):
    """-> list over ranks of (f0, f1): contiguous, reset-aligned ownership ranges covering [0, num_frames)."""
    if reset_branch is None:
        raise ValueError('reset_branch=None: the forward recurrence never restarts, a clip cannot be frame-sharded '
                         '(run independent replicas over different clips instead)')
    nseg = (num_frames + reset_branch - 1) // reset_branch
    per, extra = divmod(nseg, world_size)
    out, s = [], 0
    for r in range(world_size):
        n = per + (1 if r < extra else 0)
        out.append((min(s * reset_branch, num_frames), min((s + n) * reset_branch, num_frames)))
        s += n
    return out


def exchange_halo(local, plan, rank, halo, group=None):
    """`local`: frames owned by this rank, shape (f1-f0, ...).  Returns (frames, first_index) where `frames`
    additionally holds up to `halo` frames from the neighbouring ranks on each side."""
    world = len(plan)
    f0, f1 = plan[rank]
    num_frames = plan[-1][1]
    lo, hi = max(f0 - halo, 0), min(f1 + halo, num_frames)
    parts = {}
    ops, keep = [], []
    for other in range(world):
        if other == rank:
            continue
        o0, o1 = plan[other]
        # frames I need from `other`
        a, b = max(lo, o0), min(hi, o1)
        if a < b and (a < f0 or b > f1):
            a2, b2 = (a, min(b, f0)) if a < f0 else (max(a, f1), b)
            buf = torch.empty((b2 - a2,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            parts[a2] = buf
            ops.append(dist.P2POp(dist.irecv, buf, other, group))
        # frames `other` needs from me
        olo, ohi = max(o0 - halo, 0), min(o1 + halo, num_frames)
        a, b = max(olo, f0), min(ohi, f1)
        if a < b and (a < o0 or b > o1) and o1 > o0:
            a2, b2 = (a, min(b, o0)) if a < o0 else (max(a, o1), b)
            if a2 < b2:
                snd = local[a2 - f0:b2 - f0].contiguous()
                keep.append(snd)
                ops.append(dist.P2POp(dist.isend, snd, other, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts[f0] = local
    frames = torch.cat([parts[k] for k in sorted(parts)], 0) if len(parts) > 1 else local
    return frames, lo


def run_sharded(net, frames_lr, frames_ref, first_index, own, num_frames, T):
    """Process the owned output frames `own = (f0, f1)`.  `frames_*` hold clip frames
    [first_index, first_index + len) (own + halo).  Yields (k, result (3, 4h, 4w)).  The first owned frame
    starts a new stream (is_first_frame=True) - it sits on a reset boundary by construction."""
    f0, f1 = own
    for k in range(f0, f1):
        idx = [min(max(k - T // 2 + j, 0), num_frames - 1) - first_index for j in range(T)]   # datasets.py:233-234
        assert min(idx) >= 0 and max(idx) < frames_lr.shape[0], 'halo too small for the window'
        sel = torch.tensor(idx, device=frames_lr.device)
        out = net(frames_lr.index_select(0, sel).unsqueeze(0), frames_ref.index_select(0, sel).unsqueeze(0),
                  k == f0, False, False)['result']
        yield k, out[0]


def gather_frames(results, plan, rank, shape, device, dtype=torch.float32, group=None):
    """all ranks -> rank 0: list of num_frames tensors (rank 0) or None.  Uses all_gather on padded blocks
    (works with both NCCL and gloo)."""
    world = len(plan)
    mx = max(b - a for a, b in plan)
    mine = torch.zeros((mx,) + tuple(shape), dtype=dtype, device=device)
    for i, (k, r) in enumerate(sorted(results)):
        mine[i].copy_(r)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    if rank != 0:
        return None
    out = []
    for r, (a, b) in enumerate(plan):
        out.extend(bufs[r][i] for i in range(b - a))
    return out
