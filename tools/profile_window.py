"""Eager (no CUDA graph) run of RefVSR_MFID 270x480 for `ncu`: 1 first + 3 steady windows; only the LAST steady window
sits between cudaProfilerStart / Stop, so `ncu --profile-from-start off` lists exactly one steady-state step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cfg, net = bench.make_model('mfid', None, torch.device('cuda', 0), graphs=False)
from refvsr_b200.synth import make_clip_range
lrs, refs = make_clip_range(0, 8, bench.H, bench.W, 1, seed=1234)
lrs, refs = lrs.cuda(), refs.cuda()
NW = 4
for k in range(NW):
    ids = torch.tensor(bench.window_indices(k, 8), device='cuda')
    a, b = lrs.index_select(0, ids).unsqueeze(0), refs.index_select(0, ids).unsqueeze(0)
    torch.cuda.synchronize()
    if k == NW - 1:
        torch.cuda.profiler.start()
    out = net(a, b, k == 0, False, False)
    torch.cuda.synchronize()
    if k == NW - 1:
        torch.cuda.profiler.stop()
print('done', tuple(out['result'].shape))
