"""Eager (no CUDA graph) run of 1 first + 3 steady windows of RefVSR_MFID 270x480 for `ncu --metrics gpu__time_duration.sum`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cfg, net = bench.make_model('mfid', None, torch.device('cuda', 0), graphs=False)
from refvsr_b200.synth import make_clip_range
lrs, refs = make_clip_range(0, 8, bench.H, bench.W, 1, seed=1234)
lrs, refs = lrs.cuda(), refs.cuda()
for k in range(4):
    ids = torch.tensor(bench.window_indices(k, 8), device='cuda')
    out = net(lrs.index_select(0, ids).unsqueeze(0), refs.index_select(0, ids).unsqueeze(0), k == 0, False, False)
torch.cuda.synchronize()
print('done', tuple(out['result'].shape))
