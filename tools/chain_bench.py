"""rv_conv_chain vs the same layers through rv_conv2d (graph-timed, per layer), plus a bit-exactness check.
    python tools/chain_bench.py            # prints one line per case"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from refvsr_b200.lib import CudaOps
import test_gpu_kernels as tk

ops = CudaOps()


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3     # us


cases = [('trunk LR 60', 48, 270, 480, 'trunk', 30), ('reslist LR 17', 48, 270, 480, 'reslist', 8), ('reslist 2x 9', 48, 540, 960, 'reslist', 4),
         ('reslist LR/2 9', 48, 135, 240, 'reslist', 4), ('trunk LR C24 48', 24, 270, 480, 'trunk', 24)]
out = {}
for name, C, H, W, kind, nblk in cases:
    bufs, layers, chain_l, conv_l, oi = tk._chain_case(ops, C, H, W, torch.bfloat16, kind, nblk)
    flags = torch.empty((((H + 15) // 16) * ((W + 7) // 8),), dtype=torch.int32, device='cuda')
    x0 = bufs[0].clone()
    tk._run_per_layer(ops, bufs, layers, conv_l)
    exp = bufs[oi].clone()
    bufs[0].copy_(x0)
    cl = [(chain_l[li], s, r, d, a0, a1) for li, s, r, d, a0, a1 in layers]
    ops.conv_chain(bufs, cl, flags)
    ok = bool(torch.equal(bufs[oi], exp))
    t_pl = timeit(lambda: tk._run_per_layer(ops, bufs, layers, conv_l)) / len(layers)
    t_ch = timeit(lambda: ops.conv_chain(bufs, cl, flags)) / len(layers)
    out[name] = dict(per_layer_us=round(t_pl, 2), chain_us=round(t_ch, 2), exact=ok)
    print(f'{name:18s} per-layer kernels {t_pl:7.2f} us/layer   chain {t_ch:7.2f} us/layer   x{t_pl / t_ch:5.2f}   exact={ok}', flush=True)
print('variant', os.environ.get('REFVSR_CHAIN_VARIANT'), json.dumps(out))
