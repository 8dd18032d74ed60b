"""Two independent trunk chains (forward / backward branch of a window) on two streams inside one CUDA graph, with the persistent
grid of each conv capped (rv_set_conv_cta_cap): per-conv time of the pair against one chain alone.  Usage: dual_chain_bench.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
import refvsr_b200.lib as L                     # noqa: E402
ops = L.CudaOps()
from refvsr_b200 import packing                 # noqa: E402
from refvsr_b200.lib import ACT_RELU            # noqa: E402

H, W, C = 270, 480, 48
dt = torch.bfloat16
dev = 'cuda'
gen = torch.Generator().manual_seed(0)
wa = (torch.rand((C, C, 3, 3), generator=gen) - 0.5) * 0.1
wb = (torch.rand((C, C, 3, 3), generator=gen) - 0.5) * 0.1
la = packing.pack_conv('a', wa, torch.zeros(C), [(C, C)], 1, 1, dt, dev, True)
lb = packing.pack_conv('b', wb, torch.zeros(C), [(C, C)], 1, 1, dt, dev, True)
A = [torch.randn((H, W, C), device=dev).to(dt) for _ in range(3)]
B = [torch.randn((H, W, C), device=dev).to(dt) for _ in range(3)]
NB = 30


def chain(layer, bufs):
    x, t, y = bufs
    for _ in range(NB):
        ops.conv2d(layer, x, None, t, act_pre=ACT_RELU)
        ops.conv2d(layer, t, None, y, res=x)
        x, y = y, x


def run(cap_a, cap_b, dual):
    side = torch.cuda.Stream()

    def body():
        if dual:
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ops.set_conv_cta_cap(cap_b)
                chain(lb, B)
            ops.set_conv_cta_cap(cap_a)
            chain(la, A)
            ops.set_conv_cta_cap(0)
            cur.wait_stream(side)
        else:
            ops.set_conv_cta_cap(cap_a)
            chain(la, A)
            ops.set_conv_cta_cap(0)
    body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    n = 2 * NB * (2 if dual else 1)
    print(f'{"dual" if dual else "single"} chains, caps {cap_a:3d}/{cap_b:3d}: {best:8.1f} us total, {best / n:6.2f} us per conv', flush=True)


run(0, 0, False)
run(74, 0, False)
run(0, 0, True)
for ca, cb in ((74, 74), (80, 68), (100, 48), (111, 37), (64, 64), (148, 74)):
    run(ca, cb, True)
