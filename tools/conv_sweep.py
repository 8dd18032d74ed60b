"""GPU experiment: graph-timed conv_tc launches over channel counts / kernel sizes / epilogue options."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refvsr_b200 import packing
from refvsr_b200.lib import CudaOps, ACT_RELU
ops = CudaOps()
H, W = 270, 480
dt = torch.bfloat16

def timeit(fn, iters=20, warm=3):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def case(cin, cout, k, res=True, calloc=None, hw=(H, W)):
    h, w = hw
    calloc = calloc or cin
    wgt = torch.randn(cout, cin, k, k) * 0.05
    layer = packing.pack_conv('x', wgt, torch.zeros(cout), [(cin, calloc)], 1, k // 2, dt, 'cuda', True)
    nrot = 6
    xs = [torch.randn((h, w, calloc), device='cuda').to(dt) for _ in range(nrot)]
    rs = [torch.randn((h, w, cout), device='cuda').to(dt) for _ in range(nrot)]
    ys = [torch.empty((h, w, cout), device='cuda', dtype=dt) for _ in range(nrot)]
    t = timeit(lambda i: ops.conv2d(layer, xs[i % nrot], None, ys[i % nrot], res=rs[i % nrot] if res else None, act_pre=ACT_RELU))
    fl = 2.0 * k * k * cin * cout * h * w
    print(f'cin {cin:3d} (alloc {calloc:3d}) cout {cout:3d} k{k} res={int(res)} {h}x{w}: {t:7.1f} us  {fl / t / 1e6:7.1f} TFLOP/s', flush=True)

case(48, 48, 3)
case(48, 48, 3, res=False)
case(64, 64, 3)
case(64, 64, 3, res=False)
case(48, 48, 3, calloc=64)
case(32, 32, 3)
case(16, 16, 3)
case(48, 48, 1)
case(64, 64, 1)
case(48, 48, 5)
case(64, 64, 7)
case(32, 64, 7)
case(48, 48, 3, hw=(540, 960))
case(64, 64, 3, hw=(540, 960))
case(48, 192, 3)
case(96, 48, 3)
