"""GPU diagnostic for rv_resblock: structured weights localise failures (residual path, intermediate layout,
tap shifts of conv1 / conv2, tile borders)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from refvsr_b200 import packing
from refvsr_b200.lib import CudaOps, ACT_NONE, ACT_RELU
ops = CudaOps()


def ident(C, dy=0, dx=0, gain=1.0):
    w = torch.zeros(C, C, 3, 3)
    for n in range(C):
        w[n, n, 1 + dy, 1 + dx] = gain
    return w


def coords(H, W, C):
    y = torch.arange(H).view(H, 1, 1).expand(H, W, C)
    x = torch.arange(W).view(1, W, 1).expand(H, W, C)
    c = torch.arange(C).view(1, 1, C).expand(H, W, C)
    return (y * 0.01 + x * 0.001 + c * 0.1).float()


def run(name, H, W, C, w1, w2, b1=None, b2=None, act=ACT_NONE, x=None, dt=torch.float16):
    b1 = torch.zeros(C) if b1 is None else b1
    b2 = torch.zeros(C) if b2 is None else b2
    x = (coords(H, W, C) if x is None else x).to(dt)
    xn = x.float().permute(2, 0, 1).unsqueeze(0)
    t = F.conv2d(xn, w1, b1, 1, 1)
    if act == ACT_RELU:
        t = torch.relu(t)
    t = t.to(dt).float()
    exp = (xn + F.conv2d(t, w2, b2, 1, 1))[0].permute(1, 2, 0)
    rb = packing.pack_resblock(name, w1, b1, w2, b2, C, dt, 'cuda')
    out = torch.full((H, W, C), float('nan'), dtype=dt, device='cuda')
    try:
        ops.resblock(rb, x.cuda(), out, act)
        torch.cuda.synchronize()
    except Exception as e:
        print(f'[{name}] EXCEPTION {e}')
        return False
    o = out.float().cpu()
    nan = torch.isnan(o).float().mean().item()
    err = torch.nan_to_num((o - exp).abs(), nan=1e9)
    bad = err > 2e-2 * max(1.0, exp.abs().max().item())
    print(f'[{name}] {H}x{W}x{C}: max err {err[err < 1e8].max().item() if nan < 1 else float("nan"):.3e} bad {bad.float().mean().item():.4f} nan {nan:.4f}')
    if bad.any():
        ys, xs, ns = bad.nonzero(as_tuple=True)
        print('   first bad (y,x,n):', list(zip(ys[:10].tolist(), xs[:10].tolist(), ns[:10].tolist())))
        print('   bad y%16 hist:', torch.bincount(ys % 16, minlength=16).tolist(), ' x%8 hist:', torch.bincount(xs % 8, minlength=8).tolist())
        y, x_ = ys[0].item(), xs[0].item()
        print('   got ', [round(v, 3) for v in o[y, x_, :8].tolist()])
        print('   exp ', [round(v, 3) for v in exp[y, x_, :8].tolist()])
    return not bad.any()


C = 48
Z = torch.zeros(C, C, 3, 3)
ok = True
ok &= run('residual_only (w2=0)', 32, 24, C, ident(C), Z)
ok &= run('bias2_only', 32, 24, C, Z, Z, b2=torch.arange(C).float() * 0.01)
ok &= run('mid=bias1 -> id conv2', 32, 24, C, Z, ident(C), b1=torch.ones(C) * 0.5)
ok &= run('id,id', 32, 24, C, ident(C), ident(C))
for d in ((-1, 0), (1, 0), (0, -1), (0, 1)):
    ok &= run(f'conv1 shift{d}, conv2 id', 32, 24, C, ident(C, *d), ident(C))
for d in ((-1, 0), (1, 0), (0, -1), (0, 1)):
    ok &= run(f'conv1 id, conv2 shift{d}', 32, 24, C, ident(C), ident(C, *d))
g = torch.Generator().manual_seed(0)
ok &= run('random relu', 40, 28, C, (torch.rand(C, C, 3, 3, generator=g) - 0.5) * 0.2, (torch.rand(C, C, 3, 3, generator=g) - 0.5) * 0.2,
          b1=torch.rand(C, generator=g) - 0.5, b2=torch.rand(C, generator=g) - 0.5, act=ACT_RELU, x=torch.rand(40, 28, C, generator=g) - 0.5)
ok &= run('random 270x480 bf16', 270, 480, C, (torch.rand(C, C, 3, 3, generator=g) - 0.5) * 0.2, (torch.rand(C, C, 3, 3, generator=g) - 0.5) * 0.2,
          act=ACT_RELU, x=torch.rand(270, 480, C, generator=g) - 0.5, dt=torch.bfloat16)
print('RB_PROBE', 'ALL_OK' if ok else 'FAILED')
