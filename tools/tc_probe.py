"""GPU diagnostic: structured probes of the tcgen05 conv / matching kernels that localise a failure
(TMA box placement, swizzle, descriptor K-advance, TMEM lane mapping, epilogue) from one run's log."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refvsr_b200 import packing
from refvsr_b200.lib import CudaOps, ACT_NONE
import torch.nn.functional as F

ops = CudaOps()


def ref_conv(xs, srcs, w, b, pad):
    x = torch.cat([x.float().permute(2, 0, 1)[:r] for x, (r, a) in zip(xs, srcs)], 0).unsqueeze(0)
    return F.conv2d(x, w, b, 1, pad)[0].permute(1, 2, 0)


LAYOUT = None


def probe(name, H, W, srcs, cout, k, wfun=None, dtype=torch.float16, xfun=None):
    g = torch.Generator().manual_seed(0)
    cin = sum(r for r, _ in srcs)
    w = (torch.rand((cout, cin, k, k), generator=g) - 0.5).to(dtype).float() if wfun is None else wfun(cout, cin, k)
    b = torch.zeros(cout)
    xs = [((torch.rand((H, W, a), generator=g) - 0.5) if xfun is None else xfun(H, W, a)).to(dtype) for _, a in srcs]
    exp = ref_conv(xs, srcs, w, b, k // 2)
    lc = packing.pack_conv(name, w, b, srcs, 1, k // 2, dtype, 'cuda', True, tc_layout=LAYOUT)
    out = torch.full((H, W, cout), float('nan'), dtype=dtype, device='cuda')
    try:
        ops.conv2d(lc, xs[0].cuda(), xs[1].cuda() if len(xs) > 1 else None, out)
        torch.cuda.synchronize()
    except Exception as e:
        print(f'[{name}] EXCEPTION {e}')
        return False
    o = out.float().cpu()
    err = (o - exp).abs()
    nanfrac = torch.isnan(o).float().mean().item()
    err = torch.nan_to_num(err, nan=1e9)
    tol = 2e-2 * max(1.0, exp.abs().max().item())
    bad = (err > tol)
    print(f'[{name}] H{H} W{W} srcs{srcs} cout{cout} k{k} nb{lc.nb} layout{lc.layout}: max err {err[~torch.isnan(o)].max().item() if nanfrac < 1 else float("nan"):.3e} '
          f'bad {bad.float().mean().item():.4f} nan {nanfrac:.4f}')
    if bad.any():
        ys, xs_, ns = bad.nonzero(as_tuple=True)
        print('   first bad (y,x,n):', list(zip(ys[:12].tolist(), xs_[:12].tolist(), ns[:12].tolist())))
        print('   bad rows y%8 hist:', torch.bincount(ys % 8, minlength=8).tolist(), ' x%16 hist:', torch.bincount(xs_ % 16, minlength=16).tolist())
        print('   bad n hist (per 8):', torch.bincount(ns // 8, minlength=(cout + 7) // 8).tolist())
        y, x = ys[0].item(), xs_[0].item()
        print('   got ', [round(v, 3) for v in o[y, x, :8].tolist()])
        print('   exp ', [round(v, 3) for v in exp[y, x, :8].tolist()])
    return not bad.any()


def ident(cout, cin, k):
    w = torch.zeros(cout, cin, k, k)
    for n in range(min(cout, cin)):
        w[n, n, k // 2, k // 2] = 1.0
    return w


def shift(dy, dx):
    def f(cout, cin, k):
        w = torch.zeros(cout, cin, k, k)
        for n in range(min(cout, cin)):
            w[n, n, k // 2 + dy, k // 2 + dx] = 1.0
        return w
    return f


def coords(H, W, a):
    y = torch.arange(H).view(H, 1, 1).expand(H, W, a)
    x = torch.arange(W).view(1, W, 1).expand(H, W, a)
    c = torch.arange(a).view(1, 1, a).expand(H, W, a)
    return (y * 0.01 + x * 0.001 + c * 0.1).float()


for LAYOUT in (0, 1, 2):
    print('=== layout', LAYOUT, '===')
    ok = True
    ok &= probe('id1x1_c16', 8, 16, [(16, 16)], 16, 1, ident, xfun=coords)
    ok &= probe('id1x1_c64', 8, 16, [(64, 64)], 64, 1, ident, xfun=coords)
    ok &= probe('id1x1_c48_big', 20, 40, [(48, 48)], 48, 1, ident, xfun=coords)
    ok &= probe('rand1x1_c64', 16, 32, [(64, 64)], 32, 1)
    ok &= probe('id3x3', 16, 32, [(48, 48)], 48, 3, ident, xfun=coords)
    for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        ok &= probe(f'shift{dy}{dx}', 16, 32, [(48, 48)], 48, 3, shift(dy, dx), xfun=coords)
    ok &= probe('rand3x3', 24, 48, [(48, 48)], 48, 3)
    ok &= probe('rand3x3_2src', 24, 48, [(8, 8), (48, 48)], 48, 3)
    ok &= probe('rand3x3_n192', 24, 48, [(48, 48)], 192, 3)
    ok &= probe('rand7x7', 24, 48, [(32, 32)], 64, 7)
    ok &= probe('rand3x3_bf16', 24, 48, [(48, 48)], 48, 3, dtype=torch.bfloat16)
    ok &= probe('rand3x3_270x480', 270, 480, [(48, 48)], 48, 3)
    print('TC_PROBE layout', LAYOUT, 'ALL_OK' if ok else 'FAILED')


# matching
from oracle.oracle_ops import OracleOps
oo = OracleOps()
g = torch.Generator().manual_seed(1)
for split in (False, True):
    lr_f = torch.rand((32, 48, 16), generator=g) - 0.5
    ref_f = torch.rand((20, 28, 16), generator=g) - 0.5
    kpad = 448 if split else 192
    P, R = 32 * 48, 20 * 28
    A, B = torch.zeros((P, kpad), dtype=torch.float16), torch.zeros((R, kpad), dtype=torch.float16)
    oo.patch_pack(lr_f, A, 1 if split else 0)
    oo.patch_pack(ref_f, B, 2 if split else 0)
    ce, ie = torch.zeros(P), torch.zeros(P, dtype=torch.int32)
    oo.match_argmax(A, B, ce, ie)
    for impl in (0, 1):
        c = torch.full((P,), float('nan'), device='cuda')
        i = torch.full((P,), -1, dtype=torch.int32, device='cuda')
        try:
            ops.match_argmax(A.cuda(), B.cuda(), c, i, impl=impl)
            torch.cuda.synchronize()
            print(f'[match split={split} impl={impl}] conf err {(c.cpu() - ce).abs().max().item():.3e} idx mismatch {(i.cpu() != ie).float().mean().item():.4f}'
                  f' got {c[:4].tolist()} {i[:4].tolist()} exp {ce[:4].tolist()} {ie[:4].tolist()}')
        except Exception as e:
            print(f'[match split={split} impl={impl}] EXCEPTION {e}')
