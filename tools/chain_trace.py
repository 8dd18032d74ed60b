"""Device timeline of ONE CTA of rv_conv_chain (REFVSR_CHAIN_TRACE hook in conv_chain.cu): prints, per role, the clock64
stamps of a few layers in the middle of a 20-layer trunk chain at 270x480 (or H W from argv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from refvsr_b200.lib import CudaOps
import test_gpu_kernels as tk
ops = CudaOps()
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (270, 480)
nblk = 10
bufs, layers, chain_l, conv_l, oi = tk._chain_case(ops, 48, H, W, torch.bfloat16, 'trunk', nblk)
flags = torch.empty((((H + 15) // 16) * ((W + 7) // 8),), dtype=torch.int32, device='cuda')
cl = [(chain_l[li], s, r, d, a0, a1) for li, s, r, d, a0, a1 in layers]
for _ in range(2):
    ops.conv_chain(bufs, cl, flags)
torch.cuda.synchronize()
trace = torch.zeros((7, 2048, 3), dtype=torch.int64, device='cuda')
os.environ['REFVSR_CHAIN_TRACE'] = hex(trace.data_ptr())
ops.conv_chain(bufs, cl, flags)
torch.cuda.synchronize()
del os.environ['REFVSR_CHAIN_TRACE']
t = trace.cpu()
t0 = int(t[:, :, 2][t[:, :, 2] > 0].min())
tend = int(t[:, :, 2].max())
G = min(148, ((H + 15) // 16) * ((W + 7) // 8))
cta = int(os.environ.get('REFVSR_CHAIN_TRACE_CTA', '70'))
ntl = (((H + 15) // 16) * ((W + 7) // 8) - cta + G - 1) // G
print('variant', os.environ.get('REFVSR_CHAIN_VARIANT'), f'geometry {H}x{W}, {len(cl)} layers, CTA {cta}: {ntl} tiles per layer; total {tend - t0} cycles = {(tend - t0) / len(cl):.0f} per layer')
names = ['producer (0 start,3 weights issued,1 deps ok,2 box issued)', 'mma0 (0 start,1 weights ok,2 acc free,3 box full,4 issued)', 'mma1', 'mma2',
         'epi0 (0 start,1 tfull,2 drained,3 math done,4 staged,5 store issued+prev published)', 'epi1', 'epi2']
lo, hi = 8 * ntl, (9 if os.environ.get('CHAIN_TRACE_SHORT') else 11) * ntl          # sequence numbers of layers 8..10
for role in range(7):
    ev = [(int(e), int(n), int(c) - t0) for e, n, c in t[role].tolist() if c > 0]
    print(f'== {names[role]}: {len(ev)} events')
    for n in range(lo, hi):
        seq = [(e, c) for e, nn, c in ev if nn == n]
        if seq:
            print(f'   n={n} (l={n // ntl},k={n % ntl}): ' + ' '.join(f'e{e}@{c}' for e, c in seq))
