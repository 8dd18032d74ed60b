"""Device-side timeline of one conv_tc CTA (needs the experiments build):
    nvcc -DRV_CONV_EXPERIMENTS ... -o refvsr_b200/librefvsr_b200_exp.so ;  REFVSR_LIB=... python tools/conv_trace.py
Prints, per role (producer / MMA thread / epilogue group leaders), the clock64 deltas between events."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
exp = os.path.join(ROOT, 'refvsr_b200', 'librefvsr_b200_exp.so')
if not os.path.isfile(exp):
    src = [os.path.join(ROOT, 'refvsr_b200', 'csrc', f) for f in ('capi.cu', 'pointwise.cu', 'conv_simt.cu', 'conv_tc.cu', 'conv_rb.cu', 'conv_chain.cu', 'match.cu')]
    subprocess.run(['/usr/local/cuda/bin/nvcc', '-DRV_CONV_EXPERIMENTS', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17',
                    '-Xcompiler', '-fPIC', '-shared', '-o', exp] + src, check=True)
import torch
import refvsr_b200.lib as L
L._lib = None
L.load_library(exp)          # first load wins: every CudaOps below binds the experiments build
ops = L.CudaOps()
from refvsr_b200 import packing
from refvsr_b200.lib import ACT_RELU
H, W, C = (int(sys.argv[1]), int(sys.argv[2]), 48) if len(sys.argv) > 2 else (540, 960, 48)
print('geometry', H, W, C, 'layout', os.environ.get('REFVSR_TC_LAYOUT'), 'nmma', os.environ.get('REFVSR_NMMA'))
dt = torch.bfloat16
wgt = torch.randn(C, C, 3, 3) * 0.05
layer = packing.pack_conv('x', wgt, torch.zeros(C), [(C, C)], 1, 1, dt, 'cuda', True)
x = torch.randn((H, W, C), device='cuda').to(dt)
r = torch.randn((H, W, C), device='cuda').to(dt)
y = torch.empty((H, W, C), device='cuda', dtype=dt)
trace = torch.zeros((8, 1000, 3), dtype=torch.int64, device='cuda')
for i in range(3):
    ops.conv2d(layer, x, None, y, res=r, act_pre=ACT_RELU)
torch.cuda.synchronize()
os.environ['REFVSR_CONV_TRACE'] = hex(trace.data_ptr())
ops.conv2d(layer, x, None, y, res=r, act_pre=ACT_RELU)
torch.cuda.synchronize()
del os.environ['REFVSR_CONV_TRACE']
t = trace.cpu()
t0 = int(t[:, :, 2][t[:, :, 2] > 0].min())
names = {0: 'producer', 1: 'mma0', 2: 'epi0', 3: 'epi1', 4: 'epi2', 5: 'mma1', 6: 'mma2', 7: 'kernel marks (entry, prologue done, exit)'}
for role in range(8):
    ev = [(int(e), int(tile), int(c) - t0) for e, tile, c in t[role].tolist() if c > 0]
    print(f'== {names[role]}: {len(ev)} events; first {ev[0][2] if ev else None} last {ev[-1][2] if ev else None}')
    # per-tile summary for tiles 5..9 of this CTA (steady state)
    tiles = sorted(set(e[1] for e in ev))
    for tile in (tiles if len(tiles) <= 10 else tiles[4:9]):
        seq = [(e, c) for e, tl, c in ev if tl == tile]
        print(f'   tile {tile}: ' + ' '.join(f'e{e}@{c}' for e, c in seq))
    if len(tiles) > 6:
        starts = [min(c for e, tl, c in ev if tl == tile) for tile in tiles]
        per = [(starts[i + 1] - starts[i]) for i in range(len(starts) - 1)]
        print('   tile period (cycles): median', sorted(per)[len(per) // 2], 'first 8:', per[:8])
