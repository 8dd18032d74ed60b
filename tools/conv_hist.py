"""GPU tool: histogram of rv_conv2d geometries in one steady-state window of RefVSR_MFID 270x480, each distinct
geometry then graph-timed in isolation (same buffers) -> where the conv time of a window goes."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cfg, net = bench.make_model('mfid', None, torch.device('cuda', 0), graphs=False)
from refvsr_b200.synth import make_clip_range
lrs, refs = make_clip_range(0, 8, bench.H, bench.W, 1, seed=1234)
lrs, refs = lrs.cuda(), refs.cuda()
ops = net.Network.ops
orig = ops.conv2d
calls = []
rec = [False]
def spy(layer, src0, src1, out, **kw):
    if rec[0]:
        calls.append((layer, src0, src1, out, kw))
    return orig(layer, src0, src1, out, **kw)
ops.conv2d = spy
for k in range(4):
    rec[0] = (k == 3)
    ids = torch.tensor(bench.window_indices(k, 8), device='cuda')
    out = net(lrs.index_select(0, ids).unsqueeze(0), refs.index_select(0, ids).unsqueeze(0), k == 0, False, False)
torch.cuda.synchronize()
ops.conv2d = orig
groups = collections.OrderedDict()
for c in calls:
    layer, s0, s1, o, kw = c
    key = (layer.impl, getattr(layer, 'layout', 0), layer.kh, tuple(s0.shape), tuple(s1.shape) if s1 is not None else None,
           layer.cout, layer.nb, kw.get('res') is not None, kw.get('gate') is not None, bool(kw.get('pixel_shuffle')), str(o.dtype))
    groups.setdefault(key, []).append(c)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

rows = []
for key, cs in groups.items():
    layer, s0, s1, o, kw = cs[0]
    t = timeit(lambda: orig(layer, s0, s1, o, **kw))
    rows.append((t * len(cs), len(cs), t, key))
tot = sum(r[0] for r in rows)
print(f'{len(calls)} conv launches in a steady window, {len(rows)} geometries, sum of isolated times {tot / 1e3:.2f} ms')
print('total_us count each_us  impl layout k src0 src1 cout nb res gate ps out')
for r in sorted(rows, key=lambda r: -r[0]):
    print(f'{r[0]:8.0f} {r[1]:4d} {r[2]:7.1f}  ' + ' '.join(str(x) for x in r[3]))
