"""Full-size golden vectors at the BASELINE configurations (270x480 LR -> 1080p), produced in the build container.

    python tests/golden/make_fullsize.py mfid        # configs[2]: the UNMODIFIED reference (/root/reference), 30 blocks, T=7
    python tests/golden/make_fullsize.py small_mfid  # configs[1]: Ref 540x960 - the reference's 129600^2 fp32 similarity
                                                     # matrix (67 GB) does not fit this container, so this case comes from the
                                                     # oracle port (oracle/refvsr_oracle.py, pinned to the reference at fixture sizes)

Outputs:
  tests/golden_large/<case>.npz   full fp32 results + first-window flows / conf / idx   (git-ignored: ~100 MB; it
                                  travels to the GPU box with the working tree, like the built .so)
  tests/golden/<case>_digest.npz  committed: 8x8 block means of every result + a 128x128 centre crop + per-window
                                  sums, enough for a (weaker) check when the large file is absent.

`reset_branch` is overridden to 2 so that four calls cover: first window, steady window, forced-reset window (RefVSR.py:168-170),
steady window after a reset.  Inputs: refvsr_b200.synth.make_clip(6, 270, 480, ref_scale, seed=1234); weights:
seeded_test_weights(seed=1234).
"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FULL_CASES = {
    'mfid_270x480': dict(config='config_RefVSR_MFID', over=dict(reset_branch=2), T=7, h=270, w=480, ref_scale=1,
                         frames=6, windows=4, seed=1234, source='reference'),
    'small_mfid_270x480_ref2x': dict(config='config_RefVSR_small_MFID', over=dict(reset_branch=2), T=7, h=270, w=480,
                                     ref_scale=2, frames=6, windows=3, seed=1234, source='oracle'),
}


def digest(res):
    """(3,H,W) fp32 -> block means (3,H/8,W/8), centre crop (3,128,128)"""
    t = torch.from_numpy(res)
    bm = torch.nn.functional.avg_pool2d(t.unsqueeze(0), 8)[0].numpy()
    H, W = res.shape[1:]
    crop = res[:, H // 2 - 64:H // 2 + 64, W // 2 - 64:W // 2 + 64].copy()
    return bm, crop


def run_reference(spec):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from make_golden import load_reference
    from refvsr_b200.modules import seeded_test_weights
    RefSRNet = load_reference()
    cfg = importlib.import_module('configs.' + spec['config']).get_config('p', 'm', spec['config'])
    cfg.cuda, cfg.device, cfg.dist = False, 'cpu', False
    for k, v in spec['over'].items():
        setattr(cfg, k, v)
    torch.manual_seed(0)
    ref = RefSRNet(cfg).eval()
    seeded_test_weights(ref, seed=spec['seed'])
    net = ref.Network
    rec = {}
    orig_flow, orig_match = net.FlowNet.forward, net.feature_match.forward

    def flow_hook(a, b):
        r = orig_flow(a, b)
        rec.setdefault('flows', []).append(r.clone())
        return r

    def match_hook(lr, rf, *a, **k):
        conf, idx = orig_match(lr, rf, *a, **k)
        rec.setdefault('conf', []).append(conf.clone())
        rec.setdefault('idx', []).append(idx.clone())
        return conf, idx

    net.FlowNet.forward = flow_hook
    net.feature_match.forward = match_hook

    def call(wl, wr, first):
        rec.clear()
        out = ref(wl, wr, first, False, False)['result'][0]
        inter = None
        if first:
            inter = (torch.cat(rec['flows'], 0), torch.cat(rec['conf'], 0), torch.cat(rec['idx'], 0))
        return out, inter
    return call


def run_oracle(spec):
    from oracle.refvsr_oracle import OracleRefVSR
    from refvsr_b200 import SRNet, get_config
    from refvsr_b200.modules import seeded_test_weights
    cfg = get_config(spec['config'], device='cpu', **spec['over'])
    holder = SRNet(cfg)
    seeded_test_weights(holder, seed=spec['seed'])
    net = OracleRefVSR(cfg, holder.state_dict())

    def call(wl, wr, first):
        net.trace = {} if first else None
        out = net.forward(wl, wr, first)[0]
        inter = None
        if first:
            tr = net.trace
            t = wl.shape[1]
            # consumed flows only (the oracle skips the ones the reference never reads): forward j, then backward j
            flows = torch.cat([tr['fw'][j] for j in sorted(tr['fw'])] + [tr['bw'][j] for j in sorted(tr['bw'])], 0)
            conf = torch.cat([tr['conf'][i] for i in range(t)], 0)
            idx = torch.cat([tr['idx'][i] for i in range(t)], 0)
            inter = (flows, conf, idx)
        return out, inter
    return call


def main(name):
    from refvsr_b200.synth import make_clip, sliding_windows
    spec = FULL_CASES[name]
    call = run_reference(spec) if spec['source'] == 'reference' else run_oracle(spec)
    lrs, refs = make_clip(spec['frames'], spec['h'], spec['w'], spec['ref_scale'], seed=spec['seed'])
    big, small = {}, {}
    with torch.no_grad():
        for k, wl, wr, first in sliding_windows(lrs, refs, spec['T']):
            if k >= spec['windows']:
                break
            t0 = time.time()
            out, inter = call(wl, wr, first)
            res = out.numpy().astype(np.float32)
            print(f'{name} window {k}: {time.time() - t0:.1f} s, mean {res.mean():.6f}', flush=True)
            big[f'result_{k}'] = res
            bm, crop = digest(res)
            small[f'blockmean_{k}'] = bm
            small[f'crop_{k}'] = crop
            small[f'sum_{k}'] = np.array([res.astype(np.float64).sum(), (res.astype(np.float64) ** 2).sum()])
            if inter is not None and k == 0:
                big['flows_0'] = inter[0].numpy().astype(np.float32)
                big['conf_0'] = inter[1].numpy().astype(np.float32)
                big['idx_0'] = inter[2].numpy().astype(np.int32)
    meta = np.array([spec['T'], spec['h'], spec['w'], spec['ref_scale'], spec['frames'], spec['seed'], spec['windows']])
    big['meta'] = small['meta'] = meta
    os.makedirs(os.path.join(ROOT, 'tests', 'golden_large'), exist_ok=True)
    p1 = os.path.join(ROOT, 'tests', 'golden_large', name + '.npz')
    np.savez(p1, **big)
    p2 = os.path.join(ROOT, 'tests', 'golden', name + '_digest.npz')
    np.savez_compressed(p2, **small)
    print('written', p1, os.path.getsize(p1) >> 20, 'MiB;', p2, os.path.getsize(p2) >> 10, 'KiB')


if __name__ == '__main__':
    torch.set_num_threads(int(os.environ.get('GOLDEN_THREADS', '6')))
    for n in sys.argv[1:]:
        main({'mfid': 'mfid_270x480', 'small_mfid': 'small_mfid_270x480_ref2x'}.get(n, n))
