"""shared helpers for the model-level tests"""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# must mirror tests/golden/make_golden.py::CASES (the generator needs /root/reference, the tests do not)
CASES = {
    'small_t3_32x48': dict(config='config_RefVSR_small_L1', over=dict(num_blocks=3, reset_branch=2), T=3, h=32, w=48,
                           ref_scale=1, frames=4, seed=11),
    'mfid_t5_40x56_ref2x': dict(config='config_RefVSR_MFID', over=dict(num_blocks=2), T=5, h=40, w=56,
                                ref_scale=2, frames=3, seed=12),
    'small_t7_24x32': dict(config='config_RefVSR_small_MFID', over=dict(num_blocks=1), T=7, h=24, w=32,
                           ref_scale=1, frames=5, seed=13),
    # flag_HD_in ("8K") configs: VGG[0:7] matching at 1/4 resolution, matching_ksize 8, aa1 scale 4 / aa2 scale 8
    'small8k_t3_32x48': dict(config='config_RefVSR_small_MFID_8K', over=dict(num_blocks=1, reset_branch=2), T=3, h=32, w=48,
                             ref_scale=1, frames=5, seed=21),
    'mfid8k_t5_48x64_noreset': dict(config='config_RefVSR_MFID_8K', over=dict(num_blocks=1), T=5, h=48, w=64,
                                    ref_scale=1, frames=3, seed=22),
}


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)        # trainers/trainer.py:252-254


def build_case(name, device, ops=None, **cfg_over):
    from refvsr_b200 import SRNet, get_config
    from refvsr_b200.modules import seeded_test_weights
    from refvsr_b200.synth import make_clip
    spec = CASES[name]
    over = dict(spec['over'])
    over.update(cfg_over)
    cfg = get_config(spec['config'], device=device, **over)
    net = SRNet(cfg).eval()
    seeded_test_weights(net, seed=spec['seed'])
    if ops is not None:
        net.Network.set_ops(ops)
    net = net.to(device)
    lrs, refs = make_clip(spec['frames'], spec['h'], spec['w'], spec['ref_scale'], seed=spec['seed'])
    golden = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return spec, cfg, net, lrs, refs, golden
