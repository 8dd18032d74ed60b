"""Host logic of the product (refvsr_b200/network.py) on CPU: the schedule is run through the OracleOps test
double (oracle/oracle_ops.py) and compared with the reference's golden vectors; plus the drop-in contract
(state_dict schema, module API, state handling, error behaviour).  No CUDA needed."""
import collections
import json
import os

import numpy as np
import pytest
import torch

from util import CASES, GOLDEN_DIR, build_case, psnr


@pytest.mark.parametrize('name', list(CASES))
def test_schedule_matches_reference_golden(name, oracle_ops):
    from refvsr_b200.synth import sliding_windows
    spec, cfg, net, lrs, refs, golden = build_case(name, 'cpu', ops=oracle_ops, b200_precision='fp32')
    for k, wl, wr, first in sliding_windows(lrs, refs, spec['T']):
        outs = net(wl, wr, first, False, False)
        assert isinstance(outs, collections.OrderedDict) and list(outs) == ['result']
        o = outs['result']
        assert o.shape == (1, 3, 4 * spec['h'], 4 * spec['w']) and o.dtype == torch.float32
        g = torch.from_numpy(golden[f'result_{k}'])
        # an exact score tie may flip one argmax (fp16-split product vs fp32 bmm): PSNR bar + loose max
        assert psnr(o[0], g) > 95.0 and (o[0] - g).abs().max() < 2e-3
        if k == 0:
            N = net.Network
            mism = np.mean([(N._frame_slot(i % spec['T'], spec['h'], spec['w'])['idx'].numpy() != golden['idx_0'][i]).mean()
                            for i in range(spec['T'])])
            assert mism <= 1e-3


def test_fused_resblock_schedule_matches_reference_golden(oracle_ops):
    """the opt-in fused residual-block route (ops.resblock) computes the same network"""
    from refvsr_b200.synth import sliding_windows
    spec, cfg, net, lrs, refs, golden = build_case('small_t7_24x32', 'cpu', ops=oracle_ops, b200_precision='fp32',
                                                   b200_fuse_resblocks=True)
    net.Network.act_dtype = torch.float32
    calls = {'n': 0}
    orig = oracle_ops.resblock

    def counting(*a, **k):
        calls['n'] += 1
        return orig(*a, **k)
    oracle_ops.resblock = counting
    try:
        import refvsr_b200.packing as packing
        old = packing.resblock_fusable
        packing.resblock_fusable = lambda w1, w2, alloc, dt: tuple(w1.shape) == tuple(w2.shape) and w1.shape[2] == 3 and alloc <= 64
        for k, wl, wr, first in sliding_windows(lrs, refs, spec['T']):
            o = net(wl, wr, first)['result']
            assert psnr(o[0], torch.from_numpy(golden[f'result_{k}'])) > 95.0
    finally:
        oracle_ops.resblock = orig
        packing.resblock_fusable = old
    assert calls['n'] > 50


def test_reuse_on_off_and_unused_flows(oracle_ops):
    from refvsr_b200.synth import sliding_windows
    res = {}
    for reuse in (True, False):
        spec, cfg, net, lrs, refs, golden = build_case('small_t7_24x32', 'cpu', ops=oracle_ops, b200_precision='fp32',
                                                       b200_reuse=reuse)
        res[reuse] = [net(wl, wr, first)['result'] for k, wl, wr, first in sliding_windows(lrs, refs, spec['T'])]
        st = net.Network._state[0]
        if reuse:   # steady state keeps only what the next window can still use
            assert len(st['frame']) <= spec['T'] and len(st['fw']) <= spec['T'] and len(st['bw']) <= spec['T']
    for a, b in zip(res[True], res[False]):
        assert torch.equal(a, b)


def test_state_dict_schema_matches_reference():
    """key names, order-insensitive, and shapes: fixtures dumped from the reference by make_golden's sibling
    (tests/golden/state_dict_keys.json); falls back to counts from SURVEY 8b."""
    from refvsr_b200 import SRNet, get_config
    counts = {'config_RefVSR_small_MFID': (428, 2492070), 'config_RefVSR_MFID': (476, 5717550),
              'config_RefVSR_L1': (476, 5717550), 'config_RefVSR_small_L1': (428, 2492070),
              'config_RefVSR_small_MFID_8K': (444, 2657705), 'config_RefVSR_MFID_8K': (492, 5883185)}
    path = os.path.join(GOLDEN_DIR, 'state_dict_keys.json')
    ref_keys = json.load(open(path)) if os.path.isfile(path) else {}
    for name, (ntens, nparam) in counts.items():
        net = SRNet(get_config(name, device='cpu'))
        sd = net.state_dict()
        assert len(sd) == ntens and sum(p.numel() for p in net.parameters()) == nparam
        assert all(k.startswith('Network.') for k in sd)
        if name in ref_keys:
            assert {k: list(v.shape) for k, v in sd.items()} == ref_keys[name]
    # DataParallel-style checkpoints carry a 'module.' prefix that ckpt_manager.py:50-56 strips on CPU
    net2 = SRNet(get_config('config_RefVSR_small_MFID', device='cpu'))
    missing = net2.load_state_dict({k: v for k, v in net2.state_dict().items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_load_state_dict_invalidates_packed_weights(oracle_ops):
    from refvsr_b200.modules import seeded_test_weights
    spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=oracle_ops, b200_precision='fp32')
    wl, wr = lrs[[0, 0, 1]].unsqueeze(0), refs[[0, 0, 1]].unsqueeze(0)
    a = net(wl, wr, True)['result'].clone()
    assert len(net.Network._packed) > 0
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    seeded_test_weights(net, seed=99)
    net.load_state_dict(net.state_dict())
    assert len(net.Network._packed) == 0
    b = net(wl, wr, True)['result'].clone()
    assert (a - b).abs().max() > 1e-4
    net.load_state_dict(sd)
    c = net(wl, wr, True)['result']
    assert torch.equal(a, c)


def test_forward_contract_and_errors(oracle_ops):
    spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=oracle_ops, b200_precision='fp32')
    wl, wr = lrs[[0, 0, 1]].unsqueeze(0), refs[[0, 0, 1]].unsqueeze(0)
    with pytest.raises(RuntimeError):        # steady-state call without propagated state (RefVSR.py:256-260)
        net(wl, wr, False)
    with pytest.raises(ValueError):
        net(wl[0], wr[0], True)
    with pytest.raises(ValueError):
        net(wl, wr[:, :, :, :31], True)      # odd reference height
    # is_log returns the optional dicts (RefVSR.py:163-164,222,263)
    outs = net(wl, wr, True, True, False)
    assert 'vis' in outs and 'BW_LR_next_warp' in outs['vis'] and outs['vis']['BW_LR_next_warp'].shape == (1, 3, 32, 48)
    # training-mode call: no clamp, frame counter untouched (RefVSR.py:292-297)
    n0 = net.Network.frame_itr_num
    with torch.no_grad():                # autograd-enabled training forwards raise (inference-only product)
        o = net(wl, wr, True, False, True)["result"]
    assert net.Network.frame_itr_num == n0 and o.shape == (1, 3, 128, 192)
    # batch of 2 == two independent streams
    o2 = net(torch.cat([wl, wl], 0), torch.cat([wr, wr], 0), True)['result']
    assert o2.shape[0] == 2 and torch.equal(o2[0], o2[1])
    # input_constructor (ptflops hook, SRNet.py:47-54)
    d = net.input_constructor((1, 3, 3, 16, 16))
    assert set(d) == {'x', 'ref'} and d['x'].shape == (1, 3, 3, 16, 16)


def test_hd_input_size_contract(oracle_ops):
    spec, cfg, net, lrs, refs, golden = build_case('small8k_t3_32x48', 'cpu', ops=oracle_ops, b200_precision='fp32')
    x = torch.rand(1, 3, 3, 30, 48)
    with pytest.raises(ValueError, match='flag_HD_in'):
        net(x, x, True)


def test_reset_branch_counter(oracle_ops):
    """forced first-frame windows at call indices 0, reset_branch, 2*reset_branch, ... (RefVSR.py:168-170)"""
    from refvsr_b200.synth import sliding_windows
    spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=oracle_ops, b200_precision='fp32')
    seen = []
    for k, wl, wr, first in sliding_windows(lrs, refs, spec['T']):
        before = net.Network.frame_itr_num
        net(wl, wr, first)
        seen.append((before, net.Network.frame_itr_num))
    assert seen == [(0, 1), (1, 2), (2, 1), (1, 2)]


def test_unsupported_configs_fail_loudly():
    from refvsr_b200 import SRNet, get_config
    with pytest.raises(NotImplementedError):
        SRNet(get_config('config_RefVSR_MFID', device='cpu', scale=2))          # the x2 models are not on this path
    SRNet(get_config('config_RefVSR_MFID_8K', device='cpu', num_blocks=1))       # flag_HD_in is (schedule: HD goldens above)


def test_product_has_no_cpu_fallback():
    """without CUDA the default operator set must refuse to run (no silent oracle / eager path)."""
    from refvsr_b200 import SRNet, get_config
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    net = SRNet(get_config('config_RefVSR_small_MFID', device='cpu', num_blocks=1))
    x = torch.rand(1, 3, 3, 16, 16)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(x, x, True)
    # and nothing under refvsr_b200/ imports the oracle
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'refvsr_b200')
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_standalone_configs_match_the_reference_config_modules():
    """refvsr_b200.config.get_config restates configs/config_RefVSR_*.py; fixture dumped from the reference by
    tests/golden/make_config_values.py"""
    import json
    from refvsr_b200 import get_config
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'config_values.json')) as f:
        ref = json.load(f)
    assert len(ref) == 6
    for name, fields in ref.items():
        cfg = get_config(name, device='cpu')
        for k, v in fields.items():
            assert cfg[k] == v, (name, k, cfg[k], v)


def test_reuse_guard_non_sliding_caller(oracle_ops):
    """ADVICE r1 (medium) / VERDICT r1 item 8: a caller that passes is_first_frame=False but does NOT slide the window by
    one frame must get what the reference computes for those calls (sync guard: full recompute), or a loud error (async
    guard) - never silently stale products."""
    from oracle.refvsr_oracle import OracleRefVSR
    from refvsr_b200.synth import sliding_windows
    spec, cfg, net, lrs, refs, golden = build_case('small_t7_24x32', 'cpu', ops=oracle_ops, b200_precision='fp32')
    orc = OracleRefVSR(cfg, net.state_dict())
    wins = list(sliding_windows(lrs, refs, spec['T']))
    order = [0, 2, 3, 1]                      # 0 -> 2 jumps by two frames, 3 slides, 1 jumps backwards
    for n, k in enumerate(order):
        _, wl, wr, _ = wins[k]
        out = net(wl, wr, n == 0, False, False)['result']
        exp = orc.forward(wl, wr, n == 0)
        assert psnr(out[0], exp[0]) > 95.0, f'call {n} (window {k})'
    assert net.Network.reuse_fallbacks == 2
    # async guard: the violation surfaces as an exception on the following call
    spec, cfg, net, lrs, refs, golden = build_case('small_t7_24x32', 'cpu', ops=oracle_ops, b200_precision='fp32',
                                                   b200_reuse_check='async')
    net(wins[0][1], wins[0][2], True)
    net(wins[2][1], wins[2][2], False)        # violation: not detected yet
    with pytest.raises(RuntimeError, match='did not slide'):
        net(wins[3][1], wins[3][2], False)


def test_training_mode_forward_fails_loudly(oracle_ops):
    """is_train=True with autograd enabled would silently return a tensor without grad_fn"""
    spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=oracle_ops, b200_precision='fp32')
    wl, wr = lrs[[0, 0, 1]].unsqueeze(0), refs[[0, 0, 1]].unsqueeze(0)
    with pytest.raises(NotImplementedError, match='inference only'):
        net(wl, wr, True, False, True)
    with torch.no_grad():
        assert net(wl, wr, True, False, True)['result'].shape[1] == 3


def test_conv_chain_schedule_is_the_per_layer_schedule(oracle_ops, monkeypatch):
    """network.py routes the propagation trunks and the ResList decoders through ops.conv_chain (one persistent launch on
    the GPU); the buffer / residual indexing of those chains must describe exactly the per-layer network"""
    import refvsr_b200.packing as packing
    from refvsr_b200.synth import sliding_windows
    res = {}
    for chain in (False, True):
        spec, cfg, net, lrs, refs, golden = build_case('small_t3_32x48', 'cpu', ops=oracle_ops, b200_precision='fp32',
                                                       b200_conv_chain=chain)
        calls = {'n': 0, 'layers': 0}
        orig = oracle_ops.conv_chain

        def counting(bufs, layers, flags, max_ctas=0, _o=orig, _c=calls):
            _c['n'] += 1
            _c['layers'] += len(layers)
            return _o(bufs, layers, flags)
        monkeypatch.setattr(oracle_ops, 'conv_chain', counting)
        monkeypatch.setattr(packing, 'chain_ok', lambda w, alloc, dt: tuple(w.shape[1:]) == (w.shape[0], 3, 3) and alloc <= 48)
        res[chain] = [net(wl, wr, first)['result'] for k, wl, wr, first in sliding_windows(lrs, refs, spec['T'])]
        monkeypatch.undo()
        if chain:   # per propagation step: trunk (2 x 3 blocks) + feat_decoder (17) + feat_decoder2 (9); res1/res2 per frame; BWFW per window
            assert calls['n'] > 20 and calls['layers'] > 300, calls
        else:
            assert calls['n'] == 0
    for a, b in zip(res[True], res[False]):
        assert torch.equal(a, b)


def test_pack_cache_roundtrip(tmp_path, monkeypatch):
    """hash-keyed on-disk cache of packed conv weights (packing.py): a hit returns the identical operand image,
    different parameters or packing arguments miss, a corrupt entry is rewritten."""
    import os
    from refvsr_b200 import packing
    g = torch.Generator().manual_seed(3)
    w = torch.randn((48, 48, 3, 3), generator=g)
    b = torch.randn((48,), generator=g)
    ref = packing.pack_conv_uncached('x', w, b, [(48, 48)], 1, 1, torch.bfloat16, 'cpu', True)
    monkeypatch.setenv('REFVSR_PACK_CACHE', str(tmp_path))
    a1 = packing.pack_conv('x', w, b, [(48, 48)], 1, 1, torch.bfloat16, 'cpu', True)
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].endswith('.pt')
    calls = []
    orig = packing.pack_conv_uncached
    monkeypatch.setattr(packing, 'pack_conv_uncached', lambda *a, **k: (calls.append(1), orig(*a, **k))[1])
    a2 = packing.pack_conv('y', w, b, [(48, 48)], 1, 1, torch.bfloat16, 'cpu', True)
    assert not calls, 'second pack of the same parameters must come from the cache'
    for a in (a1, a2):
        assert torch.equal(a.wpack.view(torch.uint8), ref.wpack.view(torch.uint8)) and torch.equal(a.bias, ref.bias)
        assert (a.impl, a.nb, a.layout, a.k_real) == (ref.impl, ref.nb, ref.layout, ref.k_real)
    assert a2.name == 'y'
    packing.pack_conv('x', w + 1e-3, b, [(48, 48)], 1, 1, torch.bfloat16, 'cpu', True)          # other weights
    packing.pack_conv('x', w, b, [(48, 48)], 1, 1, torch.float16, 'cpu', True)                  # other dtype
    packing.pack_conv('x', w, b, [(48, 48)], 1, 1, torch.bfloat16, 'cpu', False)                # SIMT layout
    assert len(calls) == 3 and len(os.listdir(tmp_path)) == 4
    with open(os.path.join(tmp_path, files[0]), 'wb') as f:                                       # corrupt entry
        f.write(b'garbage')
    a3 = packing.pack_conv('x', w, b, [(48, 48)], 1, 1, torch.bfloat16, 'cpu', True)
    assert len(calls) == 4 and torch.equal(a3.wpack.view(torch.uint8), ref.wpack.view(torch.uint8))


@pytest.mark.parametrize('name', ['small_t3_32x48', 'small_t7_24x32'])
def test_push_frame_equals_sliding_windows(name):
    """Network.push_frame (one entering frame per call, the engine's ring holds the rest) must give exactly what forward() gives
    for the corresponding sliding window - across steady windows, ring wrap-around and forced resets (reset_branch)."""
    from oracle.oracle_ops import OracleOps
    from refvsr_b200.synth import sliding_windows
    from util import build_case
    spec, cfg, net, lrs, refs, golden = build_case(name, 'cpu', ops=OracleOps(), b200_precision='fp32')
    wins = list(sliding_windows(lrs, refs, spec['T']))
    with torch.no_grad():
        want = [net(wl, wr, first, False, False)['result'][0].clone() for k, wl, wr, first in wins]
    spec, cfg, net2, lrs, refs, golden = build_case(name, 'cpu', ops=OracleOps(), b200_precision='fp32')
    with pytest.raises(RuntimeError, match='previous forward'):
        net2.push_frame(wins[0][1][0, -1], wins[0][2][0, -1])
    got = []
    with torch.no_grad():
        for k, wl, wr, first in wins:
            if k == 0:
                got.append(net2(wl, wr, first, False, False)['result'][0].clone())
            else:
                assert not first
                got.append(net2.push_frame(wl[0, -1], wr[0, -1]).clone())
    assert len(want) >= 3
    for k, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a, b), f'window {k}: push_frame differs from forward()'
    assert net2.Network.frame_itr_num == net.Network.frame_itr_num
    with pytest.raises(ValueError, match='shapes'):
        net2.push_frame(wins[0][1][0, -1, :, :-2], wins[0][2][0, -1])
