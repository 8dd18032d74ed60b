"""bf16 model with fp32-grade ('split') vs single-pass fp16 matching GEMM: PSNR vs the reference goldens + index flips."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from util import CASES, build_case, psnr
from refvsr_b200.synth import sliding_windows
for name in CASES:
    for mode in ('split', 'single'):
        spec, cfg, net, lrs, refs, golden = build_case(name, 'cuda', b200_precision='bf16', b200_match=mode)
        ps, flips = [], None
        for k, wl, wr, first in sliding_windows(lrs, refs, spec['T']):
            o = net(wl.cuda(), wr.cuda(), first, False, False)['result'][0].float().cpu()
            ps.append(psnr(o, torch.from_numpy(golden[f'result_{k}'])))
            if k == 0:
                N = net.Network
                flips = np.mean([(N._frame_slot(i % spec['T'], spec['h'], spec['w'])['idx'].cpu().numpy() != golden['idx_0'][i]).mean() for i in range(spec['T'])])
        print(f'{name} bf16 match={mode}: psnr min {min(ps):.1f} mean {np.mean(ps):.1f} dB, index flips vs reference {flips:.4f}')
