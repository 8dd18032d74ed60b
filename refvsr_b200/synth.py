"""Deterministic synthetic clips (no dataset offline).  CPU generator => identical bytes on
every box with the same torch build.  A textured HR scene translates by (+1.5, -0.75) LR pixels per frame;
LR = 4x area-downsample of the full field of view, Ref = the centre half field of view (2x zoom, like the
wide vs ultra-wide camera pair of RealMCVSR) at `ref_scale` x the LR size (1 = the reference's eval
setting RefVSR.py:155, 2 = BASELINE config #2)."""
import torch
import torch.nn.functional as F


def make_clip(num_frames, h, w, ref_scale=1, seed=1234, return_hr=False):
    g = torch.Generator(device='cpu').manual_seed(seed)
    H, W = 4 * h, 4 * w
    mx, my = 6 * num_frames + 8, 3 * num_frames + 8          # margins for the translation (HR pixels)
    ch, cw = H + 2 * my, W + 2 * mx
    scene = 0.0
    for s, amp in ((32, 0.5), (8, 0.3), (2, 0.2)):           # multi-octave texture
        lo = torch.rand(1, 3, ch // s + 3, cw // s + 3, generator=g)
        scene = scene + amp * F.interpolate(lo, size=(ch, cw), mode='bicubic', align_corners=False)
    scene = (scene + 0.04 * torch.randn(1, 3, ch, cw, generator=g)).clamp(0, 1)
    lrs, refs, hrs = [], [], []
    for k in range(num_frames):
        ox, oy = mx + 6 * k, my - 3 * k
        oy = max(0, min(oy, ch - H))
        hr = scene[:, :, oy:oy + H, ox:ox + W]
        lrs.append(F.avg_pool2d(hr, 4))
        hrs.append(hr)
        cy, cx = H // 4, W // 4                              # centre half-FoV crop (2h x 2w HR pixels... x2)
        crop = hr[:, :, cy:cy + H // 2, cx:cx + W // 2]      # (2h, 2w)
        refs.append(crop if ref_scale == 2 else F.avg_pool2d(crop, 2))
    if return_hr:
        return torch.cat(lrs, 0).contiguous(), torch.cat(refs, 0).contiguous(), torch.cat(hrs, 0).contiguous()
    return torch.cat(lrs, 0).contiguous(), torch.cat(refs, 0).contiguous()


def sliding_windows(lrs, refs, t):
    """Windows as data_loader/datasets.py:222-245 builds them: centre frame k, indices clamped to the clip."""
    n = lrs.shape[0]
    for k in range(n):
        idx = [min(max(k - t // 2 + j, 0), n - 1) for j in range(t)]
        yield k, lrs[idx].unsqueeze(0), refs[idx].unsqueeze(0), (k == 0)


def make_clip_range(start, num, h, w, ref_scale=1, seed=1234):
    """Frames [start, start+num) of an endless synthetic stream: same construction as make_clip but with a
    bounded, periodic camera path (triangle wave, period 32 frames), so any rank can render exactly the frames it
    owns (bench.py multi-GPU sharding) from a fixed-size seeded scene."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    H, W = 4 * h, 4 * w
    amp = 16
    mx, my = 6 * amp + 8, 3 * amp + 8
    ch, cw = H + 2 * my, W + 2 * mx
    scene = 0.0
    for s_, a_ in ((32, 0.5), (8, 0.3), (2, 0.2)):
        lo = torch.rand(1, 3, ch // s_ + 3, cw // s_ + 3, generator=g)
        scene = scene + a_ * F.interpolate(lo, size=(ch, cw), mode='bicubic', align_corners=False)
    scene = (scene + 0.04 * torch.randn(1, 3, ch, cw, generator=g)).clamp(0, 1)
    lrs, refs = [], []
    for k in range(start, start + num):
        ph = k % (2 * amp)
        tri = ph if ph <= amp else 2 * amp - ph          # 0..amp..0
        ox, oy = mx + 6 * tri - 3 * amp, my - 3 * tri + 3 * amp // 2
        ox, oy = max(0, min(ox, cw - W)), max(0, min(oy, ch - H))
        hr = scene[:, :, oy:oy + H, ox:ox + W]
        lrs.append(F.avg_pool2d(hr, 4))
        crop = hr[:, :, H // 4:H // 4 + H // 2, W // 4:W // 4 + W // 2]
        refs.append(crop if ref_scale == 2 else F.avg_pool2d(crop, 2))
    return torch.cat(lrs, 0).contiguous(), torch.cat(refs, 0).contiguous()
