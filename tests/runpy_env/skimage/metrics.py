"""test stand-in for skimage.metrics.structural_similarity (evaluation/metrics.py:4,15-16): uniform 7x7 window SSIM with
sample covariance, mean over the valid interior, per channel - the defaults skimage applies to float images."""
import numpy as np
from scipy.ndimage import uniform_filter


def structural_similarity(im1, im2, *, win_size=7, data_range=1.0, multichannel=False, channel_axis=None, full=False, **kw):
    im1, im2 = np.asarray(im1, np.float64), np.asarray(im2, np.float64)
    if (multichannel or channel_axis is not None) and im1.ndim == 3:
        res = [structural_similarity(im1[..., c], im2[..., c], win_size=win_size, data_range=data_range, full=full)
               for c in range(im1.shape[-1])]
        if full:
            return float(np.mean([r[0] for r in res])), np.stack([r[1] for r in res], -1)
        return float(np.mean(res))
    K1, K2 = 0.01, 0.03
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    ux, uy = uniform_filter(im1, win_size), uniform_filter(im2, win_size)
    uxx, uyy, uxy = uniform_filter(im1 * im1, win_size), uniform_filter(im2 * im2, win_size), uniform_filter(im1 * im2, win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    m = float(S[pad:-pad, pad:-pad].mean())
    return (m, S) if full else m
