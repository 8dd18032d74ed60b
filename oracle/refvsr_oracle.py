"""ORACLE - test infrastructure, NOT product code.

A CPU (torch, fp32) restatement of RefVSR's per-frame forward (models/SRNet.py + models/archs), written
from the reference's source with every resampling primitive spelled out as explicit index arithmetic
(no F.grid_sample / F.interpolate), so that it documents exactly what the CUDA kernels must compute.
Each function cites the reference file:line it follows (paths relative to the reference checkout).

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.npz, which were produced by
importing the UNMODIFIED reference (tests/golden/make_golden.py) - and, where /root/reference is
present, against the live reference.  tests/test_oracle_primitives.py additionally checks every
primitive against the ATen op the reference calls.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Nothing under refvsr_b200/ does.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# resampling primitives (NCHW, fp32)
# --------------------------------------------------------------------------------------------------


def _src_index_half_pixel(n_out, n_in, scale=None, device='cpu'):
    """aten area_pixel_compute_source_index, align_corners=False, bilinear (clamped at 0)."""
    s = (n_in / n_out) if scale is None else scale
    x = torch.arange(n_out, dtype=torch.float32, device=device)
    src = (torch.tensor(s, dtype=torch.float32) * (x + 0.5) - 0.5).clamp_min(0)
    i0 = src.floor().long().clamp_max(n_in - 1)
    i1 = (i0 + 1).clamp_max(n_in - 1)
    return i0, i1, src - i0.float()


def bilinear_resize(x, Ho, Wo):
    """F.interpolate(x, size=(Ho,Wo), mode='bilinear', align_corners=False)  (SPyNet.py:120-133)"""
    H, W = x.shape[-2:]
    y0, y1, ly = _src_index_half_pixel(Ho, H, device=x.device)
    x0, x1, lx = _src_index_half_pixel(Wo, W, device=x.device)
    ly = ly.view(-1, 1)
    top = x[..., y0, :][..., x0] * (1 - lx) + x[..., y0, :][..., x1] * lx
    bot = x[..., y1, :][..., x0] * (1 - lx) + x[..., y1, :][..., x1] * lx
    return top * (1 - ly) + bot * ly


def bilinear_up2_align_corners(x):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
    (SPyNet.py:88-92, RefVSR.py:220,254,259): src = X*(in-1)/(out-1)."""
    H, W = x.shape[-2:]

    def idx(n_in):
        n_out = 2 * n_in
        sc = torch.tensor((n_in - 1) / (n_out - 1) if n_out > 1 else 0.0, dtype=torch.float32)
        src = sc * torch.arange(n_out, dtype=torch.float32, device=x.device)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        return i0, i1, src - i0.float()

    y0, y1, ly = idx(H)
    x0, x1, lx = idx(W)
    ly = ly.view(-1, 1)
    top = x[..., y0, :][..., x0] * (1 - lx) + x[..., y0, :][..., x1] * lx
    bot = x[..., y1, :][..., x0] * (1 - lx) + x[..., y1, :][..., x1] * lx
    return top * (1 - ly) + bot * ly


def _cubic_weights(t):
    A = -0.75
    w0 = ((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A
    w1 = ((A + 2) * t - (A + 3)) * t * t + 1
    w2 = ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1
    w3 = ((A * (2 - t) - 5 * A) * (2 - t) + 8 * A) * (2 - t) - 4 * A
    return torch.stack([w0, w1, w2, w3], -1)


def bicubic(x, scale):
    """F.interpolate(x, scale_factor=scale, mode='bicubic', align_corners=False) - Keys A=-0.75,
    unclamped source coordinate, clamped taps (RefVSR.py:105-106,125,140-141,288; alignment.py:41)."""
    H, W = x.shape[-2:]
    Ho, Wo = int(math.floor(H * scale)), int(math.floor(W * scale))
    inv = torch.tensor(1.0 / scale, dtype=torch.float32)

    def taps(n_out, n_in):
        src = inv * (torch.arange(n_out, dtype=torch.float32, device=x.device) + 0.5) - 0.5
        f = src.floor()
        w = _cubic_weights(src - f)                                   # (n_out, 4)
        idx = (f.long().view(-1, 1) + torch.arange(-1, 3, device=x.device).view(1, 4)).clamp(0, n_in - 1)
        return idx, w

    iy, wy = taps(Ho, H)
    ix, wx = taps(Wo, W)
    rows = x[..., ix]                                                  # (..., H, Wo, 4)
    rows = (rows * wx).sum(-1)                                         # x-pass first (aten order)
    cols = rows[..., iy, :]                                            # (..., Ho, 4, Wo)
    return (cols * wy.view(Ho, 4, 1)).sum(-2)


def nearest_down2(x):
    """F.interpolate(x, scale_factor=0.5, mode='nearest') (attention.py:65-67, flag_HD_in): src = floor(dst * 2)"""
    H, W = x.shape[-2:]
    return x[..., 0:2 * (H // 2):2, 0:2 * (W // 2):2]


def maxpool2(x):
    """nn.MaxPool2d(2, 2), torchvision vgg19.features[4] (attention.py:31-40 with vgg_range = 7)"""
    H, W = x.shape[-2:]
    x = x[..., :2 * (H // 2), :2 * (W // 2)]
    return torch.maximum(torch.maximum(x[..., 0::2, 0::2], x[..., 0::2, 1::2]),
                         torch.maximum(x[..., 1::2, 0::2], x[..., 1::2, 1::2]))


def avgpool2(x):
    """F.avg_pool2d(x, 2, 2) (SPyNet.py:66-78; attention.py:51)"""
    H, W = x.shape[-2:]
    x = x[..., : H // 2 * 2, : W // 2 * 2]
    return (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2]) * 0.25


def _gather2d(x, iy, ix, valid=None):
    """x (n,c,H,W); iy, ix (n,Ho,Wo) long (already in range) -> (n,c,Ho,Wo)"""
    n, c, H, W = x.shape
    flat = (iy * W + ix).view(n, 1, -1).expand(n, c, -1)
    out = x.reshape(n, c, H * W).gather(2, flat).view(n, c, *iy.shape[1:])
    if valid is not None:
        out = out * valid.unsqueeze(1).to(out.dtype)
    return out


def warp(x, flow):
    """models/utils.py:34-43: grid_sample(bilinear, zeros, align_corners=False) on the grid
    linspace(-1,1,Wout) + flow/((Win-1)/2).  x (n,c,Hi,Wi), flow (n,>=2,Ho,Wo)."""
    n, c, Hi, Wi = x.shape
    Ho, Wo = flow.shape[-2:]
    gx = torch.linspace(-1.0, 1.0, Wo, device=x.device).view(1, 1, Wo) + flow[:, 0] / ((Wi - 1.0) / 2.0)
    gy = torch.linspace(-1.0, 1.0, Ho, device=x.device).view(1, Ho, 1) + flow[:, 1] / ((Hi - 1.0) / 2.0)
    px = ((gx + 1) * Wi - 1) / 2
    py = ((gy + 1) * Hi - 1) / 2
    x0, y0 = px.floor(), py.floor()
    out = 0
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):                    # nw, ne, sw, se
        xi, yi = x0 + dx, y0 + dy
        wgt = (1 - (px - xi).abs()) * (1 - (py - yi).abs())
        valid = (xi >= 0) & (xi <= Wi - 1) & (yi >= 0) & (yi <= Hi - 1)
        v = _gather2d(x, yi.clamp(0, Hi - 1).long(), xi.clamp(0, Wi - 1).long(), valid)
        out = out + v * wgt.unsqueeze(1)
    return out


def flow_warp_border(x, flow):
    """mmedit/models/common/flow_warp.py:6-47 with padding_mode='border', align_corners=True
    (SPyNet.py:98-101).  flow (n,2,H,W): px = clamp(X+fx, 0, W-1)."""
    n, c, H, W = x.shape
    xs = torch.arange(W, dtype=torch.float32, device=x.device).view(1, 1, W)
    ys = torch.arange(H, dtype=torch.float32, device=x.device).view(1, H, 1)
    gx = 2.0 * (xs + flow[:, 0]) / max(W - 1, 1) - 1.0
    gy = 2.0 * (ys + flow[:, 1]) / max(H - 1, 1) - 1.0
    px = ((gx + 1) / 2 * (W - 1)).clamp(0, W - 1)
    py = ((gy + 1) / 2 * (H - 1)).clamp(0, H - 1)
    x0, y0 = px.floor(), py.floor()
    out = 0
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
        xi, yi = x0 + dx, y0 + dy
        wgt = (1 - (px - xi).abs()) * (1 - (py - yi).abs())
        valid = (xi <= W - 1) & (yi <= H - 1)
        v = _gather2d(x, yi.clamp(0, H - 1).long(), xi.clamp(0, W - 1).long(), valid)
        out = out + v * wgt.unsqueeze(1)
    return out


def extract_patches_3x3_reflect(x):
    """RefVSR_/utils.py:10-57 for ksize 3, stride 1: ReflectionPad2d(1) + Unfold -> (n, c*9, H*W)"""
    return F.unfold(F.pad(x, (1, 1, 1, 1), mode='reflect'), kernel_size=3)


def match_argmax(lr_f, ref_f, chunk=4096):
    """attention.py:69-91: L2-normalised 3x3 patches, S = ref_p @ lr_p, column max/argmax.
    Returns conf (n,1,h,w) fp32 and idx (n, h*w) int64 (lowest index wins ties, as torch.max on CPU)."""
    n, _, h, w = lr_f.shape
    lr_p = F.normalize(extract_patches_3x3_reflect(lr_f), dim=1)             # (n, 144, P)
    ref_p = F.normalize(extract_patches_3x3_reflect(ref_f).permute(0, 2, 1), dim=2)   # (n, R, 144)
    conf = torch.empty(n, h * w)
    idx = torch.empty(n, h * w, dtype=torch.long)
    for s in range(0, h * w, chunk):
        S = torch.bmm(ref_p, lr_p[:, :, s:s + chunk])
        v, i = S.max(dim=1)
        conf[:, s:s + chunk] = v
        idx[:, s:s + chunk] = i
    return conf.view(n, 1, h, w), idx


def gather_blocks(value, idx, hq, wq, ks):
    """AlignedAttention unfold -> gather -> fold (attention.py:142-144,152-154):
    out[:, :, ks*i+a, ks*j+b] = value[:, :, ks*ry+a, ks*rx+b], (ry,rx) = divmod(idx[i*wq+j], Wv/ks)."""
    n, c, Hv, Wv = value.shape
    wvk = Wv // ks
    idx = idx.view(n, hq, wq)
    ry, rx = idx // wvk, idx % wvk
    a = torch.arange(ks, device=value.device)
    iy = (ry.view(n, hq, 1, wq, 1) * ks + a.view(1, 1, ks, 1, 1)).expand(n, hq, ks, wq, ks).reshape(n, hq * ks, wq * ks)
    ix = (rx.view(n, hq, 1, wq, 1) * ks + a.view(1, 1, 1, 1, ks)).expand(n, hq, ks, wq, ks).reshape(n, hq * ks, wq * ks)
    return _gather2d(value, iy, ix)


def aligned_sample(x, affine, ks):
    """AlignedConv2d sampling (alignment.py:45-100,102-178).  x (n,c,ks*h,ks*w); affine (n,3,h,w) already
    clamp(p_conv(.)+1,-3,3).  ReflectionPad2d(1); corners AND p clamped; weights from clamped values."""
    n, c, H, W = x.shape
    h, w = affine.shape[-2:]
    xp = F.pad(x, (1, 1, 1, 1), mode='reflect')
    Hp, Wp = H + 2, W + 2
    s_x, s_y = affine[:, 0], affine[:, 1]                               # (n,h,w)
    th = (affine[:, 2] - 1.0) * 1.0472
    half = (ks - 1) // 2 + 0.5
    a = torch.arange(ks, dtype=torch.float32, device=x.device)
    u = (a.view(1, 1, 1, ks, 1) - half) * s_x.view(n, h, w, 1, 1)         # rows ("x" in the reference)
    v = (a.view(1, 1, 1, 1, ks) - half) * s_y.view(n, h, w, 1, 1)
    cs, sn = torch.cos(th).view(n, h, w, 1, 1), torch.sin(th).view(n, h, w, 1, 1)
    rr = u * cs - v * sn
    cc = u * sn + v * cs
    i = torch.arange(h, dtype=torch.float32, device=x.device).view(1, h, 1, 1, 1)
    j = torch.arange(w, dtype=torch.float32, device=x.device).view(1, 1, w, 1, 1)
    pr = rr + half + (1 + i * ks)
    pc = cc + half + (1 + j * ks)
    ltr, ltc = pr.floor(), pc.floor()
    rbr, rbc = ltr + 1, ltc + 1
    ltr, rbr, pr = ltr.clamp(0, Hp - 1), rbr.clamp(0, Hp - 1), pr.clamp(0, Hp - 1)
    ltc, rbc, pc = ltc.clamp(0, Wp - 1), rbc.clamp(0, Wp - 1), pc.clamp(0, Wp - 1)
    g_lt = (1 + (ltr - pr)) * (1 + (ltc - pc))
    g_rb = (1 - (rbr - pr)) * (1 - (rbc - pc))
    g_lb = (1 + (ltr - pr)) * (1 - (rbc - pc))
    g_rt = (1 - (rbr - pr)) * (1 + (ltc - pc))

    def tile(t):                                                       # (n,h,w,ks,ks) -> (n, ks*h, ks*w)
        return t.permute(0, 1, 3, 2, 4).reshape(n, h * ks, w * ks)

    out = 0
    for g, r_, c_ in ((g_lt, ltr, ltc), (g_rb, rbr, rbc), (g_lb, ltr, rbc), (g_rt, rbr, ltc)):
        out = out + _gather2d(xp, tile(r_).long(), tile(c_).long()) * tile(g).unsqueeze(1)
    return out


def pixel_shuffle2(x):
    """F.pixel_shuffle(x, 2): out[c, 2y+a, 2x+b] = in[4c+2a+b, y, x] (upsample.py:48-51)"""
    n, c4, h, w = x.shape
    c = c4 // 4
    return x.view(n, c, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, c, 2 * h, 2 * w)


def lrelu(x, s):
    return torch.where(x > 0, x, x * s)


# --------------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------------
class OracleRefVSR:
    """Functional restatement of models/archs/RefVSR.py::Network for scale 4, flag_HD_in False or True (the "8K"
    configs: matching on 1/4-resolution VGG[0:7] features, matching_ksize 8, both alignment levels with AlignedConv2d).
    `sd` is a state_dict with the reference's keys ('Network.' prefix optional)."""

    def __init__(self, config, sd):
        self.cfg = config
        self.C = config.mid_channels
        self.nb = config.num_blocks
        self.hd = bool(getattr(config, 'flag_HD_in', False))
        self.mk = int(getattr(config, 'matching_ksize', 2))       # config_RefVSR_*.py:31-39: 2, or 2 * scale with flag_HD_in
        self.sd = {(k[len('Network.'):] if k.startswith('Network.') else k): v.detach().float().cpu()
                   for k, v in sd.items()}
        self.frame_itr_num = 0
        self.max_frame_itr_num = config.reset_branch
        self.prev = None
        self.trace = None            # set to a dict to record intermediates
        self.compute_all_flows = False   # True: also compute the flows the reference computes but never uses

    # ---- layers ----
    def conv(self, name, x, stride=1, pad=None):
        w, b = self.sd[name + '.weight'], self.sd[name + '.bias']
        return F.conv2d(x, w, b, stride, w.shape[-1] // 2 if pad is None else pad)

    def resblock(self, p, x):                      # RefVSR_/common.py:33-39
        return x + self.conv(p + '.conv2', lrelu(self.conv(p + '.conv1', x), 0.2))

    def reslist(self, p, n, x):                    # RefVSR_/common.py:76-82
        x1 = x
        for i in range(n):
            x = self.resblock(f'{p}.RBs.{i}', x)
        return self.conv(p + '.conv_tail', x) + x1

    def basic2(self, p, x, stride0=1):             # nn.Sequential(BasicBlock, BasicBlock)
        x = lrelu(self.conv(p + '.0.0', x, stride0), 0.2)
        return lrelu(self.conv(p + '.1.0', x), 0.2)

    def prop_resblocks(self, p, x):                # RefVSR.py:327-360, sr_backbone_utils.py:85-97
        x = lrelu(self.conv(p + '.main.0', x), 0.1)
        for i in range(self.nb):
            x = x + self.conv(f'{p}.main.2.{i}.conv2', torch.relu(self.conv(f'{p}.main.2.{i}.conv1', x)))
        return x

    # ---- SPyNet (SPyNet.py:49-139) ----
    def spynet(self, ref, supp):
        h, w = ref.shape[-2:]
        w_up = w if w % 32 == 0 else 32 * (w // 32 + 1)
        h_up = h if h % 32 == 0 else 32 * (h // 32 + 1)
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        r = [(bilinear_resize(ref, h_up, w_up) - mean) / std]
        s = [(bilinear_resize(supp, h_up, w_up) - mean) / std]
        for _ in range(5):
            r.append(avgpool2(r[-1]))
            s.append(avgpool2(s[-1]))
        r, s = r[::-1], s[::-1]
        flow = ref.new_zeros(ref.shape[0], 2, h_up // 32, w_up // 32)
        for lv in range(6):
            flow_up = flow if lv == 0 else bilinear_up2_align_corners(flow) * 2.0
            x = torch.cat([r[lv], flow_warp_border(s[lv], flow_up), flow_up], 1)
            p = f'FlowNet.basic_module.{lv}.basic_module'
            for j in range(5):
                x = self.conv(f'{p}.{j}.conv', x)
                if j < 4:
                    x = torch.relu(x)
            flow = flow_up + x
        flow = bilinear_resize(flow, h, w)
        flow = torch.stack([flow[:, 0] * (float(w) / float(w_up)), flow[:, 1] * (float(h) / float(h_up))], 1)
        return flow                                 # RefVSR.py:184 resize to (h,w) is the identity

    # ---- FeatureMatching (attention.py:58-100) ----
    def match_features(self, x, pool):
        x = F.conv2d(x, self.sd['feature_match.sub_mean.weight'], self.sd['feature_match.sub_mean.bias'])
        if self.hd:
            x = nearest_down2(x)                                           # attention.py:65-67
        if pool:
            x = avgpool2(x)
        x = torch.relu(self.conv('feature_match.feature_extract.0', x))
        x = torch.relu(self.conv('feature_match.feature_extract.2', x))
        if self.hd:                                                        # vgg_range = 7: maxpool, conv 64->128, relu
            x = torch.relu(self.conv('feature_match.feature_extract.5', maxpool2(x)))
            return lrelu(self.conv('feature_match.feature_extract.map128.0', x), 0.2)
        return lrelu(self.conv('feature_match.feature_extract.map64.0', x), 0.2)

    def feature_match(self, lr, ref):
        conf, idx = match_argmax(self.match_features(lr, False), self.match_features(ref, True))
        h, hc = lr.shape[-2], conf.shape[-2]
        if h != hc:                                                        # attention.py:96-98
            conf = bicubic(conf, h / hc).clamp(0, 1)
        return conf, idx

    # ---- AlignedAttention / AlignedConv2d ----
    def align_conv1(self, x, aa='aa2'):            # alignment.py:19,42-43
        a = lrelu(self.conv(aa + '.align.conv1.0', x, pad=2), 0.2)
        return lrelu(self.resblock(aa + '.align.conv1.2', a), 0.2)

    def aligned_attention(self, aa, scale, align, lr, ref, idx, value):
        """attention.py:131-159: the output has twice the size of `lr`; blocks of scale x scale pixels of `value` (and,
        with align, of `ref` - indexed on ITS OWN block grid, as the reference does) are gathered by `idx`, then
        resampled by the per-block affine map of AlignedConv2d (kernel = stride = scale)."""
        n, _, h, w = lr.shape
        hq, wq = 2 * h // scale, 2 * w // scale
        warped = gather_blocks(value, idx, hq, wq, scale)
        if not align:
            return warped
        warped_ref = gather_blocks(ref, idx, hq, wq, scale)
        query = self.align_conv1(bicubic(lr, 2), aa)
        rf = self.align_conv1(warped_ref, aa)
        p = lrelu(self.conv(aa + '.align.p_conv.0', torch.cat([rf, query], 1), stride=scale, pad=2), 0.2)
        p = lrelu(self.resblock(aa + '.align.p_conv.2', p), 0.2)
        affine = (self.conv(aa + '.align.p_conv.4', p, pad=0) + 1.0).clamp(-3, 3)
        if self.trace is not None:
            self.trace.setdefault('affine', []).append(affine)
        return aligned_sample(warped, affine, scale)

    def aa2(self, lr, ref, idx, value):            # RefVSR.py:38: AlignedAttention(scale=matching_ksize, align=True)
        return self.aligned_attention('aa2', self.mk, True, lr, ref, idx, value)

    # ---- RAP (RefVSR.py:123-149) ----
    def rap(self, lr, ref, conf, conf_prop, idx, feat_prop, feat_prop_UP, ref_feat_down, ref_feat):
        n, _, h, w = lr.shape
        s1 = self.mk // 2                                                 # RefVSR.py:37: aa1 scale, align iff > 1
        lr_down = bicubic(lr, 0.5).clamp(0, 1)                            # RefVSR.py:125
        aligned = self.aligned_attention('aa1', s1, s1 > 1, lr_down, ref, idx, ref_feat_down)
        alpha = self.basic2('conf_fusion', torch.cat([conf_prop, conf], 1))
        feat_prop = feat_prop + alpha * self.basic2('feat_fusion', torch.cat([feat_prop, aligned], 1))
        feat_prop = self.reslist('feat_decoder', 8, feat_prop)
        aligned_up = self.aa2(lr, ref, idx, ref_feat)
        up = pixel_shuffle2(self.conv('upsample1.upsample_conv', feat_prop))
        feat_prop_UP = lrelu(self.conv('feat_fusion2_1.0.0', torch.cat([feat_prop_UP, up], 1)), 0.2)
        cpu_ = bicubic(conf_prop, 2).clamp(0, 1)
        cu_ = bicubic(conf, 2).clamp(0, 1)
        alpha2 = self.basic2('conf_fusion2', torch.cat([cpu_, cu_], 1))
        feat_prop_UP = feat_prop_UP + alpha2 * self.basic2('feat_fusion2', torch.cat([feat_prop_UP, aligned_up], 1))
        feat_prop_UP = self.reslist('feat_decoder2', 4, feat_prop_UP)
        conf_prop = torch.maximum(conf_prop, conf)
        return feat_prop, feat_prop_UP, conf_prop

    def ref_features(self, ref):                   # RefVSR.py:233-234
        ref_feat = self.reslist('res1', 4, self.basic2('ref_encoder1', ref))
        ref_feat_down = self.reslist('res2', 4, self.basic2('ref_encoder2', ref_feat, stride0=2))
        return ref_feat, ref_feat_down

    def compute_up(self, bw_up, fw_up, conf_bw, conf_fw, base):           # RefVSR.py:104-119
        cb = bicubic(conf_bw, 2).clamp(0, 1)
        cf = bicubic(conf_fw, 2).clamp(0, 1)
        cat = torch.cat([bw_up, fw_up], 1)
        out = self.conv('fusion_UP', cat, pad=0)
        alpha = self.basic2('conf_fusion_BWFW', torch.cat([cb, cf], 1))
        out = out + alpha * self.basic2('feat_fusion_BWFW', cat)
        out = self.reslist('feat_decoder_BWFW', 4, out)
        out = lrelu(pixel_shuffle2(self.conv('upsample2.upsample_conv', out)), 0.1)
        out = lrelu(self.conv('conv_hr', out), 0.1)
        return self.conv('conv_last', out) + base

    # ---- forward (RefVSR.py:151-325) ----
    @torch.no_grad()
    def forward(self, lrs, refs, is_first_frame, is_train=False):
        lrs, refs = lrs.float().cpu(), refs.float().cpu()
        n, t, c, h, w = lrs.shape
        mid = t // 2
        C = self.C
        if not is_train and self.max_frame_itr_num is not None and self.frame_itr_num == self.max_frame_itr_num:
            is_first_frame = True
        range_start = 0 if is_first_frame else (mid if not is_train else 0)
        gradio = bool(getattr(getattr(self.cfg, 'EVAL', None), 'is_gradio', False))
        # flows: only those that are consumed (the reference computes all 2(t-1), RefVSR.py:179-193)
        fw, bw = {}, {}
        for j in sorted(set(range(range_start, mid)) | ({mid} if mid < t - 1 else set())):
            fw[j] = lrs.new_zeros(n, 2, h, w) if gradio else self.spynet(lrs[:, j + 1], lrs[:, j])
        for j in range(mid, t - 1):
            bw[j] = lrs.new_zeros(n, 2, h, w) if gradio else self.spynet(lrs[:, j], lrs[:, j + 1])
        if self.compute_all_flows and not gradio:   # cost model of the reference as written (RefVSR.py:179-193)
            for j in range(0, t - 1):
                if j not in fw:
                    self.spynet(lrs[:, j + 1], lrs[:, j])
                if j not in bw:
                    self.spynet(lrs[:, j], lrs[:, j + 1])
        conf_maps, index_maps = {}, {}
        for i in range(range_start, t):                                   # RefVSR.py:196-204
            conf_maps[i], index_maps[i] = self.feature_match(lrs[:, i], refs[:, i])
        if self.trace is not None:
            self.trace.update(fw=fw, bw=bw, conf=conf_maps, idx=index_maps)

        # backward branch (RefVSR.py:211-238)
        feat_prop = lrs.new_zeros(n, C, h, w)
        feat_prop_UP = lrs.new_zeros(n, C, 2 * h, 2 * w)
        conf_prop = lrs.new_zeros(n, 1, h, w)
        for i in range(t - 1, mid - 1, -1):
            if i < t - 1:
                flow = bw[i]
                feat_prop = warp(feat_prop, flow)
                conf_prop = warp(conf_prop, flow)
                feat_prop_UP = warp(feat_prop_UP, bilinear_up2_align_corners(flow) * 2.0)
            feat_prop = self.prop_resblocks('backward_resblocks', torch.cat([lrs[:, i], feat_prop], 1))
            ref_feat, ref_feat_down = self.ref_features(refs[:, i])
            feat_prop, feat_prop_UP, conf_prop = self.rap(lrs[:, i], refs[:, i], conf_maps[i], conf_prop, index_maps[i],
                                                          feat_prop, feat_prop_UP, ref_feat_down, ref_feat)
        backward_feat_UP, conf_bw = feat_prop_UP, conf_prop

        # forward branch (RefVSR.py:241-283)
        if is_first_frame:
            feat_prop = torch.zeros_like(feat_prop)
            feat_prop_UP = torch.zeros_like(backward_feat_UP)
            conf_prop = torch.zeros_like(conf_bw)
            range_start = 0
        for i in range(range_start, mid + 1):
            if i > range_start:
                flow = fw[i - 1]
                feat_prop = warp(feat_prop, flow)
                feat_prop_UP = warp(feat_prop, bilinear_up2_align_corners(flow) * 2.0)     # quirk, RefVSR.py:254
                conf_prop = warp(conf_prop, flow)
            elif i == range_start and not is_first_frame:
                flow = self.prev['flow']
                feat_prop = warp(self.prev['feat'], flow)
                feat_prop_UP = warp(self.prev['featUP'], bilinear_up2_align_corners(flow) * 2.0)
                conf_prop = warp(self.prev['conf'], flow)
            feat_prop = self.prop_resblocks('forward_resblocks', torch.cat([lrs[:, i], feat_prop], 1))
            ref_feat, ref_feat_down = self.ref_features(refs[:, i])
            feat_prop, feat_prop_UP, conf_prop = self.rap(lrs[:, i], refs[:, i], conf_maps[i], conf_prop, index_maps[i],
                                                          feat_prop, feat_prop_UP, ref_feat_down, ref_feat)
            if (is_train and i == 0) or (not is_train and i == mid):
                self.prev = {'feat': feat_prop.clone(), 'featUP': feat_prop_UP.clone(), 'conf': conf_prop.clone(),
                             'flow': fw[i].clone() if i in fw else None}
        base = bicubic(lrs[:, mid], 4).clamp(0, 1)                        # RefVSR.py:288
        out = self.compute_up(backward_feat_UP, feat_prop_UP, conf_bw, conf_prop, base)
        if not is_train:
            if is_first_frame:
                self.frame_itr_num = 0
            self.frame_itr_num += 1
            out = out.clamp(0, 1)
        if self.trace is not None:
            self.trace.update(conf_bw=conf_bw, conf_fw=conf_prop, bw_UP=backward_feat_UP, fw_UP=feat_prop_UP)
        return out
