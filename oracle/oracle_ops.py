"""ORACLE - test infrastructure, NOT product code.

`OracleOps` implements the operator interface of refvsr_b200.lib.CudaOps on the CPU with the primitives
of oracle/refvsr_oracle.py.  Two uses, both in tests/:
  * per-kernel yardstick for the `-m gpu` tests (same arguments -> compare outputs);
  * a test double that lets the product's orchestration (refvsr_b200/network.py) run end to end on a
    machine without a GPU, so the schedule itself is checked against the golden vectors.
It is never importable from refvsr_b200/ and the product never falls back to it.
"""
import types

import torch
import torch.nn.functional as F

from . import refvsr_oracle as O


def _nchw(x):                       # (H,W,C) -> (1,C,H,W) fp32
    return x.detach().float().permute(2, 0, 1).unsqueeze(0)


def _nhwc(y):                       # (1,C,H,W) -> (H,W,C)
    return y[0].permute(1, 2, 0)


def _act(v, code):
    if code == 1:
        return torch.relu(v)
    if code == 2:
        return O.lrelu(v, 0.1)
    if code == 3:
        return O.lrelu(v, 0.2)
    if code == 4:
        return v.clamp(-3, 3)
    return v


class OracleOps:
    name = 'oracle'

    def launch_count(self):
        return 0

    # ---- conv ----
    def pack_conv(self, name, weight, bias, srcs, stride, pad, act_dtype, device, prefer_tc, bias_add=0.0):
        return types.SimpleNamespace(name=name, weight=weight.detach().float().cpu(),
                                     bias=bias.detach().float().cpu() + bias_add, srcs=list(srcs), stride=stride,
                                     pad=pad, cout=weight.shape[0], alloc0=srcs[0][1],
                                     alloc1=srcs[1][1] if len(srcs) > 1 else 0)

    def conv2d(self, layer, src0, src1, out, gate=None, res=None, act_pre=0, act_post=0, pixel_shuffle=False):
        xs = [_nchw(src0)[:, :layer.srcs[0][0]]]
        assert src0.shape[2] == layer.alloc0
        if src1 is not None:
            assert src1.shape[2] == layer.alloc1
            xs.append(_nchw(src1)[:, :layer.srcs[1][0]])
        y = F.conv2d(torch.cat(xs, 1), layer.weight, layer.bias, layer.stride, layer.pad)
        y = _act(y, act_pre)
        if gate is not None:
            y = y * _nchw(gate)[:, :layer.cout]
        if res is not None:
            y = y + _nchw(res)[:, :layer.cout]
        y = _act(y, act_post)
        if pixel_shuffle:
            y = O.pixel_shuffle2(y)
        o = _nhwc(y)
        assert out.shape[:2] == o.shape[:2], (layer.name, out.shape, o.shape)
        out[..., :o.shape[2]] = o.to(out.dtype)

    def pack_resblock(self, name, w1, b1, w2, b2, alloc, act_dtype, device):
        return types.SimpleNamespace(name=name, cout=w1.shape[0], alloc=alloc, w1=w1.detach().float().cpu(), b1=b1.detach().float().cpu(),
                                     w2=w2.detach().float().cpu(), b2=b2.detach().float().cpu())

    def resblock(self, rb, src, out, act_mid, act_post=0):
        x = _nchw(src)[:, :rb.cout]
        t = _act(F.conv2d(x, rb.w1, rb.b1, 1, 1), act_mid)
        t = t.to(src.dtype).float()                      # the fused kernel stores the intermediate in the activation dtype
        y = _act(x + F.conv2d(t, rb.w2, rb.b2, 1, 1), act_post)
        out[..., :rb.cout] = _nhwc(y).to(out.dtype)

    def space_to_depth2(self, src, out):
        H, W, C = src.shape
        z = src.view(H // 2, 2, W // 2, 2, C).permute(0, 2, 1, 3, 4).reshape(H // 2, W // 2, 4 * C)
        out.copy_(z)

    # ---- image / pyramid prep ----
    def prep_image(self, src, out, mat12=None, pool2=False):
        x = src.detach().float().unsqueeze(0)
        if mat12 is not None:
            m = torch.tensor(mat12, dtype=torch.float32).view(3, 4)
            x = F.conv2d(x, m[:, :3].reshape(3, 3, 1, 1).contiguous(), m[:, 3].contiguous())
        if pool2:
            x = O.avgpool2(x)
        out.zero_()
        out[..., :3] = _nhwc(x).to(out.dtype)

    def spynet_resize_norm(self, src, out):
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        x = (O.bilinear_resize(src.float().unsqueeze(0), out.shape[0], out.shape[1]) - mean) / std
        out.copy_(_nhwc(x))

    def avgpool2(self, src, out):
        out.copy_(_nhwc(O.avgpool2(_nchw(src))))

    def maxpool2(self, src, out):
        out.copy_(_nhwc(O.maxpool2(_nchw(src).float())).to(out.dtype))

    def resize_planes(self, src, out, inv_scale, mode='bicubic', clamp01=False):
        x = src.float().unsqueeze(0)
        if mode == 'bicubic':
            y = O.bicubic(x, 1.0 / inv_scale)
        else:
            assert inv_scale == 2.0
            y = O.nearest_down2(x)
        if clamp01:
            y = y.clamp(0, 1)
        assert tuple(y.shape[1:]) == tuple(out.shape), (y.shape, out.shape)
        out.copy_(y[0])

    def spynet_level_input(self, ref, supp, flow_prev, out8, flow_up):
        H, W = ref.shape[:2]
        fu = torch.zeros(1, 2, H, W) if flow_prev is None else O.bilinear_up2_align_corners(_nchw(flow_prev)) * 2.0
        x = torch.cat([_nchw(ref), O.flow_warp_border(_nchw(supp), fu), fu], 1)
        out8.copy_(_nhwc(x).to(out8.dtype))
        flow_up.copy_(_nhwc(fu))

    def flow_resize(self, flow, out):
        H, W = flow.shape[:2]
        h, w = out.shape[:2]
        f = O.bilinear_resize(_nchw(flow), h, w)
        f = torch.stack([f[:, 0] * (float(w) / float(W)), f[:, 1] * (float(h) / float(H))], 1)
        out.copy_(_nhwc(f))

    # ---- warp ----
    def warp3(self, feat, featUP, conf, flow, out_feat, out_featUP, out_conf):
        self.warp(feat, flow, out_feat)
        self.warp(conf, flow, out_conf)
        self.warp(featUP, flow, out_featUP, flow_up2=True)

    def warp(self, src, flow, out, flow_up2=False):
        x = _nchw(src if src.dim() == 3 else src.unsqueeze(-1))
        f = _nchw(flow)
        if flow_up2:
            f = O.bilinear_up2_align_corners(f) * 2.0
        y = _nhwc(O.warp(x, f))
        out.copy_((y if out.dim() == 3 else y[..., 0]).to(out.dtype))

    # ---- matching ----
    def patch_pack(self, feat, out, mode):
        p = F.normalize(O.extract_patches_3x3_reflect(_nchw(feat)), dim=1)[0].t() * 64.0    # (P, 144)
        K = p.shape[1]
        hi = p.half()
        lo = (p - hi.float()).half()
        out.zero_()
        if mode == 0:
            out[:, :K] = hi
        elif mode == 1:
            out[:, :K], out[:, K:2 * K], out[:, 2 * K:3 * K] = hi, lo, hi
        else:
            out[:, :K], out[:, K:2 * K], out[:, 2 * K:3 * K] = hi, hi, lo

    def match_argmax(self, A, B, conf, idx, impl=0):
        a, b = A.float(), B.float()
        for s in range(0, a.shape[0], 4096):
            S = b @ a[s:s + 4096].t()
            v, i = S.max(dim=0)
            conf.view(-1)[s:s + 4096] = v / 4096.0
            idx[s:s + 4096] = i.to(idx.dtype)

    # ---- reference alignment ----
    def gather_blocks(self, value, idx, hq, wq, ks, out):
        out.copy_(_nhwc(O.gather_blocks(_nchw(value), idx.long().view(1, -1), hq, wq, ks)).to(out.dtype))

    def aligned_sample(self, x, affine, ks, out):
        out.copy_(_nhwc(O.aligned_sample(_nchw(x), _nchw(affine), ks)).to(out.dtype))

    def bicubic_up2_image(self, src, out):
        out.zero_()
        out[..., :3] = _nhwc(O.bicubic(src.float().unsqueeze(0), 2)).to(out.dtype)

    # ---- confidence maps ----
    def conf_pair(self, a, b, out, up2=False):
        x = torch.stack([a.float(), b.float()], 0).unsqueeze(0)
        if up2:
            x = O.bicubic(x, 2).clamp(0, 1)
        out.zero_()
        out[..., :2] = _nhwc(x).to(out.dtype)

    def conf_max(self, a, b, out):
        out.copy_(torch.maximum(a, b))

    def pack_chain(self, name, weight, bias, alloc, act_dtype, device):
        c = weight.shape[0]
        return types.SimpleNamespace(name=name, nb=(alloc + 15) // 16 * 16, weight=weight.detach().float().cpu(),
                                     bias=bias.detach().float().cpu(), srcs=[(c, alloc)], stride=1, pad=1, cout=c,
                                     alloc0=alloc, alloc1=0)

    def conv_chain(self, bufs, layers, flags, max_ctas=0):
        """the chain's contract: exactly the layers, one after the other"""
        for pk, src, res, dst, a0, a1 in layers:
            assert dst != src
            self.conv2d(pk, bufs[src], None, bufs[dst], res=bufs[res] if res >= 0 else None, act_pre=a0, act_post=a1)

    def frames_differ(self, pairs, flag):
        flag.fill_(0 if all(torch.equal(a, b) for a, b in pairs) else 1)

    # ---- tail ----
    def reconstruct(self, x, lr, scale, clamp01, out):
        base = O.bicubic(lr.float().unsqueeze(0), scale).clamp(0, 1)
        y = _nchw(x)[:, :3] + base
        if clamp01:
            y = y.clamp(0, 1)
        out.copy_(y[0])
