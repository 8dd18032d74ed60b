"""Parameter containers that reproduce the reference's `state_dict()` schema key for key.

Nothing in this file computes anything: these modules only *hold* nn.Conv2d parameters under the same
attribute paths as the reference so that `ckpt_manager.CKPT_Manager.load_ckpt` (ckpt_manager.py:22-60,
load_state_dict(strict=False)), `SRNet.weights_init` (models/SRNet.py:24-38) and DataParallel / DDP
wrapping keep working.  The forward pass lives in engine.py and only reads `.weight` / `.bias`.

Schema source: models/archs/RefVSR.py:15-101, SPyNet.py:142-191, RefVSR_/attention.py:14-56,
RefVSR_/alignment.py:11-37, RefVSR_/common.py:25-109, mmedit/models/common/sr_backbone_utils.py:42-83,
mmedit/models/common/upsample.py:8-38 (476 tensors for MFID, 428 for the small models).
"""
import math

import torch
import torch.nn as nn


class _Holder(nn.Module):
    """A module whose forward must never run: compute happens in the CUDA engine."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('refvsr_b200 parameter holders are not callable; use Network.forward')


def _conv(cin, cout, k, stride=1, pad=None):
    return nn.Conv2d(cin, cout, k, stride, k // 2 if pad is None else pad, bias=True)


class ConvModuleHolder(_Holder):          # mmcv ConvModule: key `.conv.`
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = _conv(cin, cout, k)


class SPyNetBasicModule(_Holder):          # SPyNet.py:142-191
    def __init__(self):
        super().__init__()
        chans = [(8, 32), (32, 64), (64, 32), (32, 16), (16, 2)]
        self.basic_module = nn.Sequential(*[ConvModuleHolder(a, b, 7) for a, b in chans])


class SPyNet(_Holder):                     # SPyNet.py:12-47
    def __init__(self):
        super().__init__()
        self.basic_module = nn.ModuleList([SPyNetBasicModule() for _ in range(6)])

    def load_ckpt(self, pretrained):       # SPyNet.py:45-47 (plain torch.load; mmcv not needed)
        import os
        if not os.path.isfile(pretrained):
            raise FileNotFoundError(pretrained)
        sd = torch.load(pretrained, map_location='cpu')
        sd = sd.get('state_dict', sd)
        self.load_state_dict(sd, strict=False)


class ResBlock(_Holder):                   # RefVSR_/common.py:25-39
    def __init__(self, c):
        super().__init__()
        self.conv1 = _conv(c, c, 3)
        self.conv2 = _conv(c, c, 3)


class ResList(_Holder):                    # RefVSR_/common.py:64-82
    def __init__(self, n, c):
        super().__init__()
        self.RBs = nn.ModuleList([ResBlock(c) for _ in range(n)])
        self.conv_tail = _conv(c, c, 3)


def BasicBlock(cin, cout, k, stride=1):    # RefVSR_/common.py:96-109 -> Sequential(conv, act): key `.0.`
    return nn.Sequential(_conv(cin, cout, k, stride), nn.LeakyReLU(0.2, inplace=True))


class MeanShift(nn.Conv2d):                # RefVSR_/common.py:84-94 (frozen 1x1 conv)
    def __init__(self, rgb_range, rgb_mean, rgb_std, sign=-1):
        super().__init__(3, 3, kernel_size=1)
        std = torch.tensor(rgb_std)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1) / std.view(3, 1, 1, 1)
        self.bias.data = sign * rgb_range * torch.tensor(rgb_mean) / std
        self.weight.requires_grad = False
        self.bias.requires_grad = False


class FeatureMatching(_Holder):            # attention.py:14-56
    def __init__(self, scale, flag_HD_in):
        super().__init__()
        self.vgg_range = (4 if scale == 4 else 7) if not flag_HD_in else 7
        fe = nn.Sequential()
        fe.add_module('0', _conv(3, 64, 3))
        fe.add_module('1', nn.ReLU(inplace=True))
        fe.add_module('2', _conv(64, 64, 3))
        fe.add_module('3', nn.ReLU(inplace=True))
        if self.vgg_range == 7:
            fe.add_module('4', nn.MaxPool2d(2, 2))
            fe.add_module('5', _conv(64, 128, 3))
            fe.add_module('6', nn.ReLU(inplace=True))
        width = 64 if self.vgg_range == 4 else 128
        fe.add_module(f'map{width}', BasicBlock(width, 16, 1))
        self.feature_extract = fe
        self.sub_mean = MeanShift(1, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))


class AlignedConv2d(_Holder):              # alignment.py:11-37 (p_conv registered before conv1)
    def __init__(self, stride):
        super().__init__()
        self.p_conv = nn.Sequential(_conv(64, 32, 5, stride, 2), nn.LeakyReLU(0.2, True), ResBlock(32),
                                    nn.LeakyReLU(0.2, True), _conv(32, 3, 1, 1, 0))
        self.conv1 = nn.Sequential(_conv(3, 32, 5, 1, 2), nn.LeakyReLU(0.2, True), ResBlock(32),
                                   nn.LeakyReLU(0.2, True))


class AlignedAttention(_Holder):           # attention.py:102-117
    def __init__(self, scale, align):
        super().__init__()
        self.scale = scale
        self.has_align = bool(align)
        if align:
            self.align = AlignedConv2d(stride=scale)


class ResidualBlockNoBN(_Holder):          # sr_backbone_utils.py:42-83
    def __init__(self, c):
        super().__init__()
        self.conv1 = _conv(c, c, 3)
        self.conv2 = _conv(c, c, 3)
        for m in (self.conv1, self.conv2):  # default_init_weights(m, 0.1)
            nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in', nonlinearity='relu')
            m.weight.data *= 0.1
            nn.init.constant_(m.bias, 0)


class ResidualBlocksWithInputConv(_Holder):  # RefVSR.py:327-360
    def __init__(self, cin, cout, num_blocks):
        super().__init__()
        self.main = nn.Sequential(_conv(cin, cout, 3), nn.LeakyReLU(0.1, inplace=True),
                                  nn.Sequential(*[ResidualBlockNoBN(cout) for _ in range(num_blocks)]))


class PixelShufflePack(_Holder):           # upsample.py:8-38
    def __init__(self, cin, cout, scale):
        super().__init__()
        self.upsample_conv = _conv(cin, cout * scale * scale, 3)
        nn.init.kaiming_normal_(self.upsample_conv.weight, a=0, mode='fan_in', nonlinearity='relu')
        nn.init.constant_(self.upsample_conv.bias, 0)


def build_parameter_tree(net, config):
    """Attach all sub-modules to `net` in the reference's registration order (RefVSR.py:26-94)."""
    C, nb = config.mid_channels, config.num_blocks
    net.FlowNet = SPyNet()
    for p in net.FlowNet.parameters():
        p.requires_grad = False
    net.feature_match = FeatureMatching(config.scale, config.flag_HD_in)
    ks = config.matching_ksize
    net.aa1 = AlignedAttention(ks // 2, ks // 2 > 1)
    net.aa2 = AlignedAttention(ks, True)
    net.ref_encoder1 = nn.Sequential(BasicBlock(3, C, 3), BasicBlock(C, C, 3))
    net.res1 = ResList(4, C)
    net.ref_encoder2 = nn.Sequential(BasicBlock(C, C, 3, 2), BasicBlock(C, C, 3))
    net.res2 = ResList(4, C)
    net.conf_fusion = nn.Sequential(BasicBlock(2, 16, 3), BasicBlock(16, C, 3))
    net.feat_fusion = nn.Sequential(BasicBlock(2 * C, C, 3), BasicBlock(C, C, 3))
    net.feat_decoder = ResList(8, C)
    net.conf_fusion2 = nn.Sequential(BasicBlock(2, 16, 3), BasicBlock(16, C, 3))
    net.feat_fusion2_1 = nn.Sequential(BasicBlock(2 * C, C, 3))
    net.feat_fusion2 = nn.Sequential(BasicBlock(2 * C, C, 3), BasicBlock(C, C, 3))
    net.feat_decoder2 = ResList(4, C)
    net.conf_fusion_BWFW = nn.Sequential(BasicBlock(2, 16, 3), BasicBlock(16, C, 3))
    net.feat_fusion_BWFW = nn.Sequential(BasicBlock(2 * C, C, 3), BasicBlock(C, C, 3))
    net.feat_decoder_BWFW = ResList(4, C)
    net.backward_resblocks = ResidualBlocksWithInputConv(C + 3, C, nb)
    net.forward_resblocks = ResidualBlocksWithInputConv(C + 3, C, nb)
    net.fusion_UP = nn.Conv2d(2 * C, C, 1, 1, 0, bias=True)
    net.upsample1 = PixelShufflePack(C, C, 2)
    if config.scale == 4:
        net.upsample2 = PixelShufflePack(C, C, 2)
    net.conv_hr = nn.Conv2d(C, C, 3, 1, 1)
    net.conv_last = nn.Conv2d(C, 3, 3, 1, 1)


@torch.no_grad()
def seeded_test_weights(module, seed=1234, gain=0.62):
    """Deterministic, platform-independent weights for tests / golden vectors / benchmarks (there are no
    checkpoints offline).  Variance-preserving uniform init per tensor from a CPU generator, in
    state_dict order; residual-branch second convs are damped like the reference's 0.1-scaled kaiming
    init (sr_backbone_utils.py:68-83) so 30-block chains stay O(1); SPyNet's and the affine
    regressor's last layers are boosted so that flows / affine maps are non-trivial under random
    weights.  The frozen MeanShift is left at its analytic values."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    sd = module.state_dict()
    for k, v in sd.items():
        if 'sub_mean' in k:
            continue
        if k.endswith('.bias'):
            v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.1)
            if k.endswith('p_conv.4.bias'):
                v.copy_(torch.tensor([0.15, -0.2, 0.3])[: v.numel()])
            continue
        fan_in = v[0].numel()
        bound = gain * math.sqrt(3.0 / fan_in) * math.sqrt(2.0)
        w = (torch.rand(v.shape, generator=g) * 2 - 1) * bound
        if '.conv2.weight' in k or 'conv_tail.weight' in k:
            w *= 0.3
        if 'FlowNet' in k and '.4.conv.weight' in k:
            w *= 0.5           # flows of about a pixel at the finest level instead of tens
        if k.endswith('p_conv.4.weight'):
            w *= 0.5
        if k.startswith('Network.conv_last') or k.startswith('conv_last'):
            w *= 0.15          # output = bicubic base + small residual, like a trained model
        v.copy_(w)
    return module
