import torch.nn as nn


class _Registry:
    def register_module(self, *a, **k):
        def deco(cls):
            return cls
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return deco

    def get(self, name):
        return None


CONV_LAYERS = _Registry()
UPSAMPLE_LAYERS = _Registry()
PLUGIN_LAYERS = _Registry()
ACTIVATION_LAYERS = _Registry()
NORM_LAYERS = _Registry()


class ConvModule(nn.Module):
    """conv (bias=True when norm_cfg is None) + optional ReLU; state-dict key `.conv.`
    (what SPyNet.py:152-191 relies on)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'),
                 inplace=True, **kw):
        super().__init__()
        assert norm_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation,
                              groups, bias=True)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            assert act_cfg['type'] == 'ReLU'
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def build_activation_layer(cfg):
    t = cfg['type']
    kw = {k: v for k, v in cfg.items() if k != 'type'}
    return getattr(nn, t)(**kw)


def build_conv_layer(cfg, *args, **kwargs):
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=''):
    return 'bn', nn.BatchNorm2d(num_features)


def build_upsample_layer(cfg, *args, **kwargs):
    raise NotImplementedError
