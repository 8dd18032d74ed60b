"""Drop-in for models/SRNet.py (reference lines cited inline)."""
import collections
import importlib

import numpy as np
import torch
import torch.nn as nn


class SRNet(nn.Module):
    """models/SRNet.py:11-61.  `config.network` selects the architecture module exactly like the
    reference's importlib hook (SRNet.py:20-21); 'RefVSR' resolves to refvsr_b200.network.Network, any
    other name is looked up as `models.archs.<name>` so third-party archs keep working."""

    def __init__(self, config):
        super().__init__()
        self.rank = torch.distributed.get_rank() if getattr(config, 'dist', False) else -1
        self.config = config
        self.device = config.device
        if config.network == 'RefVSR':
            from .network import Network
            self.Network = Network(config)
        else:
            lib = importlib.import_module('models.archs.{}'.format(config.network))
            self.Network = lib.Network(config)
        self.data = collections.OrderedDict()

    def weights_init(self, m):                       # SRNet.py:24-38
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            torch.nn.init.xavier_uniform_(m.weight, gain=self.config.wi)
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)
        elif type(m) in (torch.nn.BatchNorm2d, torch.nn.InstanceNorm2d):
            if m.weight is not None:
                torch.nn.init.constant_(m.weight, 1)
                torch.nn.init.constant_(m.bias, 0)
        elif type(m) == torch.nn.Linear:
            torch.nn.init.normal_(m.weight, 0, self.config.win)
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)

    def init(self):                                   # SRNet.py:40-45
        if self.config.wi is not None and self.config.win is not None:
            self.Network.apply(self.weights_init)
            if self.Network.FlowNet is not None:
                self.Network.FlowNet.load_ckpt(pretrained='./ckpt/SPyNet.pytorch')

    def input_constructor(self, res):                 # SRNet.py:47-54 (ptflops hook)
        b, f, c, h, w = res[:]
        imgs = torch.FloatTensor(np.random.randn(b, f, c, h, w)).to(self.device)
        return {'x': imgs, 'ref': imgs}

    def forward(self, x, ref, is_first_frame=True, is_log=False, is_train=False):   # SRNet.py:57-61
        return self.Network.forward(x, ref, is_first_frame, is_log=is_log, is_train=is_train)

    def push_frame(self, lr, ref):
        """streaming twin of forward() for callers that keep their own decoded-frame ring: after the clip's first window, hand
        over only the entering (LR, Ref) frame -> (3, 4h, 4w); see Network.push_frame (no counterpart in the reference)"""
        return self.Network.push_frame(lr, ref)
