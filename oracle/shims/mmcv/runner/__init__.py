import os


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None, **kw):
    """No-op when the file is absent (RefVSR.py:27 hard-codes ./ckpt/SPyNet.pytorch)."""
    if not os.path.isfile(filename):
        return None
    import torch
    sd = torch.load(filename, map_location='cpu')
    if 'state_dict' in sd:
        sd = sd['state_dict']
    model.load_state_dict(sd, strict=strict)
    return sd


def auto_fp16(*a, **k):
    def deco(f):
        return f
    return deco
