"""Standalone config factories for the fields the hot path reads:
network, scale, flag_HD_in, num_blocks, mid_channels, matching_ksize, reset_branch, frame_num, device,
dist, is_amp, EVAL.is_gradio, save_sample.  Values copied from configs/config.py:8-118 and
configs/config_RefVSR_{small_,}{L1,MFID}{,_8K}.py; the reference's own `configs` package keeps working
unchanged with run.py/eval.py - this module exists so tests/bench need no easydict."""


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


_MODELS = {
    # name: (num_blocks, mid_channels, frame_num, is_amp, flag_HD_in, frame_itr_num, reset_branch)
    # (values printed from the reference's own config modules; reset_branch = frame_itr_num except MFID_8K: None)
    'config_RefVSR_small_L1': (24, 24, 13, True, False, 26, 26),        # config_RefVSR_small_L1.py
    'config_RefVSR_small_MFID': (24, 24, 7, True, False, 9, 9),         # config_RefVSR_small_MFID.py:20-47
    'config_RefVSR_L1': (30, 48, 13, False, False, 26, 26),             # config_RefVSR_L1.py
    'config_RefVSR_MFID': (30, 48, 7, False, False, 9, 9),              # config_RefVSR_MFID.py:21-47
    'config_RefVSR_small_MFID_8K': (24, 24, 3, True, True, 9, 9),       # config_RefVSR_small_MFID_8K.py
    'config_RefVSR_MFID_8K': (30, 48, 7, False, True, 9, None),         # config_RefVSR_MFID_8K.py:26-48
}


def get_config(name, device='cuda', **overrides):
    if not name.startswith('config_'):
        name = 'config_' + name
    nb, c, t, amp, hd, itr, reset = _MODELS[name]
    cfg = Config()
    cfg.config = name
    cfg.network = 'RefVSR'
    cfg.trainer = 'trainer'
    cfg.scale = 4
    cfg.flag_HD_in = hd
    cfg.matching_ksize = 2 * (4 if hd else 1)       # config_RefVSR_MFID.py:31-39
    cfg.num_blocks, cfg.mid_channels = nb, c
    cfg.frame_num, cfg.frame_itr_num = t, itr
    cfg.reset_branch = reset
    cfg.is_amp = amp
    cfg.dist = False
    cfg.cuda = device != 'cpu'
    cfg.device = device
    cfg.wi = cfg.win = None
    cfg.save_sample = False
    cfg.is_train = False
    cfg.EVAL = Config(is_gradio=False, is_replicate=False)
    for k, v in overrides.items():
        cfg[k] = v
    return cfg
