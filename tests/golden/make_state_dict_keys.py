"""Dump {config: {state_dict key: shape}} of the UNMODIFIED reference's SRNet (needs /root/reference).
    python tests/golden/make_state_dict_keys.py  ->  tests/golden/state_dict_keys.json"""
import importlib
import json
import os
import sys

from make_golden import ROOT, load_reference

if __name__ == '__main__':
    SRNet = load_reference()
    out = {}
    for name in ('config_RefVSR_small_L1', 'config_RefVSR_small_MFID', 'config_RefVSR_L1', 'config_RefVSR_MFID',
                 'config_RefVSR_small_MFID_8K', 'config_RefVSR_MFID_8K'):
        cfg = importlib.import_module('configs.' + name).get_config('p', 'm', name)
        cfg.cuda, cfg.device, cfg.dist = False, 'cpu', False
        out[name] = {k: list(v.shape) for k, v in SRNet(cfg).state_dict().items()}
    json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_keys.json'), 'w'))
    print({k: len(v) for k, v in out.items()})
