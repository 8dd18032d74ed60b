"""Dump the fields of the reference's config modules that the hot path reads -> tests/golden/config_values.json.
Run in the build container (needs /root/reference):  python tests/golden/make_config_values.py"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'oracle', 'shims'), '/root/reference']
FIELDS = ['num_blocks', 'mid_channels', 'frame_num', 'frame_itr_num', 'is_amp', 'flag_HD_in', 'reset_branch',
          'matching_ksize', 'scale', 'network']
out = {}
for name in ['config_RefVSR_small_L1', 'config_RefVSR_small_MFID', 'config_RefVSR_L1', 'config_RefVSR_MFID',
             'config_RefVSR_small_MFID_8K', 'config_RefVSR_MFID_8K']:
    c = importlib.import_module('configs.' + name).get_config('p', 'm', name)
    out[name] = {f: c[f] for f in FIELDS}
with open(os.path.join(ROOT, 'tests', 'golden', 'config_values.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps(out)[:300])
