import os, sys, subprocess
# run the kx=1 / kx=2 single-tap probes under every constant base_offset
code = r'''
import sys, os
sys.path.insert(0, os.getcwd())
import torch
sys.argv=['x']
import importlib.util
spec = importlib.util.spec_from_file_location('tp', 'tools/tc_probe_lib.py'); tp = importlib.util.module_from_spec(spec); spec.loader.exec_module(tp)
tp.LAYOUT = 1
for (dy,dx) in ((0,0),(0,1),(-1,0),(1,1)):
    tp.probe(f'shift{dy}{dx}', 16, 32, [(48, 48)], 48, 3, tp.shift(dy, dx), xfun=tp.coords)
'''
for bo in range(8):
    env = dict(os.environ, REFVSR_BO_FORCE=str(bo))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print('==== base_offset', bo)
    print('\n'.join(l for l in (r.stdout + r.stderr).splitlines() if 'max err' in l or 'got' in l or 'exp' in l or 'Error' in l))
