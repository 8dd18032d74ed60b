import os, subprocess, sys
os.environ['REFVSR_LIB'] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'refvsr_b200', 'librefvsr_b200_exp.so')
code = r'''
import sys, os
sys.path.insert(0, os.getcwd())
import refvsr_b200.lib as L
L._lib = None
L.load_library(os.path.join(os.getcwd(), 'refvsr_b200', 'librefvsr_b200_exp.so'))   # first load wins
import importlib.util
spec = importlib.util.spec_from_file_location('cs', 'tools/conv_sweep_lib.py'); cs = importlib.util.module_from_spec(spec); spec.loader.exec_module(cs)
cs.case(48, 48, 3)
cs.case(48, 48, 1)
cs.case(48, 48, 3, hw=(540, 960))
'''
s = open('tools/conv_sweep.py').read()
open('tools/conv_sweep_lib.py', 'w').write(s[:s.index('case(48, 48, 3)\n')])
for dbg, label in [(0, 'baseline'), (32, 'no bias LDS'), (64, 'no output stores'), (128, 'no residual loads'), (224, 'no bias / stores / residual'), (1, 'no MMA')]:
    env = dict(os.environ, REFVSR_CONV_DBG=str(dbg))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(f'==== dbg={dbg} ({label})')
    print('\n'.join(l for l in (r.stdout + r.stderr).splitlines() if 'cin' in l or 'Error' in l or 'error' in l))
