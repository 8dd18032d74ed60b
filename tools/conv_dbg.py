import os, subprocess, sys
code = r'''
import sys, os
sys.path.insert(0, os.getcwd())
import importlib.util
spec = importlib.util.spec_from_file_location('cs', 'tools/conv_sweep_lib.py'); cs = importlib.util.module_from_spec(spec); spec.loader.exec_module(cs)
cs.case(48, 48, 3)
cs.case(48, 48, 1)
cs.case(48, 48, 3, hw=(540, 960))
'''
s = open('tools/conv_sweep.py').read()
open('tools/conv_sweep_lib.py', 'w').write(s[:s.index('case(48, 48, 3)\n')])
for dbg, label in [(0, 'baseline'), (16, 'MMAs rotate over 3 accumulators (timing only)'), (1, 'no MMA')]:
    env = dict(os.environ, REFVSR_CONV_DBG=str(dbg))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(f'==== dbg={dbg} ({label})')
    print('\n'.join(l for l in (r.stdout + r.stderr).splitlines() if 'cin' in l or 'Error' in l or 'error' in l))
