"""Runs the dominant kernels in isolation at the BASELINE shapes (same code path as bench.py's roofline
leg) so that `ncu -k regex:<kernel>` can capture them without replaying a whole window."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cfg, net = bench.make_model(sys.argv[1] if len(sys.argv) > 1 else 'mfid', None, torch.device('cuda', 0))
r = bench.kernel_rooflines(net, bench.measured_peaks())
print(json.dumps(r, indent=1))
