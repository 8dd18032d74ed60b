"""One conv_tc launch per geometry for `ncu --set full`: the trunk conv (3x3 48->48 +ReLU +residual) at 270x480 and
540x960, after warm-up launches that ncu skips (--launch-skip)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refvsr_b200 import packing
from refvsr_b200.lib import CudaOps, ACT_RELU
ops = CudaOps()
dt = torch.bfloat16
for (h, w) in ((270, 480), (540, 960)):
    C = 48
    wgt = torch.randn(C, C, 3, 3) * 0.05
    layer = packing.pack_conv('x', wgt, torch.zeros(C), [(C, C)], 1, 1, dt, 'cuda', True)
    x = torch.randn((h, w, C), device='cuda').to(dt)
    r = torch.randn((h, w, C), device='cuda').to(dt)
    y = torch.empty((h, w, C), device='cuda', dtype=dt)
    for i in range(3):          # 2 warm-up + 1 profiled per geometry  (ncu: --launch-skip 2 --launch-count 1, then 5 / 1)
        ops.conv2d(layer, x, None, y, res=r, act_pre=ACT_RELU)
    torch.cuda.synchronize()
print('done')
