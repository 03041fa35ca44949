import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmambair_b200 import ops
dt = torch.bfloat16
B, C, H, W = 8, 96, 64, 64
L = H * W
ys = torch.randn(B, 4, C, L, device="cuda").to(dt); z = torch.randn(B, C, L, device="cuda").to(dt)
lw, lb = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
for _ in range(3):
    ops.merge_norm_gate(ys, z, lw, lb, C, H, W, in_place_order=True)
torch.cuda.synchronize()
