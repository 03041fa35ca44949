"""one vmb_conv3x3 launch per listed site (for ncu): python tools/conv_one.py up4_3 up2_1 ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vmambair_b200 import ops

SITES = {"patch_embed": (3, 48, 64, 0), "down1_2": (48, 24, 64, 1), "down2_3": (96, 48, 32, 1), "down3_4": (192, 96, 16, 1),
         "up4_3": (384, 768, 8, 2), "up3_2": (192, 384, 16, 2), "up2_1": (96, 192, 32, 2)}
B = int(os.environ.get("B", "8"))
for name in sys.argv[1:]:
    cin, cout, hw, mode = SITES[name]
    x = torch.randn(B, cin, hw, hw, device="cuda", dtype=torch.bfloat16)
    wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 30, torch.bfloat16)
    for _ in range(2):
        ops.conv3x3(x, wp, None, cout, mode)
    torch.cuda.synchronize()
