"""Per-stage device time of one fused OSS block (eager, CUDA events around every library call)."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmambair_b200 import archs, ops, fused
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
C = int(sys.argv[2]) if len(sys.argv) > 2 else 96
H = W = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dt = torch.bfloat16
torch.manual_seed(0)
blk = archs.MamberBlock(dim=C, num_heads=1, ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias").cuda()
x = torch.randn(B, C, H, W, device="cuda").to(dt)
for _ in range(3):
    fused.block_forward(blk, x)
torch.cuda.synchronize()
rec = []
ops.set_timing(rec)
for _ in range(5):
    fused.block_forward(blk, x)
torch.cuda.synchronize()
ops.set_timing(None)
per = len(rec) // 5
tot = collections.OrderedDict()
for i, r in enumerate(rec):
    key = f"{i % per:02d}:{r[0]}"
    tot[key] = tot.get(key, 0.0) + r[2].elapsed_time(r[3]) / 5
for k, v in tot.items():
    print(f"{k:20s} {v*1e3:9.1f} us")
print("sum", round(sum(tot.values()) * 1e3, 1), "us  B,C,H =", B, C, H)
