"""One selective-scan forward launch configuration, a few launches (for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmambair_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8); ap.add_argument("--C", type=int, default=96); ap.add_argument("--L", type=int, default=4096)
ap.add_argument("--dtype", default="bf16"); ap.add_argument("--n", type=int, default=4); ap.add_argument("--ckpt", action="store_true")
a = ap.parse_args()
dt = {"bf16": torch.bfloat16, "fp32": torch.float32}[a.dtype]
D, K, N = 4 * a.C, 4, 16
torch.manual_seed(0)
u = torch.randn(a.B, D, a.L, device="cuda").to(dt); delta = (0.5 * torch.rand(a.B, D, a.L, device="cuda")).to(dt)
A = -0.5 * torch.rand(D, N, device="cuda"); Bm = torch.randn(a.B, K, N, a.L, device="cuda").to(dt); Cm = torch.randn(a.B, K, N, a.L, device="cuda").to(dt)
Dv = torch.randn(D, device="cuda"); bias = 0.5 * torch.rand(D, device="cuda")
for _ in range(a.n):
    ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dv, bias, True, a.ckpt)
torch.cuda.synchronize()
