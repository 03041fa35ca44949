"""A few selective-scan backward launches at (B, 4*96, 4096) bf16 (for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmambair_b200 import ops
ap = argparse.ArgumentParser(); ap.add_argument("--B", type=int, default=4); ap.add_argument("--n", type=int, default=3)
a = ap.parse_args()
dt, D, K, N, L = torch.bfloat16, 384, 4, 16, 4096
torch.manual_seed(0)
u = torch.randn(a.B, D, L, device="cuda").to(dt); delta = (0.5 * torch.rand(a.B, D, L, device="cuda")).to(dt)
A = -0.5 * torch.rand(D, N, device="cuda"); Bm = torch.randn(a.B, K, N, L, device="cuda").to(dt); Cm = torch.randn(a.B, K, N, L, device="cuda").to(dt)
Dv = torch.randn(D, device="cuda"); bias = 0.5 * torch.rand(D, device="cuda"); dout = torch.randn_like(u)
out, ck = ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dv, bias, True, True)
for _ in range(a.n):
    ops.selective_scan_bwd(u, delta, A, Bm, Cm, Dv, bias, dout, ck, True)
torch.cuda.synchronize()
