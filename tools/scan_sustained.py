"""Sustained (back-to-back) timing of scan fwd: no host sync between launches, clocks sampled."""
import sys, os, json, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmambair_b200 import ops
dev = "cuda"
for dn, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
    for B in (1, 8, 32):
        K, N, C, L = 4, 16, 96, 4096
        D = K * C
        u = torch.randn(B, D, L, device=dev).to(dt); delta = (0.5 * torch.rand(B, D, L, device=dev)).to(dt)
        A = -0.5 * torch.rand(D, N, device=dev); Bm = torch.randn(B, K, N, L, device=dev).to(dt); Cm = torch.randn(B, K, N, L, device=dev).to(dt)
        Dv = torch.randn(D, device=dev); bias = 0.5 * torch.rand(D, device=dev)
        fn = lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, Dv, bias, True, True)
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        reps = 20
        s.record()
        for _ in range(reps):
            g.replay()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / (reps * 20)
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        print(json.dumps(dict(dtype=dn, B=B, ms=round(ms, 4), us_per_img=round(ms * 1e3 / B, 2), clk_after=clk)), flush=True)
