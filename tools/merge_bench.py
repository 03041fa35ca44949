"""merge + out_norm + gate + pool: single-kernel path (VMB_MERGE_FUSED=1) vs the two-kernel path, sustained (CUDA-graph replay)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmambair_b200 import ops
dt = torch.bfloat16
for B, C, H, W in ((8, 96, 64, 64), (8, 48, 64, 64), (4, 96, 64, 64), (8, 96, 32, 32), (2, 96, 128, 128)):
    L = H * W
    sets = [dict(ys=torch.randn(B, 4, C, L, device="cuda").to(dt), z=torch.randn(B, C, L, device="cuda").to(dt)) for _ in range(4)]
    lw, lb = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    row = dict(B=B, C=C, H=H, W=W)
    for fused in ("0", "1"):
        os.environ["VMB_MERGE_FUSED"] = fused
        ops.merge_norm_gate(sets[0]["ys"], sets[0]["z"], lw, lb, C, H, W, in_place_order=True); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(12):
                ops.merge_norm_gate(sets[i % 4]["ys"], sets[i % 4]["z"], lw, lb, C, H, W, in_place_order=True)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(20): g.replay()
        e.record(); torch.cuda.synchronize()
        row["fused_us" if fused == "1" else "two_kernel_us"] = round(s.elapsed_time(e) / 240 * 1e3, 2)
    print(json.dumps(row), flush=True)
