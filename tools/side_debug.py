"""run-to-run reproducibility of the training step's gradients, single stream vs weight gradients on the side stream
(debugging aid for fused_train._Side)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import vmambair_b200.archs as archs
from vmambair_b200 import fused_train
from vmambair_b200.optim import FlatAdam
torch.manual_seed(11)
archs.set_train_path("fused")
net = archs.MambaSISR6(dim=16, num_blocks=[2, 1, 1, 1], num_refinement_blocks=2).cuda().train()
opt = FlatAdam(net.parameters(), lr=1e-4)
lq, gt = torch.rand(2, 3, 32, 32, device="cuda"), torch.rand(2, 3, 128, 128, device="cuda")
names, offs = [], []
for n, p in net.named_parameters():
    names.append(n); offs.append(p.numel())
def run():
    opt.flat_grad.zero_()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=os.environ.get("AMP", "1") == "1"):
        out = net(lq)
    loss = F.l1_loss(out.float(), gt)
    loss.backward()
    torch.cuda.synchronize()
    return opt.flat_grad.clone(), float(loss), out.detach().float().clone()
def dist(a, b): return float((a - b).norm() / b.norm())
def worst_param(a, b):
    o, best = 0, (0.0, "")
    for n, k in zip(names, offs):
        d = float((a[o:o + k] - b[o:o + k]).norm() / b[o:o + k].norm().clamp_min(1e-12))
        if d > best[0]: best = (d, n)
        o += k
    return best
fused_train._Side.on = False
runs = [run() for _ in range(4)]
for i in range(1, 4):
    print(f"single stream run {i} vs run 0: grads {dist(runs[i][0], runs[0][0]):.2e}  output {dist(runs[i][2], runs[0][2]):.2e}  loss {runs[i][1]:.7f} vs {runs[0][1]:.7f}  worst param {worst_param(runs[i][0], runs[0][0])}", flush=True)
print(f"run 2 vs run 1: {dist(runs[2][0], runs[1][0]):.2e}")
fused_train._Side.on = True
for mask in (0, 255):
    fused_train._Side.mask = mask
    for j in range(3):
        g = run()
        print(f"side on, mask {mask:3d}, run {j} vs single run 1: grads {dist(g[0], runs[1][0]):.2e}  output {dist(g[2], runs[1][2]):.2e}  worst param {worst_param(g[0], runs[1][0])}", flush=True)
