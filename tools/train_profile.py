"""Kernel-level breakdown of one eager training forward+backward (full net, 4 x 3x64x64, bf16 autocast) with torch.profiler / CUPTI."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
import bench

torch.manual_seed(0)
net = bench.build_net(os.environ.get("NET", "full")).cuda().train()
lq = torch.rand(4, 3, 64, 64, device="cuda"); gt = torch.rand(4, 3, 256, 256, device="cuda")
def fb():
    for p in net.parameters(): p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(lq)
    F.l1_loss(out.float(), gt).backward()
for _ in range(3): fb()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fb(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
tot = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    n = e.name
    for pre in ("void ", "vmb::"):
        n = n.replace(pre, "")
    n = n.split("(")[0][:80]
    tot[n][0] += 1
    tot[n][1] += e.device_time_total
busy = sum(v[1] for v in tot.values())
print(f"kernel time {busy / 1e3:.3f} ms, {len(evs)} kernels")
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t:9.1f} us  {c:5d} x {t / c:7.2f} us  {t / busy * 100:5.1f} %  {n}")
