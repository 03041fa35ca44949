"""Kernel-level breakdown of one eager training forward+backward (full net, 4 x 3x64x64, bf16 autocast) with torch.profiler / CUPTI."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
import bench
from vmambair_b200.optim import FlatAdam

torch.manual_seed(0)
net = bench.build_net(os.environ.get("NET", "full")).cuda().train()
opt = FlatAdam(net.parameters(), lr=2e-4, betas=(0.9, 0.99), ema_decay=0.999)  # as train_bench: flat .grad views, direct accumulation, side stream
lq = torch.rand(4, 3, 64, 64, device="cuda"); gt = torch.rand(4, 3, 256, 256, device="cuda")
def fb():
    opt.flat_grad.zero_()
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
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"kernel time {busy / 1e3:.3f} ms, {len(evs)} kernels, first kernel start -> last kernel end {(t1 - t0) / 1e3:.3f} ms (eager launch)")
by_stream = collections.defaultdict(float)
for e in evs:
    by_stream[getattr(e, "stream", None) if hasattr(e, "stream") else 0] += e.device_time_total
print("kernel time per stream:", {k: round(v / 1e3, 2) for k, v in by_stream.items()})
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t:9.1f} us  {c:5d} x {t / c:7.2f} us  {t / busy * 100:5.1f} %  {n}")
