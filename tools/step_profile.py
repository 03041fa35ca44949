"""Kernel-level breakdown of one CUDA-graph replay of the bench step (torch.profiler / CUPTI): per-kernel totals,
launch count, and the idle time between kernels (step wall minus the sum of kernel durations)."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from vmambair_b200.engine import InferenceEngine

dev = torch.device("cuda", 0)
eng = InferenceEngine(bench.build_net(os.environ.get("NET", "light")), 8, 64, 64, dtype=torch.bfloat16, device=dev)
for _ in range(5):
    eng.step_device()
torch.cuda.synchronize()
REPS = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        eng.step_device()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
evs.sort(key=lambda e: e.time_range.start)
tot = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    n = e.name
    for pre in ("void ", "vmb::"):
        n = n.replace(pre, "")
    n = n.split("(")[0][:70]
    tot[n][0] += 1
    tot[n][1] += e.device_time_total
busy = sum(v[1] for v in tot.values())
span = (evs[-1].time_range.end - evs[0].time_range.start)
print(f"replays {REPS}: span {span / REPS / 1e3:.3f} ms per step, kernel time {busy / REPS / 1e3:.3f} ms, idle {(span - busy) / REPS / 1e3:.3f} ms, "
      f"{len(evs) // REPS} kernels per step")
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t / REPS:9.1f} us  {c // REPS:4d} x {t / c:7.2f} us  {t / busy * 100:5.1f} %  {n}")
