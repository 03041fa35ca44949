import sys, torch
sys.path.insert(0, ".")
import bench
import torch.nn.functional as F
torch.manual_seed(0)
net = bench.build_net("full").cuda().train()
lq = torch.rand(4, 3, 64, 64, device="cuda"); gt = torch.rand(4, 3, 256, 256, device="cuda")
def fb():
    for p in net.parameters(): p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(lq)
    F.l1_loss(out.float(), gt).backward()
for _ in range(2): fb()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fb(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
