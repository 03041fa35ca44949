"""a few channel_branch_bwd launches (B = 4, C = 96, SISR variant) for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vmambair_b200.archs as archs
from vmambair_b200 import ops, fused_train
torch.manual_seed(0)
blk = archs.MamberBlock(dim=96, num_heads=1, ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias").cuda()
a = blk.attn
pooled = torch.randn(4, 96, device="cuda") * 64
c = fused_train.channel_gate(a, pooled.requires_grad_(), 4096)
for _ in range(3):
    c = fused_train.channel_gate(a, pooled, 4096)
    c.backward(torch.randn_like(c))
torch.cuda.synchronize()
