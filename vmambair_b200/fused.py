"""Fused inference pipeline of one OSS block: 15 launches of this library's kernels instead of the ~85
un-fused PyTorch ops of the reference MamberBlock.forward (SRGAN/VmambaIR/archs/MambaSISR6_arch.py:486-515).

   norm1+in_conv(+SiLU z) -> dwconv3x3+SiLU -> transpose (the (W,H) copy of x) -> x_proj (+dt_proj folded: W_dt W_x is a
   CxC map, so delta, B, C of directions 0/2 come out of ONE GEMM on x and of 1/3 out of one GEMM on x^T) ->
   direction-aware selective scan (reversal inside the kernel, no gathered operands) -> 4-way sum in the reference's
   order + out_norm + gate + pooling -> channel-direction OSS (1 CTA/image) -> channel gate + out_conv + residual ->
   norm2+project_in -> dwconv3x3 + GELU gate -> project_out + residual
   (shapes with L % 8 != 0 take the explicit cross_scan gather instead of the direction-aware scan)

No CPU path, no silent fallback: every stage goes through the C-ABI and raises on failure.
"""
from __future__ import annotations

import os

import torch

from . import ops

_DTYPES = (torch.float32, torch.bfloat16, torch.float16)


_warned = set()
_DW_T = os.environ.get("VMB_DW_T", "1") == "1"  # depthwise conv writes the transposed copy itself (0: separate vmb_transpose_hw pass)


def unsupported_reason(block, x: torch.Tensor):
    """None when the fused kernels cover this call, else the first kernel limit it violates.  Mirrors the checks of the
    launchers: LayerNorm / gate prologues keep the whole K resident (csrc/pixlin.cu pixlin_launch: K <= 384), the depthwise
    conv indexes (batch, channel) planes on gridDim.y (csrc/api.cu vmb_dwconv3x3: batch * channels <= 65535), the channel
    branch holds its (dc, C) problem in one CTA's shared memory (csrc/oss_ops.cu channel_launch), dstate <= 16."""
    if not (x.is_cuda and x.dim() == 4):
        return "not a CUDA (B,C,H,W) tensor"
    if x.dtype not in _DTYPES:
        return f"dtype {x.dtype}"
    a = block.attn
    if a.d_state > 16 or a.dc_state > 16:
        return f"d_state {a.d_state} / dc_state {a.dc_state} > 16"
    B, C = x.shape[0], x.shape[1]
    h2 = block.ffn.project_in.weight.shape[0]
    if C > 384:
        return f"C = {C} > 384 (LayerNorm / gate prologue keeps K resident)"
    if B * max(C, h2 // 2) > 65535 or B > 65535:
        return f"batch * channels = {B * max(C, h2 // 2)} > 65535 (depthwise-conv grid)"
    RN = a.dtc_rank + 2 * a.dc_state
    dc = a.dc_inner
    smem = 4 * (dc * C + 2 * RN * (C | 1) + 4 * dc * C + C + 64 + 2 * RN * dc + 2 * dc * a.dtc_rank + 5 * dc + 1)
    if smem > 227 * 1024:
        return f"channel branch needs {smem} B of shared memory"
    return None


def available(block, x: torch.Tensor) -> bool:
    why = unsupported_reason(block, x)
    if why is not None and x.is_cuda:
        key = (why, tuple(x.shape[1:]))
        if key not in _warned:  # once per (reason, shape): the composed torch + scan-operator path takes over, loudly
            _warned.add(key)
            import warnings
            warnings.warn(f"vmambair_b200: fused OSS block not available ({why}); using the composed path for x{tuple(x.shape)}")
    return why is None


def _f32(t):
    """fp32 kernel parameter; prefers the un-rounded fp32 master kept by engine.cast_for_inference"""
    if t is None:
        return None
    m = getattr(t, "_vmb_master", None)
    if m is not None and m.shape == t.shape and m.device == t.device:
        return m.contiguous()
    return t.detach().float().contiguous()


def _prepare(block, dtype, device):
    """Weights in kernel layout (cached per block; rebuilt when any parameter changed)."""
    ver = (dtype, device, sum(p._version for p in block.parameters()), tuple(p.data_ptr() for p in block.parameters()))
    cache = getattr(block, "_vmb_cache", None)
    if cache is not None and cache["ver"] == ver:
        return cache
    a, f = block.attn, block.ffn
    C, R, N = a.d_inner, a.dt_rank, a.d_state
    w = lambda t: ops.pad_weight(t.detach().to(dtype))
    ln = lambda m: (1 if hasattr(m.body, "bias") else 2, _f32(m.body.weight), _f32(getattr(m.body, "bias", None)))
    xw, dtw = a.x_proj_weight.detach().float(), a.dt_projs_weight.detach().float()
    big = torch.cat([torch.cat([dtw[k] @ xw[k, :R], xw[k, R:]], 0) for k in range(4)], 0)  # (4(C+2N), C)
    has_cio = hasattr(a, "conv_cin")
    c = dict(
        ver=ver, C=C, N=N, R=R, h=f.project_out.weight.shape[1],
        ln1=ln(block.norm1), ln2=ln(block.norm2),
        w_in=w(a.in_conv.weight.view(2 * C, C)), b_in=_f32(a.in_conv.bias),
        dw=_f32(a.conv2d.weight.view(C, 9)), dw_b=_f32(a.conv2d.bias),
        w_big=w(big),
        w_big02=w(torch.cat([big[0:C + 2 * N], big[2 * (C + 2 * N):3 * (C + 2 * N)]], 0)),
        w_big13=w(torch.cat([big[(C + 2 * N):2 * (C + 2 * N)], big[3 * (C + 2 * N):]], 0)),
        A=(-torch.exp(a.A_logs.detach().float())).contiguous(), Ds=_f32(a.Ds), dt_bias=_f32(a.dt_projs_bias.view(-1)),
        on_w=_f32(a.out_norm.body.weight), on_b=_f32(a.out_norm.body.bias),
        ch=dict(cin_w=_f32(a.conv_cin.weight.view(-1)) if has_cio else None, cin_b=_f32(a.conv_cin.bias) if has_cio else None,
                xc_proj=_f32(a.xc_proj_weight), dtc_w=_f32(a.dtc_projs_weight), dtc_b=_f32(a.dtc_projs_bias),
                Ac_logs=_f32(a.Ac_logs), Dsc=_f32(a.Dsc),
                cout_w=_f32(a.conv_cout.weight.view(-1)) if has_cio else None, cout_b=_f32(a.conv_cout.bias) if has_cio else None,
                cn_w=_f32(a.channel_norm.body.weight), cn_b=_f32(a.channel_norm.body.bias),
                dc=a.dc_inner, Rc=a.dtc_rank, N=a.dc_state),
        gate_mode=1 if a.gate == "mul" else 2,
        w_out=w(a.out_conv.weight.view(C, C)), b_out=_f32(a.out_conv.bias),
        w_pin=w(f.project_in.weight.view(-1, C)), b_pin=_f32(f.project_in.bias),
        fdw=_f32(f.dwconv.weight.view(-1, 9)), fdw_b=_f32(f.dwconv.bias),
        w_pout=w(f.project_out.weight.view(C, -1)), b_pout=_f32(f.project_out.bias),
    )
    block._vmb_cache = c
    return c


@torch.no_grad()
def block_forward(block, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """out: optional (B,C,H,W) destination with dense planes -- e.g. a channel slice of the decoder's concatenation buffer
    (unet.py): the block's last kernel stores there and torch.cat never runs."""
    B, C, H, W = x.shape
    L = H * W
    c = _prepare(block, x.dtype, x.device)
    N = c["N"]
    x3 = x.contiguous().view(B, C, L)
    # norm1 + in_conv; SiLU on the z half
    xz = ops.pixlin(x3, c["w_in"], c["b_in"], ln=c["ln1"], act=(C, 2 * C), static_w=True)
    grouped = L % 8 == 0 and os.environ.get("VMB_FUSED_SCAN", "grouped") == "grouped"
    if grouped and _DW_T:
        xc, xt = ops.dwconv3x3_t(xz[:, :C], c["dw"], c["dw_b"], C, H, W)  # x and its (W,H) copy from one launch
    else:
        xc = ops.dwconv3x3(xz[:, :C], c["dw"], c["dw_b"], C, H, W, 0)
    if grouped:
        # direction-aware scan: directions 0/2 read x and the GEMM on x, 1/3 read the transposed copies; reversed
        # directions walk the same memory backwards -- no gathered / flipped (B,4C,L) operands exist
        if not _DW_T:
            xt = ops.transpose_hw(xc, H, W)
        d02 = ops.pixlin(xc, c["w_big02"], static_w=True).view(B, 2, C + 2 * N, L)
        d13 = ops.pixlin(xt, c["w_big13"], static_w=True).view(B, 2, C + 2 * N, L)
        src = [(xc, d02[:, 0]), (xt, d13[:, 0]), (xc, d02[:, 1]), (xt, d13[:, 1])]
        ys = ops.selective_scan_fwd_grouped([s[0] for s in src], [s[1][:, :C] for s in src], [s[1][:, C:C + N] for s in src],
                                            [s[1][:, C + N:] for s in src], [0, 0, 1, 1], c["A"], c["Ds"], c["dt_bias"], True)
        y2, pooled = ops.merge_norm_gate(ys, xz[:, C:], c["on_w"], c["on_b"], C, H, W, in_place_order=True)
        return _block_tail(c, x3, y2, pooled, B, C, H, W, out)
    # delta (dt_proj o x_proj), B, C of the four directions from one GEMM on the un-permuted x
    dbl = ops.pixlin(xc, c["w_big"], static_w=True)  # (B, 4*(C+2N), L)
    dbl4 = dbl.view(B, 4, C + 2 * N, L)
    xs = ops.cross_scan([xc] * 4, C, H, W)
    dts = ops.cross_scan([dbl4[:, k, :C] for k in range(4)], C, H, W)
    bc = ops.cross_scan([dbl4[:, k, C:] for k in range(4)], 2 * N, H, W)
    ys, _ = ops.selective_scan_fwd(xs.view(B, 4 * C, L), dts.view(B, 4 * C, L), c["A"], bc[:, :, :N], bc[:, :, N:], c["Ds"],
                                   c["dt_bias"], True, need_ckpt=False)
    y2, pooled = ops.merge_norm_gate(ys.view(B, 4, C, L), xz[:, C:], c["on_w"], c["on_b"], C, H, W)
    return _block_tail(c, x3, y2, pooled, B, C, H, W, out)


def _block_tail(c, x3, y2, pooled, B, C, H, W, dst=None):
    L = H * W
    cg = ops.channel_branch(pooled, 1.0 / L, c["ch"], C)
    x1 = ops.pixlin(y2, c["w_out"], c["b_out"], residual=x3, gate=cg, gate_mode=c["gate_mode"], static_w=True)
    t = ops.pixlin(x1, c["w_pin"], c["b_pin"], ln=c["ln2"], static_w=True)
    g = ops.dwconv3x3(t, c["fdw"], c["fdw_b"], c["h"], H, W, 1)
    if dst is not None:
        assert dst.shape == (B, C, H, W) and dst.stride(3) == 1 and dst.stride(2) == W and dst.stride(1) == L
        ops.pixlin(g, c["w_pout"], c["b_pout"], residual=x1, static_w=True, out=dst.view(B, C, L))
        return dst
    out = ops.pixlin(g, c["w_pout"], c["b_pout"], residual=x1, static_w=True)
    return out.view(B, C, H, W)
