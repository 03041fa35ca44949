"""Training path of one OSS block on the fused kernels: forward AND backward of every stage as this library's kernels
(SURVEY.md 8 a15 -- the autograd of LayerNorm, in_conv, depthwise conv + SiLU, cross-scan, x_proj / dt_proj, the selective scan,
merge + out_norm + gate, out_conv, the EFFN), replacing the torch-autograd-over-eager-ops path of round 1.

Reference semantics: MamberBlock.forward under autograd, SRGAN/VmambaIR/archs/MambaSISR6_arch.py:486-515 (SS2D_1.forward :486-498,
forward_corev1 :395-436, FeedForward :201-218); the training step that drives it is MambaSISRModel.optimize_parameters
(SRGAN/VmambaIR/models/MambaSISR_model.py:120-147).

Structure (two autograd Functions with the tiny channel-direction branch between them left to torch autograd on (B, C) tensors):
    _Front:  x -> norm1 + in_conv -> dwconv + SiLU -> x_proj/dt_proj (folded GEMM) -> cross-scan -> selective scan (checkpoints)
             -> merge + out_norm + SiLU(z) gate -> (y2, pooled sums)
    c = SS2D_1.cforward_pooled(pooled / L)                      (torch ops, fp32, a few hundred bytes per image)
    _Tail:   (y2, c, x) -> channel gate + out_conv + residual -> norm2 + project_in -> dwconv + GELU gate -> project_out + residual

Backward kernels: vmb_pixlin with the transposed weight (data gradients of the five 1x1 convs), vmb_selective_scan_bwd,
vmb_cross_scan (the gathers pi_k and their inverses = pi_k with H and W swapped, bit-exact permutations),
vmb_merge_norm_gate_bwd, vmb_layernorm_fwd/bwd, vmb_dwconv3x3_bwd + vmb_dwconv3x3 (flipped taps), vmb_channel_gate_bwd.
The five weight-gradient contractions dW = sum_{b,p} dY X^T run vmb_pixlin_wgrad (mma.sync, split over pixels; fp32 parity mode:
the library GEMM); bias gradients are row sums.  Activations are saved in the compute dtype (bf16 under autocast), parameters stay fp32 and receive fp32 gradients.
"""
from __future__ import annotations

import torch

from . import ops


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


def _wt(w2d, dtype):
    """kernel-layout weight (rows zero-padded to 16 elements) and its transpose, in the compute dtype"""
    w = w2d.detach().to(dtype)
    return ops.pad_weight(w), ops.pad_weight(w.t().contiguous())


def _wgrad(dy, x):
    """dW[m,k] = sum_{b,p} dy[b,m,p] x[b,k,p]  (fp32 result; this library's split-pixel mma kernel for 16-bit activations)"""
    return ops.pixlin_wgrad(dy, x)


class _Front(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n1w, n1b, w_in, b_in, cw, cb, x_proj_w, dt_w, dt_b, A_logs, Ds, on_w, on_b):
        with torch.autocast("cuda", enabled=False):
            B, C, H, W = x.shape
            L, dt_ = H * W, x.dtype
            N, R = A_logs.shape[1], dt_w.shape[2]
            ln_mode = 1 if n1b is not None else 2
            x3 = x.contiguous().view(B, C, L)
            Win, WinT = _wt(w_in.view(2 * C, C), dt_)
            xz = ops.pixlin(x3, Win, _f32(b_in), ln=(ln_mode, _f32(n1w), _f32(n1b)))  # (B,2C,L): [x_pre | z_pre], no activation
            cw9 = _f32(cw.view(C, 9))
            xc = ops.dwconv3x3(xz[:, :C], cw9, _f32(cb), C, H, W, 0)
            xw, dtw = x_proj_w.detach().float(), dt_w.detach().float()
            big = torch.cat([torch.cat([dtw[k] @ xw[k, :R], xw[k, R:]], 0) for k in range(4)], 0)  # (4(C+2N), C)
            Wbig, WbigT = _wt(big, dt_)
            dbl4 = ops.pixlin(xc, Wbig).view(B, 4, C + 2 * N, L)   # delta, B, C of the four directions, pixel order
            xs = ops.cross_scan([xc] * 4, C, H, W)
            dts = ops.cross_scan([dbl4[:, k, :C] for k in range(4)], C, H, W)
            bc = ops.cross_scan([dbl4[:, k, C:] for k in range(4)], 2 * N, H, W)
            A = (-torch.exp(A_logs.detach().float())).contiguous()
            ys, ckpt = ops.selective_scan_fwd(xs.view(B, 4 * C, L), dts.view(B, 4 * C, L), A, bc[:, :, :N], bc[:, :, N:], _f32(Ds),
                                              _f32(dt_b.reshape(-1)), True, need_ckpt=True)
            y2, pooled, ws = ops.merge_norm_gate(ys.view(B, 4, C, L), xz[:, C:], _f32(on_w), _f32(on_b), C, H, W,
                                                 z_preact=True, return_ws=True)
            ctx.save_for_backward(x3, xz, xc, xs, dts, bc, ckpt, ws, A, n1w, n1b, w_in, cw, cb, x_proj_w, dt_w, dt_b, Ds, on_w, on_b,
                                  WinT, WbigT)
            ctx.geom = (B, C, H, W, N, R, ln_mode)
        return y2.view(B, C, H, W), pooled

    @staticmethod
    def backward(ctx, dy2, dpooled):
        with torch.autocast("cuda", enabled=False):
            (x3, xz, xc, xs, dts, bc, ckpt, ws, A, n1w, n1b, w_in, cw, cb, x_proj_w, dt_w, dt_b, Ds, on_w, on_b,
             WinT, WbigT) = ctx.saved_tensors
            B, C, H, W, N, R, ln_mode = ctx.geom
            L, dt_ = H * W, x3.dtype
            dy2 = dy2.to(dt_).contiguous().view(B, C, L)
            dpooled = None if dpooled is None else dpooled.float().contiguous()
            dxz = torch.empty_like(xz)  # [d x_pre | d z_pre]
            dm, d_onw, d_onb = ops.merge_norm_gate_bwd(ws, xz[:, C:], dy2, dpooled, _f32(on_w), _f32(on_b), C, L, dz_out=dxz[:, C:])
            dys = ops.cross_scan([dm] * 4, C, H, W)  # gradient of the merged output, gathered into the four scan orders
            du, ddelta, dA, dB, dC, dD, dbias = ops.selective_scan_bwd(
                xs.view(B, 4 * C, L), dts.view(B, 4 * C, L), A, bc[:, :, :N], bc[:, :, N:], _f32(Ds), _f32(dt_b.reshape(-1)),
                dys.view(B, 4 * C, L), ckpt, True)
            # back to pixel order: pi_k^-1 is pi_k with H and W swapped
            dxc4 = ops.cross_scan([du.view(B, 4, C, L)[:, k] for k in range(4)], C, W, H)
            ddbl = torch.empty((B, 4, C + 2 * N, L), dtype=dt_, device=x3.device)
            ops.cross_scan([ddelta.view(B, 4, C, L)[:, k] for k in range(4)], C, W, H, out=ddbl[:, :, :C])
            ops.cross_scan([dB[:, k] for k in range(4)], N, W, H, out=ddbl[:, :, C:C + N])
            ops.cross_scan([dC[:, k] for k in range(4)], N, W, H, out=ddbl[:, :, C + N:])
            ddbl = ddbl.view(B, 4 * (C + 2 * N), L)
            xw, dtw = x_proj_w.detach().float(), dt_w.detach().float()
            dxc = ops.pixlin(ddbl, WbigT)
            dxc = (dxc.float() + dxc4.float().sum(1)).to(dt_)
            dbig = _wgrad(ddbl, xc).view(4, C + 2 * N, C)
            d_xproj = torch.empty_like(xw)
            d_xproj[:, R:] = dbig[:, C:]
            d_xproj[:, :R] = torch.bmm(dtw.transpose(1, 2), dbig[:, :C])      # W_dt^T dW
            d_dtw = torch.bmm(dbig[:, :C], xw[:, :R].transpose(1, 2))         # dW W_x[:R]^T
            # depthwise conv + SiLU
            cw9 = _f32(cw.view(C, 9))
            dv, d_cw, d_cb = ops.dwconv3x3_bwd(xz[:, :C], cw9, _f32(cb), dxc, C, H, W, 0)
            ops.dwconv3x3(dv, cw9.flip(-1).contiguous(), None, C, H, W, 2, out=dxz[:, :C])
            # in_conv + norm1
            dxn = ops.pixlin(dxz, WinT)
            xn = ops.layernorm_fwd(x3, ln_mode, _f32(n1w), _f32(n1b))
            d_win = _wgrad(dxz, xn).view_as(w_in)
            d_bin = dxz.float().sum((0, 2))
            dx, d_n1w, d_n1b = ops.layernorm_bwd(x3, dxn, ln_mode, _f32(n1w))
            dA_logs = (dA * A).to(A.dtype)  # A = -exp(A_logs)
        return (dx.view(B, C, H, W), d_n1w, d_n1b, d_win, d_bin, d_cw.view_as(cw), d_cb, d_xproj, d_dtw, dbias.view_as(dt_b),
                dA_logs, dD, d_onw, d_onb)


class _Tail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y2, c, x, w_out, b_out, n2w, n2b, w_pin, b_pin, fdw, fdwb, w_pout, b_pout, gate_mode):
        with torch.autocast("cuda", enabled=False):
            B, C, H, W = x.shape
            L, dt_ = H * W, x.dtype
            h = w_pout.shape[1]
            ln_mode = 1 if n2b is not None else 2
            x3, y23 = x.contiguous().view(B, C, L), y2.contiguous().view(B, C, L)
            cg = c.detach().float().contiguous()
            Wout, WoutT = _wt(w_out.view(C, C), dt_)
            x1 = ops.pixlin(y23, Wout, _f32(b_out), residual=x3, gate=cg, gate_mode=gate_mode)
            Wpin, WpinT = _wt(w_pin.view(2 * h, C), dt_)
            t = ops.pixlin(x1, Wpin, _f32(b_pin), ln=(ln_mode, _f32(n2w), _f32(n2b)))
            fdw9 = _f32(fdw.view(2 * h, 9))
            gg = ops.dwconv3x3(t, fdw9, _f32(fdwb), h, H, W, 1)
            Wpout, WpoutT = _wt(w_pout.view(C, h), dt_)
            out = ops.pixlin(gg, Wpout, _f32(b_pout), residual=x1)
            ctx.save_for_backward(y23, cg, x1, t, gg, w_out, b_out, n2w, n2b, w_pin, b_pin, fdw, fdwb, w_pout, b_pout,
                                  WoutT, WpinT, WpoutT)
            ctx.geom = (B, C, H, W, h, ln_mode, gate_mode)
        return out.view(B, C, H, W)

    @staticmethod
    def backward(ctx, dout):
        with torch.autocast("cuda", enabled=False):
            (y23, cg, x1, t, gg, w_out, b_out, n2w, n2b, w_pin, b_pin, fdw, fdwb, w_pout, b_pout,
             WoutT, WpinT, WpoutT) = ctx.saved_tensors
            B, C, H, W, h, ln_mode, gate_mode = ctx.geom
            L, dt_ = H * W, x1.dtype
            dout3 = dout.to(dt_).contiguous().view(B, C, L)
            # project_out
            dgg = ops.pixlin(dout3, WpoutT)
            d_wpout = _wgrad(dout3, gg).view_as(w_pout)
            d_bpout = dout3.float().sum((0, 2)) if b_pout is not None else None
            # depthwise conv + GELU gate
            fdw9 = _f32(fdw.view(2 * h, 9))
            dv, d_fdw, d_fdwb = ops.dwconv3x3_bwd(t, fdw9, _f32(fdwb), dgg, h, H, W, 1)
            dt = ops.dwconv3x3(dv, fdw9.flip(-1).contiguous(), None, 2 * h, H, W, 2)
            # project_in + norm2 (+ the residual branch of the EFFN)
            dx1n = ops.pixlin(dt, WpinT)
            x1n = ops.layernorm_fwd(x1, ln_mode, _f32(n2w), _f32(n2b))
            d_wpin = _wgrad(dt, x1n).view_as(w_pin)
            d_bpin = dt.float().sum((0, 2)) if b_pin is not None else None
            dx1, d_n2w, d_n2b = ops.layernorm_bwd(x1, dx1n, ln_mode, _f32(n2w), add=dout3)
            # out_conv with the channel gate in front, residual behind
            dyg = ops.pixlin(dx1, WoutT)
            dy2, dc = ops.channel_gate_bwd(dyg, y23, cg, gate_mode)
            wb = ops.pixlin_wgrad(dx1, y23, per_batch=True)  # (B, C_out, C_in): the channel gate scales it per image
            if gate_mode == 1:
                d_wout = (wb * (1.0 + cg)[:, None, :]).sum(0)
            else:
                d_wout = wb.sum(0) + torch.einsum("bm,bk->mk", dx1.float().sum(2), cg)
            d_bout = dx1.float().sum((0, 2)) if b_out is not None else None
        return (dy2.view(B, C, H, W), dc, dx1.view(B, C, H, W), d_wout.view_as(w_out), d_bout, d_n2w, d_n2b, d_wpin, d_bpin,
                d_fdw.view_as(fdw), d_fdwb, d_wpout, d_bpout, None)


def block_forward(block, x: torch.Tensor) -> torch.Tensor:
    """MamberBlock.forward with autograd on the fused kernels (x: (B,C,H,W) CUDA tensor in the compute dtype)."""
    a, f = block.attn, block.ffn
    L = x.shape[2] * x.shape[3]
    n1, n2, on = block.norm1.body, block.norm2.body, a.out_norm.body
    y2, pooled = _Front.apply(x, n1.weight, getattr(n1, "bias", None), a.in_conv.weight, a.in_conv.bias, a.conv2d.weight,
                              a.conv2d.bias, a.x_proj_weight, a.dt_projs_weight, a.dt_projs_bias, a.A_logs, a.Ds, on.weight, on.bias)
    with torch.autocast("cuda", enabled=False):
        c = a.cforward_pooled(pooled * (1.0 / L))  # (B, C) fp32, torch autograd (channel-direction OSS on the pooled descriptor)
    return _Tail.apply(y2, c, x, a.out_conv.weight, a.out_conv.bias, n2.weight, getattr(n2, "bias", None), f.project_in.weight,
                       f.project_in.bias, f.dwconv.weight, f.dwconv.bias, f.project_out.weight, f.project_out.bias,
                       1 if a.gate == "mul" else 2)
