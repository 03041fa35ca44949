"""Thin torch-tensor front-end of the C-ABI (device pointers + strides -> library call).

Mirrors the reference binding `selective_scan_cuda_core.fwd/bwd`
(Mamba/kernels/selective_scan/csrc/selective_scan/cus/selective_scan.cpp:157-349): same argument
meaning, same checks, same RuntimeError behaviour.  Outputs are allocated here with torch (the
library owns nothing), on the input's device and current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


_LAUNCHES = 0          # kernels of this library launched so far (host-side count)
_TIMING = None         # when a list: (tag, bytes, start_event, end_event) per library call (bench roofline)


def launch_count() -> int:
    return _LAUNCHES


def set_timing(records):
    """records: None (off) or a list that receives (tag, algorithmic_bytes, start_event, end_event)."""
    global _TIMING
    _TIMING = records


class _timed:
    def __init__(self, tag, nbytes, dev, kernels):
        self.tag, self.nbytes, self.dev, self.kernels = tag, nbytes, dev, kernels

    def __enter__(self):
        if _TIMING is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *a):
        global _LAUNCHES
        _LAUNCHES += self.kernels
        if _TIMING is not None:
            self.e.record(torch.cuda.current_stream(self.dev))
            _TIMING.append((self.tag, self.nbytes, self.s, self.e))


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check_common(u, delta, A, B, C_, D, delta_bias, who):
    if u.dtype not in _DT:
        raise RuntimeError(f"{who}: input dtype must be float32/float16/bfloat16, got {u.dtype}")
    if A.dtype != torch.float32:
        raise RuntimeError(f"{who}: A must be float32")
    for n, t in (("delta", delta), ("B", B), ("C", C_)):
        if t.dtype != u.dtype:
            raise RuntimeError(f"{who}: {n}.dtype {t.dtype} != u.dtype {u.dtype}")
    for n, t in (("u", u), ("delta", delta), ("A", A), ("B", B), ("C", C_)):
        if not t.is_cuda:
            raise RuntimeError(f"{who}: {n} must be a CUDA tensor")
    if u.dim() != 3 or B.dim() != 4 or C_.dim() != 4 or A.dim() != 2:
        raise RuntimeError(f"{who}: expected u (B,D,L), A (D,N), B/C (B,G,N,L)")
    b, d, l = u.shape
    n = A.shape[1]
    g = B.shape[1]
    if tuple(delta.shape) != (b, d, l) or tuple(A.shape) != (d, n) or tuple(B.shape) != (b, g, n, l) \
            or tuple(C_.shape) != (b, g, n, l):
        raise RuntimeError(f"{who}: shape mismatch u{tuple(u.shape)} delta{tuple(delta.shape)} A{tuple(A.shape)} "
                           f"B{tuple(B.shape)} C{tuple(C_.shape)}")
    for nme, t in (("u", u), ("delta", delta), ("B", B), ("C", C_)):
        if t.stride(-1) != 1 and t.size(-1) != 1:
            raise RuntimeError(f"{who}: {nme}.stride(-1) must be 1")
    if not A.is_contiguous():
        raise RuntimeError(f"{who}: A must be contiguous")
    for nme, t in (("D", D), ("delta_bias", delta_bias)):
        if t is not None:
            if t.dtype != torch.float32 or not t.is_cuda or tuple(t.shape) != (d,) or not t.is_contiguous():
                raise RuntimeError(f"{who}: {nme} must be a contiguous float32 CUDA tensor of shape ({d},)")
    return b, d, l, n, g


def ckpt_interval() -> int:
    return _lib.lib().vmb_scan_ckpt_interval()


def selective_scan_fwd(u, delta, A, B, C_, D=None, delta_bias=None, delta_softplus=False, need_ckpt=True):
    """-> (out, ckpt).  out has delta's dtype/shape; ckpt is the opaque fp32 checkpoint tensor."""
    b, d, l, n, g = _check_common(u, delta, A, B, C_, D, delta_bias, "selective_scan_fwd")
    L = _lib.lib()
    out = torch.empty_like(delta)
    if out.stride(-1) != 1:
        out = torch.empty(delta.shape, dtype=delta.dtype, device=delta.device)
    ck = ckpt_interval()
    ckpt = torch.empty((b, d, (l + ck - 1) // ck, n), dtype=torch.float32, device=u.device) if need_ckpt else None
    a = _lib.ScanFwdArgs(
        _ptr(u), _ptr(delta), _ptr(A), _ptr(B), _ptr(C_), _ptr(D), _ptr(delta_bias), _ptr(out), _ptr(ckpt),
        b, d, l, n, g,
        u.stride(0), u.stride(1), delta.stride(0), delta.stride(1), out.stride(0), out.stride(1),
        B.stride(0), B.stride(1), B.stride(2), C_.stride(0), C_.stride(1), C_.stride(2),
        int(bool(delta_softplus)), _DT[u.dtype])
    es = u.element_size()
    nbytes = es * (3 * b * d * l + 2 * b * g * n * l) + 4 * (d * n + 2 * d)  # every operand once
    with torch.cuda.device(u.device), _timed("scan_fwd", nbytes, u.device, 1):
        _lib.check(L.vmb_selective_scan_fwd(C.byref(a), _stream(u)), "selective_scan_fwd")
    return out, ckpt


def selective_scan_bwd(u, delta, A, B, C_, D, delta_bias, dout, ckpt, delta_softplus=False, zeroed=None):
    """-> (du, ddelta, dA, dB, dC, dD, ddelta_bias); dB/dC returned in B/C dtype (reference cpp:347).
    zeroed: optional (dA, dD, dbias) pre-zeroed fp32 buffers (the training path hands out slices of one zeroed workspace)."""
    b, d, l, n, g = _check_common(u, delta, A, B, C_, D, delta_bias, "selective_scan_bwd")
    if dout.dtype != u.dtype or tuple(dout.shape) != (b, d, l) or (dout.stride(-1) != 1 and dout.size(-1) != 1):
        raise RuntimeError("selective_scan_bwd: dout must match u in dtype/shape with stride(-1)==1")
    L = _lib.lib()
    du = torch.empty((b, d, l), dtype=u.dtype, device=u.device)
    ddelta = torch.empty((b, d, l), dtype=u.dtype, device=u.device)
    dA = torch.zeros_like(A) if zeroed is None else zeroed[0]
    dB = torch.empty((b, g, n, l), dtype=u.dtype, device=u.device)
    dC = torch.empty((b, g, n, l), dtype=u.dtype, device=u.device)
    ws = torch.empty(L.vmb_scan_bwd_workspace_bytes(b, g, n, l), dtype=torch.uint8, device=u.device)
    dD = (torch.zeros_like(D) if zeroed is None else zeroed[1]) if D is not None else None
    dbias = (torch.zeros_like(delta_bias) if zeroed is None else zeroed[2]) if delta_bias is not None else None
    a = _lib.ScanBwdArgs(
        _ptr(u), _ptr(delta), _ptr(A), _ptr(B), _ptr(C_), _ptr(D), _ptr(delta_bias), _ptr(dout), _ptr(ckpt),
        _ptr(du), _ptr(ddelta), _ptr(dA), _ptr(dB), _ptr(dC), _ptr(dD), _ptr(dbias), _ptr(ws),
        b, d, l, n, g,
        u.stride(0), u.stride(1), delta.stride(0), delta.stride(1), dout.stride(0), dout.stride(1),
        du.stride(0), du.stride(1), ddelta.stride(0), ddelta.stride(1),
        B.stride(0), B.stride(1), B.stride(2), C_.stride(0), C_.stride(1), C_.stride(2),
        int(bool(delta_softplus)), _DT[u.dtype])
    es = u.element_size()
    nbytes = es * (5 * b * d * l + 4 * b * g * n * l) + 4 * b * d * ((l + 63) // 64) * n
    with torch.cuda.device(u.device), _timed("scan_bwd", nbytes, u.device, 2):
        _lib.check(L.vmb_selective_scan_bwd(C.byref(a), _stream(u)), "selective_scan_bwd")
    return du, ddelta, dA, dB, dC, dD, dbias


# ----------------------------------------------------------------------------- fused OSS-block stages
def _run(name, args, t, tag, nbytes=0, kernels=1):
    fn = getattr(_lib.lib(), name)
    with torch.cuda.device(t.device), _timed(tag, nbytes, t.device, kernels):
        _lib.check(fn(C.byref(args), _stream(t)), name)


def pixlin(x, w, bias=None, residual=None, ln=None, gate=None, gate_mode=0, act=(0, 0), out_dtype=None, out=None, static_w=False):
    """out[b,m,p] = epi(sum_k w[m,k] pro(x)[b,k,p]).  x: (B,K,P) view with stride(-1)==1; w: (M,>=K) x.dtype, rows may be
    zero-padded to a multiple of 16 (pad_weight) for vector loads.
    ln = (mode, weight_fp32, bias_fp32|None); gate: (B,K) fp32; residual: (B,M,P) view.
    static_w: w / bias / ln parameters were written long before this call (cached inference weights), so the kernel may fetch them
    while the preceding kernel of the stream is still running; leave False when they are that kernel's output."""
    B, K, P = x.shape
    M = w.shape[0]
    assert w.shape[1] >= K and w.dtype == x.dtype and w.stride(1) == 1 and x.stride(2) == 1
    out_dtype = out_dtype or x.dtype
    if out is None:
        out = torch.empty((B, M, P), dtype=out_dtype, device=x.device)
    ln_mode, ln_w, ln_b = ln if ln is not None else (0, None, None)
    a = _lib.PixlinArgs(
        _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), _ptr(out), _ptr(ln_w), _ptr(ln_b), _ptr(gate),
        ln_mode, gate_mode if gate is not None else 0, act[0], act[1], B, K, M, P,
        x.stride(0), x.stride(1), residual.stride(0) if residual is not None else 0,
        residual.stride(1) if residual is not None else 0, out.stride(0), out.stride(1),
        gate.stride(0) if gate is not None else 0, w.stride(0), _DT[x.dtype], _DT[out_dtype], int(static_w))
    _run("vmb_pixlin", a, x, "pixlin", 0)
    return out


def pad_weight(w: torch.Tensor) -> torch.Tensor:
    """(M,K) -> (M, ceil16(K)) zero-padded copy: rows 16 B aligned for the vectorised weight staging."""
    M, K = w.shape
    kp = (K + 15) // 16 * 16
    if kp == K:
        return w.contiguous()
    out = torch.zeros((M, kp), dtype=w.dtype, device=w.device)
    out[:, :K] = w
    return out


def dwconv3x3(x, w9, bias, c_out, H, W, mode, out=None):
    """x: (B, Cin, H*W) view; mode 0: SiLU(dw(x[:c_out])); mode 1: gelu(dw(x[:c_out])) * dw(x[c_out:2c_out]);
    mode 2: dw(x[:c_out]) (no activation: the transposed conv of the backward pass, called with flipped taps)."""
    B = x.shape[0]
    if out is None:
        out = torch.empty((B, c_out, H * W), dtype=x.dtype, device=x.device)
    a = _lib.DwconvArgs(_ptr(x), _ptr(w9), _ptr(bias), _ptr(out), B, c_out, H, W, mode,
                        x.stride(0), x.stride(1), out.stride(0), out.stride(1), _DT[x.dtype])
    _run("vmb_dwconv3x3", a, x, "dwconv")
    return out


def dwconv3x3_t(x, w9, bias, c_out, H, W):
    """SiLU(dw(x[:c_out])) and its (W,H)-transposed copy from ONE launch -> (out (B,c_out,H*W), out_t (B,c_out,W*H))."""
    B = x.shape[0]
    out = torch.empty((B, c_out, H * W), dtype=x.dtype, device=x.device)
    out_t = torch.empty_like(out)
    a = _lib.DwconvArgs(_ptr(x), _ptr(w9), _ptr(bias), _ptr(out), B, c_out, H, W, 0,
                        x.stride(0), x.stride(1), out.stride(0), out.stride(1), _DT[x.dtype])
    fn = _lib.lib().vmb_dwconv3x3_t
    with torch.cuda.device(x.device), _timed("dwconv", 0, x.device, 1):
        _lib.check(fn(C.byref(a), _ptr(out_t), _stream(x)), "vmb_dwconv3x3_t")
    return out, out_t


def cross_scan(srcs, rows, H, W, out=None):
    """srcs: 4 views (B, rows, L) sharing strides -> (B, 4, rows, L) in scan order.  `out`: optional (B, 4, rows, L) view
    (e.g. a channel slice of a wider tensor) with contiguous rows.  The inverse orders are the same call with H and W swapped."""
    s0 = srcs[0]
    B = s0.shape[0]
    if out is None:
        out = torch.empty((B, 4, rows, H * W), dtype=s0.dtype, device=s0.device)
    assert out.shape == (B, 4, rows, H * W) and out.stride(3) == 1 and out.stride(2) == H * W and out.dtype == s0.dtype
    arr = (C.c_void_p * 4)(*[t.data_ptr() for t in srcs])
    for t in srcs:
        assert t.stride() == s0.stride() and t.stride(2) == 1
    a = _lib.CrossScanArgs(arr, _ptr(out), B, rows, H, W, s0.stride(0), s0.stride(1), out.stride(0), _DT[s0.dtype], out.stride(1))
    _run("vmb_cross_scan", a, s0, "cross_scan")
    return out


def pixel_shuffle2_nhwc(x, bias=None):
    """nn.PixelShuffle(2) on a channels-last tensor: x logical (B,4C,H,W) with NHWC storage -> (B,C,2H,2W), NHWC storage.
    Bit-identical to F.pixel_shuffle(x, 2).  bias: optional fp32 (4C), added to x first (the bias of the conv that produced x)."""
    B, C4, H, W = x.shape
    assert C4 % 4 == 0 and x.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((B, C4 // 4, 2 * H, 2 * W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    a = _lib.PixelShuffleArgs(_ptr(x), _ptr(out), B, H, W, C4 // 4, _DT[x.dtype])
    if bias is None:
        _run("vmb_pixel_shuffle2_nhwc", a, x, "pixel_shuffle")
    else:
        assert bias.dtype == torch.float32 and bias.numel() == C4 and bias.is_contiguous() and bias.device == x.device
        fn = _lib.lib().vmb_pixel_shuffle2_nhwc_bias
        with torch.cuda.device(x.device), _timed("pixel_shuffle", 0, x.device, 1):
            _lib.check(fn(C.byref(a), _ptr(bias), _stream(x)), "vmb_pixel_shuffle2_nhwc_bias")
    return out


def cross_scan_multi(segs, H, W):
    """segs: up to 4 tuples (srcs, rows, out|None) as for cross_scan, all of one geometry -> list of outputs, ONE launch."""
    assert 1 <= len(segs) <= 4
    arr = (_lib.CrossScanArgs * len(segs))()
    outs = []
    s00 = segs[0][0][0]
    B = s00.shape[0]
    for i, (srcs, rows, out) in enumerate(segs):
        s0 = srcs[0]
        if out is None:
            out = torch.empty((B, 4, rows, H * W), dtype=s0.dtype, device=s0.device)
        assert out.shape == (B, 4, rows, H * W) and out.stride(3) == 1 and out.stride(2) == H * W and out.dtype == s00.dtype
        for t in srcs:
            assert t.stride() == s0.stride() and t.stride(2) == 1 and t.dtype == s00.dtype and t.shape[0] == B
        arr[i] = _lib.CrossScanArgs((C.c_void_p * 4)(*[t.data_ptr() for t in srcs]), _ptr(out), B, rows, H, W, s0.stride(0),
                                    s0.stride(1), out.stride(0), _DT[s0.dtype], out.stride(1))
        outs.append(out)
    fn = _lib.lib().vmb_cross_scan_multi
    with torch.cuda.device(s00.device), _timed("cross_scan", 0, s00.device, 1):
        _lib.check(fn(arr, len(segs), _stream(s00)), "vmb_cross_scan_multi")
    return outs


CONV_PLAIN, CONV_UNSHUFFLE2, CONV_SHUFFLE2, CONV_ADD_NEAREST = 0, 1, 2, 3


def pack_conv3x3_weight(w: torch.Tensor, dtype) -> torch.Tensor:
    """nn.Conv2d weight (Cout, Cin, 3, 3) -> the kernel layout [tap = 3 ky + kx][Cout rounded up to 64][Cin rounded up to 16], zero
    padded, in the I/O dtype (include/vmambair_b200.h, vmb_conv3x3)."""
    Cout, Cin = w.shape[0], w.shape[1]
    assert w.shape[2:] == (3, 3)
    out = torch.zeros((9, (Cout + 63) // 64 * 64, (Cin + 15) // 16 * 16), dtype=dtype, device=w.device)
    out[:, :Cout, :Cin] = w.detach().permute(2, 3, 0, 1).reshape(9, Cout, Cin).to(dtype)
    return out


def conv3x3(x, wp, bias, c_out, mode=CONV_PLAIN, out=None, add=None, add_scale=1, nhwc=False):
    """Dense 3x3 conv (stride 1, pad 1) with the PixelUnshuffle(2) / PixelShuffle(2) / channel-slice / "+ nearest-upsampled image"
    store folded in.  x: (B, Cin, H, W), NCHW view with dense planes (any batch / channel stride) or, nhwc=True, channels_last
    contiguous.  wp: pack_conv3x3_weight(...).  bias: fp32 (c_out) or None.  out: optional NCHW destination with dense planes (e.g.
    a channel slice of the decoder's concatenation buffer).  add (mode CONV_ADD_NEAREST): (B, c_out, H/s, W/s), x.dtype."""
    B, Cin, H, W = x.shape
    if nhwc:
        assert x.is_contiguous(memory_format=torch.channels_last)
        x_bs, x_cs = H * W * Cin, 1
    else:
        assert x.stride(3) == 1 and x.stride(2) == W, "conv3x3: NCHW input planes must be dense"
        x_bs, x_cs = x.stride(0), x.stride(1)
    assert wp.dtype == x.dtype and wp.is_contiguous() and wp.shape == (9, (c_out + 63) // 64 * 64, (Cin + 15) // 16 * 16)
    shape = {CONV_PLAIN: (B, c_out, H, W), CONV_UNSHUFFLE2: (B, 4 * c_out, H // 2, W // 2),
             CONV_SHUFFLE2: (B, c_out // 4, 2 * H, 2 * W), CONV_ADD_NEAREST: (B, c_out, H, W)}[mode]
    if out is None:
        out = torch.empty(shape, dtype=x.dtype, device=x.device)
    assert tuple(out.shape) == shape and out.dtype == x.dtype and out.stride(3) == 1 and out.stride(2) == shape[3]
    add_bs = add_cs = 0
    if mode == CONV_ADD_NEAREST:
        assert add is not None and add.dtype == x.dtype and add.shape == (B, c_out, H // add_scale, W // add_scale)
        add = add.contiguous()
        add_bs, add_cs = add.stride(0), add.stride(1)
    a = _lib.Conv3x3Args(_ptr(x), _ptr(wp), _ptr(bias), _ptr(out), _ptr(add) if mode == CONV_ADD_NEAREST else None,
                         B, Cin, c_out, H, W, int(nhwc), mode, add_scale, x_bs, x_cs, out.stride(0), out.stride(1), add_bs, add_cs,
                         _DT[x.dtype])
    _run("vmb_conv3x3", a, x, "conv3x3")
    return out


def transpose_hw(x, H, W):
    """x: (B, C, H*W) contiguous -> (B, C, W*H) with every plane transposed."""
    assert x.is_contiguous()
    out = torch.empty_like(x)
    a = _lib.TransposeArgs(_ptr(x), _ptr(out), x.shape[0] * x.shape[1], H, W, _DT[x.dtype])
    _run("vmb_transpose_hw", a, x, "transpose")
    return out


def selective_scan_fwd_grouped(us, deltas, Bs, Cs, revs, A, D, delta_bias, delta_softplus=True):
    """Direction-aware forward: group g reads us[g] (B,R,L), deltas[g] (B,R,L), Bs[g]/Cs[g] (B,N,L) (views sharing
    strides) and walks them backwards when revs[g]; returns (B, G, R, L) with group g written in ITS SOURCE's memory order."""
    G = len(us)
    b, R, L = us[0].shape
    N = A.shape[1]
    for lst in (us, deltas, Bs, Cs):
        for t in lst:
            assert t.stride() == lst[0].stride() and t.stride(2) == 1 and t.dtype == us[0].dtype
    out = torch.empty((b, G, R, L), dtype=us[0].dtype, device=us[0].device)
    arr = lambda lst: (C.c_void_p * 4)(*([t.data_ptr() for t in lst] + [0] * (4 - G)))
    outs = [out[:, g] for g in range(G)]
    a = _lib.ScanGroupedArgs(
        arr(us), arr(deltas), arr(Bs), arr(Cs), arr(outs), (C.c_int * 4)(*(list(map(int, revs)) + [0] * (4 - G))),
        _ptr(A), _ptr(D), _ptr(delta_bias), b, G * R, L, N, G,
        us[0].stride(0), us[0].stride(1), deltas[0].stride(0), deltas[0].stride(1), out.stride(0), out.stride(2),
        Bs[0].stride(0), Bs[0].stride(1), Cs[0].stride(0), Cs[0].stride(1), int(bool(delta_softplus)), _DT[us[0].dtype])
    es = us[0].element_size()
    n_u = len({t.data_ptr() for t in us})  # forward and reversed directions share one source tensor: counted once
    nbytes = es * b * L * (n_u * R + G * R + G * R + 2 * G * N) + 4 * (G * R * N + 2 * G * R)  # every unique operand / result once
    _run("vmb_selective_scan_fwd_grouped", a, us[0], "scan_fwd", nbytes)
    return out


def merge_norm_gate(ys, z, ln_w, ln_b, C_, H, W, in_place_order=False, z_preact=False, return_ws=False):
    """ys: (B,4,C,L) contiguous; z: (B,C,L) view -> (y2 (B,C,L), pooled sums (B,C) fp32 [, workspace]).
    z_preact: z holds the pre-activation (SiLU applied here); return_ws: also return the fp32 workspace (merged scan output
    (B,C,L) followed by the per-pixel (sum, sum of squares)) that vmb_merge_norm_gate_bwd consumes."""
    B = ys.shape[0]
    assert ys.is_contiguous() and z.stride(2) == 1
    y2 = torch.empty((B, C_, H * W), dtype=ys.dtype, device=ys.device)
    pooled = torch.empty((B, C_), dtype=torch.float32, device=ys.device)
    ws = torch.empty(_lib.lib().vmb_merge_workspace_bytes(B, C_, H, W), dtype=torch.uint8, device=ys.device)
    a = _lib.MergeArgs(_ptr(ys), _ptr(z), _ptr(ln_w), _ptr(ln_b), _ptr(y2), _ptr(pooled), B, C_, H, W,
                       z.stride(0), z.stride(1), _DT[ys.dtype], _ptr(ws), int(in_place_order), int(z_preact), int(return_ws))
    _run("vmb_merge_norm_gate", a, ys, "merge", 0, 2)
    return (y2, pooled, ws) if return_ws else (y2, pooled)


def channel_branch(pooled, inv_count, prm, C_):
    B = pooled.shape[0]
    c = torch.empty((B, C_), dtype=torch.float32, device=pooled.device)
    a = _lib.ChannelArgs(_ptr(pooled), inv_count, _ptr(prm["cin_w"]), _ptr(prm["cin_b"]), _ptr(prm["xc_proj"]),
                         _ptr(prm["dtc_w"]), _ptr(prm["dtc_b"]), _ptr(prm["Ac_logs"]), _ptr(prm["Dsc"]), _ptr(prm["cout_w"]),
                         _ptr(prm["cout_b"]), _ptr(prm["cn_w"]), _ptr(prm["cn_b"]), _ptr(c), B, C_, prm["dc"], prm["Rc"], prm["N"])
    _run("vmb_channel_branch", a, pooled, "channel")
    return c


# ----------------------------------------------------------------------------- training path: backward stages
def layernorm_fwd(x, mode, w, b):
    """x: (B,C,L) view -> LayerNorm over C per pixel, materialised (B,C,L) contiguous (mode 1 WithBias / 2 BiasFree)."""
    B, C_, L = x.shape
    assert x.stride(2) == 1
    y = torch.empty((B, C_, L), dtype=x.dtype, device=x.device)
    a = _lib.LnFwdArgs(_ptr(x), _ptr(w), _ptr(b), _ptr(y), None, B, C_, L, mode, x.stride(0), x.stride(1), y.stride(0), y.stride(1),
                       _DT[x.dtype])
    _run("vmb_layernorm_fwd", a, x, "ln_fwd")
    return y


def layernorm_bwd(x, g, mode, w, add=None, need_param_grads=True, zeroed=None, phase="both", stats=None):
    """-> (dx (B,C,L), dw (C) fp32, db (C) fp32 | None): LayerNorm backward of g [+ add] (the residual branch's gradient).
    phase "data": dx only -> (dx, stats);  phase "params" (stats of the "data" call): dw / db only -> (dw, db)."""
    B, C_, L = x.shape
    assert x.stride(2) == 1 and g.stride(2) == 1 and g.dtype == x.dtype and (add is None or (add.stride(2) == 1 and add.dtype == x.dtype))
    assert phase in ("both", "data", "params") and (phase != "params" or stats is not None)
    dx = torch.empty((B, C_, L), dtype=x.dtype, device=x.device) if phase != "params" else None
    if stats is None:
        stats = torch.empty((B, L, 2), dtype=torch.float32, device=x.device)
    dw = db = None
    if phase != "data":
        if zeroed is not None:
            dw, db = zeroed[0], (zeroed[1] if mode == 1 else None)
        elif need_param_grads:
            dw = torch.zeros(C_, dtype=torch.float32, device=x.device)
            db = torch.zeros(C_, dtype=torch.float32, device=x.device) if mode == 1 else None
    a = _lib.LnBwdArgs(_ptr(x), _ptr(g), _ptr(add), _ptr(w), _ptr(dx), _ptr(dw), _ptr(db), _ptr(stats), B, C_, L, mode,
                       x.stride(0), x.stride(1), g.stride(0), g.stride(1), add.stride(0) if add is not None else 0,
                       add.stride(1) if add is not None else 0, dx.stride(0) if dx is not None else 0, dx.stride(1) if dx is not None else 0,
                       _DT[x.dtype])
    _run("vmb_layernorm_bwd", a, x, "ln_bwd", 0, (1 if dx is not None else 0) + (1 if dw is not None else 0))
    if phase == "data":
        return dx, stats
    if phase == "params":
        return dw, db
    return dx, dw, db


def merge_norm_gate_bwd(ws, z, dy2, dpooled, ln_w, ln_b, C_, L, dz_out, zeroed=None, phase="both"):
    """backward of merge_norm_gate(z_preact=True).  ws: the forward's workspace; dy2 (B,C,L) contiguous; dpooled (B,C) fp32|None;
    dz_out: (B,C,L) view receiving the gradient w.r.t. the pre-activation z.  -> (dm (B,C,L), d ln_w, d ln_b).
    phase "data": -> dm (dz_out written);  phase "params": -> (d ln_w, d ln_b) only."""
    B = dy2.shape[0]
    assert dy2.is_contiguous() and z.stride(2) == 1 and (dz_out is None or dz_out.stride(2) == 1) and phase in ("both", "data", "params")
    merged = ws.view(torch.float32)
    stats = merged[B * C_ * L:]
    dm = torch.empty((B, C_, L), dtype=dy2.dtype, device=dy2.device) if phase != "params" else None
    dz = dz_out if phase != "params" else None
    dw = db = None
    if phase != "data":
        dw = torch.zeros(C_, dtype=torch.float32, device=dy2.device) if zeroed is None else zeroed[0]
        db = torch.zeros(C_, dtype=torch.float32, device=dy2.device) if zeroed is None else zeroed[1]
    a = _lib.MergeBwdArgs(_ptr(merged), _ptr(stats), _ptr(z), _ptr(dy2), _ptr(dpooled), _ptr(ln_w), _ptr(ln_b), _ptr(dm), _ptr(dz),
                          _ptr(dw), _ptr(db), B, C_, L, z.stride(0), z.stride(1), dz.stride(0) if dz is not None else 0,
                          dz.stride(1) if dz is not None else 0, _DT[dy2.dtype])
    _run("vmb_merge_norm_gate_bwd", a, dy2, "merge_bwd", 0, (1 if dm is not None else 0) + (1 if dw is not None else 0))
    if phase == "data":
        return dm
    if phase == "params":
        return dw, db
    return dm, dw, db


def dwconv3x3_bwd(x, w9, bias, g, c_out, H, W, mode, zeroed=None, phase="both", dv=None):
    """backward of dwconv3x3 (mode 0 / 1) up to the conv output: -> (dv (B, channels, L), dw9 (channels, 9) fp32, dbias (channels) fp32|None);
    the input gradient is dwconv3x3(dv, w9.flip(-1), None, channels, H, W, 2).
    phase "data": -> dv only;  phase "params" (dv of the "data" call, g unused): -> (dw9, dbias) only."""
    B = x.shape[0]
    ch = c_out * (2 if mode else 1)
    assert x.stride(2) == 1 and phase in ("both", "data", "params")
    if phase == "params":
        assert dv is not None and dv.dtype == x.dtype
        g = None
    else:
        assert g.stride(2) == 1 and g.dtype == x.dtype
        dv = torch.empty((B, ch, H * W), dtype=x.dtype, device=x.device)
    dw = dbias = None
    if phase != "data":
        dw = torch.zeros((ch, 9), dtype=torch.float32, device=x.device) if zeroed is None else zeroed[0]
        dbias = (torch.zeros(ch, dtype=torch.float32, device=x.device) if zeroed is None else zeroed[1]) if bias is not None else None
    a = _lib.DwconvBwdArgs(_ptr(x), _ptr(w9), _ptr(bias), _ptr(g), _ptr(dv), _ptr(dw), _ptr(dbias), B, c_out, H, W, mode,
                           x.stride(0), x.stride(1), g.stride(0) if g is not None else 0, g.stride(1) if g is not None else 0,
                           dv.stride(0), dv.stride(1), _DT[x.dtype])
    _run("vmb_dwconv3x3_bwd", a, x, "dwconv_bwd", 0, (1 if g is not None else 0) + (1 if dw is not None else 0))
    if phase == "data":
        return dv
    if phase == "params":
        return dw, dbias
    return dv, dw, dbias


def channel_gate_bwd(dyg, y2, gate, mode):
    """dyg, y2: (B,C,L) contiguous; gate (B,C) fp32 -> (dy2 (B,C,L), dgate (B,C) fp32)."""
    B, C_, L = y2.shape
    assert dyg.is_contiguous() and y2.is_contiguous() and gate.is_contiguous() and gate.dtype == torch.float32
    dy2 = torch.empty_like(y2)
    dg = torch.empty((B, C_), dtype=torch.float32, device=y2.device)
    a = _lib.GateBwdArgs(_ptr(dyg), _ptr(y2), _ptr(gate), _ptr(dy2), _ptr(dg), B, C_, L, mode, _DT[y2.dtype])
    _run("vmb_channel_gate_bwd", a, y2, "gate_bwd")
    return dy2, dg


def pixlin_wgrad(dy, x, per_batch=False, out=None, dbias=None):
    """dW[m,k] = sum_{b,p} dy[b,m,p] x[b,k,p] -> fp32 (M,K)  (or (B,M,K) without the batch sum).  16-bit activations with 16 B aligned
    rows run this library's mma.sync split-pixel kernel; fp32 (parity mode) or unaligned rows go to the library GEMM.
    out: optional fp32 accumulator (the result is ADDED: pre-zero it for a plain result); dbias: optional (M) fp32 accumulator receiving
    sum_{b,p} dy (the bias gradient)."""
    B, M, L = dy.shape
    K = x.shape[1]
    ok = (dy.dtype in (torch.bfloat16, torch.float16) and x.dtype == dy.dtype and dy.stride(2) == 1 and x.stride(2) == 1
          and all(s % 8 == 0 for s in (dy.stride(0), dy.stride(1), x.stride(0), x.stride(1)))
          and dy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0)
    if not ok:
        wb = torch.bmm(dy, x.transpose(1, 2)).float()
        if dbias is not None:
            dbias += dy.float().sum((0, 2))
        res = wb if per_batch else wb.sum(0)
        if out is not None:
            out.add_(res)  # same contract as the kernel: ADD into the (pre-zeroed or accumulating) buffer
            return out
        return res
    if out is None:
        out = torch.zeros((B, M, K) if per_batch else (M, K), dtype=torch.float32, device=dy.device)
    a = _lib.WgradArgs(_ptr(dy), _ptr(x), _ptr(out), B, M, K, L, dy.stride(0), dy.stride(1), x.stride(0), x.stride(1),
                       int(per_batch), _DT[dy.dtype], _ptr(dbias))
    _run("vmb_pixlin_wgrad", a, dy, "wgrad")
    return out


def prep_block_weights(jobs, dtype):
    """jobs: list of (type, src, src2, dst, dst2, M, K, N2, ld, ld2) -- see vmb_prep_block_weights; one launch."""
    a = _lib.PrepArgs()
    a.njobs, a.dtype = len(jobs), _DT[dtype]
    for i, (t, src, src2, dst, dst2, M, K, N2, ld, ld2) in enumerate(jobs):
        a.jobs[i] = _lib.PrepJob(src.data_ptr(), src2.data_ptr() if src2 is not None else None, dst.data_ptr(),
                                 dst2.data_ptr() if dst2 is not None else None, t, M, K, N2, ld, ld2)
    _run("vmb_prep_block_weights", a, jobs[0][1], "prep")


def sum4_add(x4, add):
    """x4 (B,4,C,L) contiguous, add (B,C,L) contiguous -> add + x4.sum(1) (fp32 accumulation, result in the input dtype)"""
    B = x4.shape[0]
    assert x4.is_contiguous() and add.is_contiguous() and x4.dtype == add.dtype
    out = torch.empty_like(add)
    global _LAUNCHES
    with torch.cuda.device(x4.device):
        _lib.check(_lib.lib().vmb_sum4_add(x4.data_ptr(), add.data_ptr(), out.data_ptr(), B, add[0].numel(), _DT[x4.dtype],
                                           torch.cuda.current_stream(x4.device).cuda_stream), "sum4_add")
    _LAUNCHES += 1
    return out


def _channel_args(pooled, inv_count, prm, C_, c_out):
    B = pooled.shape[0]
    return _lib.ChannelArgs(_ptr(pooled), inv_count, _ptr(prm["cin_w"]), _ptr(prm["cin_b"]), _ptr(prm["xc_proj"]),
                            _ptr(prm["dtc_w"]), _ptr(prm["dtc_b"]), _ptr(prm["Ac_logs"]), _ptr(prm["Dsc"]), _ptr(prm["cout_w"]),
                            _ptr(prm["cout_b"]), _ptr(prm["cn_w"]), _ptr(prm["cn_b"]), _ptr(c_out), B, C_, prm["dc"], prm["Rc"], prm["N"])


def channel_branch_bwd_supported(C_, prm) -> bool:
    return prm["N"] <= 16 and 2 * prm["dc"] * 16 <= 512 and \
        _lib.lib().vmb_channel_branch_bwd_smem_bytes(C_, prm["dc"], prm["Rc"], prm["N"]) <= 227 * 1024


def channel_branch_bwd(pooled, inv_count, prm, C_, dc_out, into=None):
    """backward of channel_branch: -> (dpooled (B,C) fp32, dict of fp32 parameter gradients keyed like prm).
    into: optional {key: contiguous fp32 accumulator} -- those gradients are ADDED there (e.g. the parameters' .grad buffers)"""
    B = pooled.shape[0]
    dev = pooled.device
    dc, Rc, N = prm["dc"], prm["Rc"], prm["N"]
    RN = Rc + 2 * N
    sizes = dict(cin_w=dc, cin_b=dc, xc_proj=2 * RN * dc, dtc_w=2 * dc * Rc, dtc_b=2 * dc, Ac_logs=2 * dc * N, Dsc=2 * dc, cout_w=dc,
                 cout_b=1, cn_w=C_, cn_b=C_)
    into = into or {}
    own = {k: n for k, n in sizes.items() if k not in into}
    g, o = {}, 0
    if own:
        flat = torch.zeros(sum(own.values()), dtype=torch.float32, device=dev)  # one zero fill for every accumulator
        for k, n in own.items():
            g[k] = flat[o:o + n]
            o += n
    for k, t in into.items():
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == sizes[k]
        g[k] = t
    has_cin, has_cout = prm["cin_w"] is not None, prm["cout_w"] is not None
    dpooled = torch.empty((B, C_), dtype=torch.float32, device=dev)
    scratch = torch.empty(_lib.lib().vmb_channel_branch_bwd_scratch_bytes(B, dc, C_) // 4, dtype=torch.float32, device=dev)
    dc_out = dc_out.float().contiguous()
    a = _lib.ChannelBwdArgs(_channel_args(pooled, inv_count, prm, C_, None), _ptr(dc_out), _ptr(dpooled),
                            _ptr(g["cin_w"]) if has_cin else None, _ptr(g["cin_b"]) if has_cin else None, _ptr(g["xc_proj"]),
                            _ptr(g["dtc_w"]), _ptr(g["dtc_b"]), _ptr(g["Ac_logs"]), _ptr(g["Dsc"]),
                            _ptr(g["cout_w"]) if has_cout else None, _ptr(g["cout_b"]) if has_cout else None, _ptr(g["cn_w"]),
                            _ptr(g["cn_b"]), _ptr(scratch))
    _run("vmb_channel_branch_bwd", a, pooled, "channel_bwd")
    return dpooled, g
