"""GPU parity of the selective-scan operator (through the C-ABI) against the CPU oracle and the
golden vectors produced by the reference's selective_scan_ref.
Tolerances: fp32 1e-3 rel / 1e-5 abs is BASELINE.json's bar, checked against the fp64 oracle where
fp32 reassociation noise of two sequential evaluations already exceeds 1e-5 abs (see test_oracle.py);
16-bit I/O uses the reference test's own tolerances (test_selective_scan.py:398-400)."""
import os

import numpy as np
import pytest
import torch

from oracle import cscan

pytestmark = pytest.mark.gpu

TOL = {torch.float32: (1e-3, 1e-5), torch.float16: (3e-3, 5e-3), torch.bfloat16: (3e-2, 5e-2)}


def make_inputs(b, D, L, N, G, dtype, seed=0, has_D=True, has_bias=True, model_like=False):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    rn = lambda *s: torch.randn(*s, generator=g)
    if model_like:  # A = -(1..N), small dt -- what the OSS block feeds the operator at init
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1)
        delta = 0.1 * rn(b, D, L) - 4.0
    else:  # distributions of the reference test (test_selective_scan.py:406-441)
        A = -0.5 * r(D, N)
        delta = 0.5 * r(b, D, L)
    Bm, Cm, u = rn(b, G, N, L), rn(b, G, N, L), rn(b, D, L)
    Dv = rn(D) if has_D else None
    bias = 0.5 * r(D) if has_bias else None
    cast = lambda t: t.to(dtype)
    return cast(u), cast(delta), A, cast(Bm), cast(Cm), Dv, bias


def run_ours(u, delta, A, Bm, Cm, Dv, bias, softplus, need_ckpt=True):
    from vmambair_b200 import ops
    dev = "cuda"
    mv = lambda t: None if t is None else t.to(dev)
    out, ckpt = ops.selective_scan_fwd(mv(u), mv(delta), mv(A), mv(Bm), mv(Cm), mv(Dv), mv(bias), softplus, need_ckpt)
    torch.cuda.synchronize()
    return out, ckpt


def assert_close_ref(out, ref64, dtype, ref32=None):
    rtol, atol = TOL[dtype]
    out = out.double().cpu()
    err = (out - ref64).abs()
    bound = atol + rtol * ref64.abs()
    if dtype == torch.float32 and ref32 is not None:
        # fp32 noise floor: the reference's own fp32 sequential evaluation is off the fp64 truth by
        # E = max|ref32 - ref64| (>> 1e-5 abs on slowly-decaying states at L=4096).  Our kernel, like the
        # reference CUDA kernel, evaluates exp with MUFU.EX2 (2 ulp vs libm's 0.5 ulp), so we allow 8*E on top
        # of BASELINE's 1e-3 rel / 1e-5 abs.  test_fwd_model_like_distribution holds the strict bound.
        bound = bound + 8.0 * (ref32.double() - ref64).abs().max()
    bad = err > bound
    assert not bad.any(), f"{int(bad.sum())} / {bad.numel()} out of tolerance; max err {err.max():.3e}"


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_golden_vectors(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "scan_cases.npz"))
    g = lambda k: torch.from_numpy(z[f"{name}/{k}"]) if f"{name}/{k}" in z else None
    sp = bool(z[f"{name}/softplus"])
    out, _ = run_ours(g("u"), g("delta"), g("A"), g("B"), g("C"), g("D"), g("bias"), sp)
    torch.testing.assert_close(out.cpu(), g("out"), rtol=1e-3, atol=1e-5)


SHAPES = [  # b, D, L, N, G
    (2, 32, 64, 16, 4),      # RB=8
    (1, 64, 1000, 16, 2),    # RB=32/16, ragged L
    (2, 192, 4096, 16, 4),   # C=48 OSS scan
    (1, 384, 4096, 16, 4),   # C=96 OSS scan (north-star shape)
    (3, 8, 96, 16, 2),       # channel scan shape (dc_inner=4, L=C)
    (2, 2, 48, 16, 2),       # RealSR channel scan (1 row per group)
    (1, 24, 777, 1, 1),      # dstate=1 (reference test parametrisation)
    (1, 16, 300, 40, 2),     # dstate > 16: several state tiles
    (1, 6, 513, 16, 3),      # 2 rows per group, odd everything
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("softplus", [True, False])
def test_fwd_vs_oracle(shape, dtype, softplus):
    b, D, L, N, G = shape
    ins = make_inputs(b, D, L, N, G, dtype, seed=L)
    out, ckpt = run_ours(*ins, softplus)
    ref64 = cscan.scan_fwd(*ins, softplus, fp64=True)
    ref32 = cscan.scan_fwd(*ins, softplus)
    assert_close_ref(out, ref64, dtype, ref32)


@pytest.mark.parametrize("flags", [(False, False), (True, False), (False, True)])
def test_fwd_optional_args(flags):
    has_D, has_bias = flags
    ins = make_inputs(2, 32, 200, 16, 2, torch.float32, seed=5, has_D=has_D, has_bias=has_bias)
    out, _ = run_ours(*ins, True)
    ref64 = cscan.scan_fwd(*ins, True, fp64=True)
    assert_close_ref(out, ref64, torch.float32, cscan.scan_fwd(*ins, True))


def test_fwd_model_like_distribution():
    """dt ~ softplus(-4) ~ 0.02, A = -(1..16): the operating point of the OSS block.  Strict 1e-3 / 1e-5."""
    ins = make_inputs(2, 192, 4096, 16, 4, torch.float32, seed=9, model_like=True)
    out, _ = run_ours(*ins, True)
    ref64 = cscan.scan_fwd(*ins, True, fp64=True)
    assert_close_ref(out, ref64, torch.float32, None)


def test_fwd_long_sequence_config4():
    """L = 65536 (BASELINE config 4 scan stress), one batch, C=48."""
    ins = make_inputs(1, 192, 65536, 16, 4, torch.bfloat16, seed=3, model_like=True)
    out, _ = run_ours(*ins, True, need_ckpt=False)
    ref64 = cscan.scan_fwd(*ins, True, fp64=True)
    assert_close_ref(out, ref64, torch.bfloat16)


def test_fwd_noncontiguous_rows():
    """u/delta given as strided views (batch/dim strides != dense) -- the reference accepts these (cpp:180-181)."""
    b, D, L, N, G = 2, 16, 128, 16, 2
    u, delta, A, Bm, Cm, Dv, bias = make_inputs(b, D, L, N, G, torch.float32, seed=1)
    big_u = torch.zeros(b, D, 2 * L); big_u[:, :, :L] = u
    from vmambair_b200 import ops
    dev = "cuda"
    uu = big_u.to(dev)[:, :, :L]
    out, _ = ops.selective_scan_fwd(uu, delta.to(dev), A.to(dev), Bm.to(dev), Cm.to(dev), Dv.to(dev), bias.to(dev), True)
    ref64 = cscan.scan_fwd(u, delta, A, Bm, Cm, Dv, bias, True, fp64=True)
    assert_close_ref(out, ref64, torch.float32, cscan.scan_fwd(u, delta, A, Bm, Cm, Dv, bias, True))


def test_linearity_in_u_full_size():
    """size-independent property at BASELINE size: scan(u1+u2) == scan(u1)+scan(u2) (D=0)."""
    u, delta, A, Bm, Cm, _, bias = make_inputs(8, 384, 4096, 16, 4, torch.float32, seed=2, has_D=False)
    u2 = torch.randn_like(u)
    o1, _ = run_ours(u, delta, A, Bm, Cm, None, bias, True)
    o2, _ = run_ours(u2, delta, A, Bm, Cm, None, bias, True)
    o12, _ = run_ours(u + u2, delta, A, Bm, Cm, None, bias, True)
    torch.testing.assert_close(o12, o1 + o2, rtol=1e-4, atol=1e-3)


def test_errors_raise():
    from vmambair_b200 import ops
    u, delta, A, Bm, Cm, Dv, bias = [t.cuda() for t in make_inputs(1, 8, 32, 16, 2, torch.float32)]
    with pytest.raises(RuntimeError):
        ops.selective_scan_fwd(u, delta.half(), A, Bm, Cm, Dv, bias, True)
    with pytest.raises(RuntimeError):
        ops.selective_scan_fwd(u, delta, A, Bm[:, :, :8], Cm, Dv, bias, True)
    with pytest.raises(RuntimeError):
        ops.selective_scan_fwd(u.double(), delta.double(), A, Bm.double(), Cm.double(), Dv, bias, True)


# ----------------------------------------------------------------------------- backward
BWD_TOL = {torch.float32: 2e-3, torch.float16: 6e-3, torch.bfloat16: 4e-2}


def run_ours_bwd(u, delta, A, Bm, Cm, Dv, bias, dout, softplus):
    from vmambair_b200 import ops
    mv = lambda t: None if t is None else t.cuda()
    a = [mv(t) for t in (u, delta, A, Bm, Cm, Dv, bias)]
    out, ckpt = ops.selective_scan_fwd(*a, softplus, True)
    grads = ops.selective_scan_bwd(*a, mv(dout), ckpt, softplus)
    torch.cuda.synchronize()
    return out, grads


def check_grads(got, ref, dtype, names=("du", "ddelta", "dA", "dB", "dC", "dD", "dbias")):
    tol = BWD_TOL[dtype]
    for n, g, r in zip(names, got, ref):
        if r is None:
            assert g is None, n
            continue
        g = g.double().cpu()
        r = r.double()
        scale = r.abs().max().clamp_min(1e-6)
        err = (g - r).abs().max()
        # scale-aware bound (reference test uses rtol up to 5x / atol up to 10x of its output tolerance, :490-502)
        assert err <= tol * scale + 1e-5, f"{n}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_bwd_golden_vectors(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "scan_cases.npz"))
    g = lambda k: torch.from_numpy(z[f"{name}/{k}"]) if f"{name}/{k}" in z else None
    sp = bool(z[f"{name}/softplus"])
    _, grads = run_ours_bwd(g("u"), g("delta"), g("A"), g("B"), g("C"), g("D"), g("bias"), g("dout"), sp)
    ref = [g(k) for k in ("du", "ddelta", "dA", "dB", "dC", "dD", "dbias")]
    check_grads(grads, ref, torch.float32)


BWD_SHAPES = [
    (2, 32, 64, 16, 4),
    (1, 64, 1000, 16, 2),
    (1, 192, 2048, 16, 4),
    (3, 8, 96, 16, 2),
    (2, 2, 48, 16, 2),
    (1, 24, 300, 1, 1),
    (1, 16, 200, 40, 2),
    (1, 6, 513, 16, 3),
]


@pytest.mark.parametrize("shape", BWD_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("softplus", [True, False])
def test_bwd_vs_oracle(shape, dtype, softplus):
    b, D, L, N, G = shape
    ins = make_inputs(b, D, L, N, G, dtype, seed=L + 1)
    dout = torch.randn(b, D, L, generator=torch.Generator().manual_seed(L)).to(dtype)
    _, grads = run_ours_bwd(*ins, dout, softplus)
    ref = cscan.scan_bwd(*ins, dout, softplus)
    check_grads(grads, ref, dtype)


def test_bwd_optional_args_none():
    ins = make_inputs(2, 16, 130, 16, 2, torch.float32, seed=4, has_D=False, has_bias=False)
    dout = torch.randn(2, 16, 130)
    _, grads = run_ours_bwd(*ins, dout, True)
    ref = cscan.scan_bwd(*ins, dout, True)
    assert grads[5] is None and grads[6] is None
    check_grads(grads, ref, torch.float32)


def test_bwd_model_like_full_row():
    """C=48 OSS scan shape, model-like distribution, fp32."""
    ins = make_inputs(1, 192, 4096, 16, 4, torch.float32, seed=11, model_like=True)
    dout = torch.randn(1, 192, 4096, generator=torch.Generator().manual_seed(1))
    _, grads = run_ours_bwd(*ins, dout, True)
    ref = cscan.scan_bwd(*ins, dout, True)
    check_grads(grads, ref, torch.float32)
