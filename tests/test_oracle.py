"""Pins the oracle (oracle/) against vectors produced by the REAL reference function
(tests/golden/make_golden.py ran selective_scan_ref from
Mamba/kernels/selective_scan/test_selective_scan.py:168-234 in the build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import cscan
from oracle.selective_scan_ref import selective_scan_oracle, selective_scan_oracle_bwd

CASES = ["a", "b", "c", "d", "e"]


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "scan_cases.npz"))
    g = lambda k: torch.from_numpy(z[f"{name}/{k}"]) if f"{name}/{k}" in z else None
    return {k: g(k) for k in ("u", "delta", "A", "B", "C", "D", "bias", "out", "last_state", "dout",
                              "du", "ddelta", "dA", "dB", "dC", "dD", "dbias")} | {
        "softplus": bool(z[f"{name}/softplus"])}


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_matches_reference(golden_dir, name):
    c = load_case(golden_dir, name)
    out, last = selective_scan_oracle(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"],
                                      c["softplus"], return_last_state=True)
    torch.testing.assert_close(out, c["out"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(last, c["last_state"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference(golden_dir, name):
    c = load_case(golden_dir, name)
    out, last = cscan.scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"],
                               c["softplus"], return_last_state=True)
    torch.testing.assert_close(out, c["out"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(last, c["last_state"], rtol=1e-5, atol=1e-5)
    out64 = cscan.scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"],
                           c["softplus"], fp64=True)
    torch.testing.assert_close(out64.float(), c["out"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("impl", ["torch", "c"])
def test_bwd_oracle_matches_reference_autograd(golden_dir, name, impl):
    c = load_case(golden_dir, name)
    fn = selective_scan_oracle_bwd if impl == "torch" else cscan.scan_bwd
    got = fn(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["bias"], c["dout"], c["softplus"])
    names = ["du", "ddelta", "dA", "dB", "dC", "dD", "dbias"]
    for n, g in zip(names, got):
        if c[n] is None:
            assert g is None
            continue
        torch.testing.assert_close(g.float(), c[n], rtol=2e-4, atol=2e-4, msg=lambda m: f"{n}: {m}")


def test_c_oracle_groups_and_sizes():
    """C oracle == torch oracle on a bigger ragged case (self-consistency of the two restatements)."""
    torch.manual_seed(0)
    b, D, L, N, G = 2, 24, 777, 16, 4
    u = torch.randn(b, D, L); dl = 0.5 * torch.rand(b, D, L); A = -0.5 * torch.rand(D, N)
    B = torch.randn(b, G, N, L); C = torch.randn(b, G, N, L); Dv = torch.randn(D); bias = 0.5 * torch.rand(D)
    o1 = selective_scan_oracle(u, dl, A, B, C, Dv, bias, True)
    o2 = cscan.scan_fwd(u, dl, A, B, C, Dv, bias, True)
    # two fp32 sequential evaluations (different sum order over n, libm vs SLEEF exp) already differ by ~7e-5 abs at L=777
    torch.testing.assert_close(o1, o2, rtol=1e-4, atol=2e-4)
