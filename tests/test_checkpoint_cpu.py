"""CPU tests of the BasicSR checkpoint conventions (vmambair_b200.checkpoint; reference base_model.py:213-244,281-309)."""
import torch
import pytest

from vmambair_b200 import archs, checkpoint


def tiny():
    torch.manual_seed(0)
    return archs.MambaSISR6(dim=8, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1)


def test_reference_style_file_loads_strict(tmp_path):
    src = tiny()
    with torch.no_grad():
        for p in src.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    # what the reference writes after DataParallel training: both copies, "module." prefixes on one of them
    ref_file = {"params": {"module." + k: v.clone() for k, v in src.state_dict().items()},
                "params_ema": {k: v.clone() * 0.5 for k, v in src.state_dict().items()}}
    path = tmp_path / "net_g_latest.pth"
    torch.save(ref_file, path)
    a = checkpoint.load_network(tiny(), str(path), strict=True, param_key="params")
    b = checkpoint.load_network(tiny(), str(path), strict=True, param_key="params_ema")
    for k, v in src.state_dict().items():
        assert torch.equal(a.state_dict()[k], v) and torch.equal(b.state_dict()[k], v * 0.5)


def test_params_ema_falls_back_to_params(tmp_path):
    src = tiny()
    path = tmp_path / "only_params.pth"
    checkpoint.save_network(src, str(path))
    got = checkpoint.load_network(tiny(), str(path), param_key="params_ema")
    assert all(torch.equal(got.state_dict()[k], v) for k, v in src.state_dict().items())
    with pytest.raises(KeyError):
        checkpoint.select_params({"state": {}}, "params_ema")


def test_save_roundtrip_and_layout(tmp_path):
    net, ema = tiny(), tiny()
    path = tmp_path / "pair.pth"
    checkpoint.save_network(torch.nn.DataParallel(net), str(path), ema=ema)
    raw = torch.load(path, weights_only=True)
    assert set(raw) == {"params", "params_ema"}
    assert list(raw["params"]) == list(net.state_dict()) and not any(k.startswith("module.") for k in raw["params"])


def test_non_strict_skips_size_mismatch(tmp_path):
    src = tiny()
    state = {k: v.clone() for k, v in src.state_dict().items()}
    k0 = "patch_embed.proj.weight"
    state[k0] = torch.zeros(3, 3)
    path = tmp_path / "mismatch.pth"
    torch.save({"params": state}, path)
    with pytest.raises(RuntimeError):
        checkpoint.load_network(tiny(), str(path), strict=True)
    dst = tiny()
    before = dst.state_dict()[k0].clone()
    checkpoint.load_network(dst, str(path), strict=False)
    assert torch.equal(dst.state_dict()[k0], before)
    k1 = "refinement.0.attn.A_logs"
    assert torch.equal(dst.state_dict()[k1], src.state_dict()[k1])
