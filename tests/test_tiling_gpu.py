"""Tiled / padded inference on the fused networks (vmambair_b200.tiling.TiledInference) against the reference's loops driven by
the CPU oracle network (RealSR/VmambaIR/utils.py:97-160; Deraining/basicsr/test_deraining.py:73-85)."""
import pytest
import torch
import torch.nn.functional as F

import vmambair_b200.archs as archs
from oracle import oss_ref
from tests.test_tiling_cpu import _reference_tile_loop
from vmambair_b200.tiling import TiledInference

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def test_tiled_realsr_matches_reference_tile_loop_on_the_oracle():
    torch.manual_seed(2)
    net = archs.MambaRealSR11(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    img = torch.rand(1, 3, 72, 88)
    model = lambda t: oss_ref.net_forward(sd, t, "realsr")
    ref = _reference_tile_loop(model, img, 48, 8, 4)
    got = TiledInference(net, scale=4, multiple=8, tile_size=48, tile_pad=8, dtype=torch.float32, max_batch=2)(img).cpu()
    assert got.shape == ref.shape == (1, 3, 288, 352)
    torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-4)


def test_padded_deraining_matches_reference_padding_on_the_oracle():
    """61 x 70 input: reflect-padded to 64 x 72, restored, cropped (test_deraining.py:73-85)"""
    torch.manual_seed(3)
    net = archs.Mamber32(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    img = torch.rand(2, 3, 61, 70)
    h, w, factor = 61, 70, 8
    Hp, Wp = ((h + factor) // factor) * factor, ((w + factor) // factor) * factor
    padh, padw = (Hp - h if h % factor != 0 else 0), (Wp - w if w % factor != 0 else 0)
    ref = oss_ref.net_forward(sd, F.pad(img, (0, padw, 0, padh), "reflect"), "mamber32")[:, :, :h, :w]
    got = TiledInference(net, scale=1, multiple=8, dtype=torch.float32)(img).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-4)
