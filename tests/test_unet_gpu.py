"""vmb_conv3x3 and the native U-Net path (vmambair_b200/unet.py) against plain PyTorch fp32 references of the
same ops: F.conv2d + F.pixel_unshuffle / F.pixel_shuffle / torch.cat / F.interpolate, i.e. what the reference modules run
(SRGAN/VmambaIR/archs/MambaSISR6_arch.py:520-553,590-607,640), and against the CPU oracle network."""
import pytest
import torch
import torch.nn.functional as F

import vmambair_b200.archs as archs
from oracle import oss_ref
from vmambair_b200 import ops, unet

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"

TOL = {torch.float32: (1e-4, 1e-5), torch.bfloat16: (1.6e-2, 2e-3), torch.float16: (2e-3, 3e-4)}


def _ref(x, w, b, mode, add=None, s=1):
    y = F.conv2d(x.float(), w.float(), None if b is None else b.float(), padding=1)
    if mode == ops.CONV_UNSHUFFLE2:
        return F.pixel_unshuffle(y, 2)
    if mode == ops.CONV_SHUFFLE2:
        return F.pixel_shuffle(y, 2)
    if mode == ops.CONV_ADD_NEAREST:
        return y + F.interpolate(add.float(), scale_factor=s, mode="nearest")
    return y


def _check(got, ref, dtype, scale=1.0):
    rtol, atol = TOL[dtype]
    torch.testing.assert_close(got.float(), ref, rtol=rtol, atol=atol * scale)


# (Cin, Cout, H, W, B): the conv sites of the nets at the benchmark geometry + ragged / tiny cases
SITES = [(3, 48, 64, 64, 2), (48, 24, 64, 64, 2), (96, 48, 32, 32, 2), (192, 96, 16, 16, 3), (384, 768, 8, 8, 2),
         (192, 384, 16, 16, 2), (96, 192, 32, 32, 2), (16, 8, 10, 22, 1), (20, 36, 24, 40, 2), (5, 7, 9, 13, 1), (32, 12, 8, 8, 1)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("site", SITES)
def test_conv3x3_plain_and_shuffles(site, dtype):
    Cin, Cout, H, W, B = site
    torch.manual_seed(Cin * 131 + Cout)
    x = torch.randn(B, Cin, H, W, device=DEV).to(dtype)
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV) / (3.0 * Cin ** 0.5)).to(dtype)
    b = torch.randn(Cout, device=DEV)
    wp = ops.pack_conv3x3_weight(w, dtype)
    mag = float(_ref(x, w, b, ops.CONV_PLAIN).abs().max())
    for bias in (b, None):
        _check(ops.conv3x3(x, wp, bias, Cout), _ref(x, w, bias, ops.CONV_PLAIN), dtype, mag)
    if H % 2 == 0 and W % 2 == 0:
        _check(ops.conv3x3(x, wp, None, Cout, ops.CONV_UNSHUFFLE2), _ref(x, w, None, ops.CONV_UNSHUFFLE2), dtype, mag)
    if Cout % 4 == 0:
        _check(ops.conv3x3(x, wp, b, Cout, ops.CONV_SHUFFLE2), _ref(x, w, b, ops.CONV_SHUFFLE2), dtype, mag)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_conv3x3_strided_views_in_and_out(dtype):
    """input = a channel slice of a wider tensor (an encoder output living in the concatenation buffer), output = the other slice"""
    torch.manual_seed(5)
    B, Cin, Cout, H, W = 2, 96, 192, 32, 32
    wide = torch.randn(B, 2 * Cin, H, W, device=DEV).to(dtype)
    x = wide[:, Cin:]
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV) / 30).to(dtype)
    wp = ops.pack_conv3x3_weight(w, dtype)
    cat = torch.full((B, 2 * (Cout // 4), 2 * H, 2 * W), 7.0, device=DEV, dtype=dtype)
    ops.conv3x3(x, wp, None, Cout, ops.CONV_SHUFFLE2, out=cat[:, :Cout // 4])
    _check(cat[:, :Cout // 4], _ref(x, w, None, ops.CONV_SHUFFLE2), dtype, 2.0)
    assert bool((cat[:, Cout // 4:] == 7.0).all()), "the other half of the concatenation buffer was touched"
    big = torch.zeros(B, 4 * Cout + 8, H // 2, W // 2, device=DEV, dtype=dtype)
    ops.conv3x3(x, wp, None, Cout, ops.CONV_UNSHUFFLE2, out=big[:, 8:])
    _check(big[:, 8:], _ref(x, w, None, ops.CONV_UNSHUFFLE2), dtype, 2.0)
    assert bool((big[:, :8] == 0).all())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("case", [(96, 3, 64, 64, 2, 4, True), (96, 3, 40, 72, 1, 4, True), (96, 3, 32, 32, 2, 1, False),
                                  (32, 3, 24, 24, 2, 2, True), (16, 5, 12, 20, 1, 1, True)])
def test_conv3x3_last_conv_add_nearest(case, dtype):
    """SR tail: conv_last on the NHWC tensor + F.interpolate(inp_img, scale, 'nearest') -> NCHW; Mamber32: output conv + inp_img"""
    Cin, Cout, H, W, B, s, nhwc = case
    torch.manual_seed(11)
    x = torch.randn(B, Cin, H, W, device=DEV).to(dtype)
    if nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV) / (3.0 * Cin ** 0.5)).to(dtype)
    b = torch.randn(Cout, device=DEV)
    img = torch.rand(B, Cout, H // s, W // s, device=DEV).to(dtype)
    got = ops.conv3x3(x, ops.pack_conv3x3_weight(w, dtype), b, Cout, ops.CONV_ADD_NEAREST, add=img, add_scale=s, nhwc=nhwc)
    assert got.is_contiguous() and got.shape == (B, Cout, H, W)
    _check(got, _ref(x, w, b, ops.CONV_ADD_NEAREST, img, s), dtype, 4.0)


def test_conv3x3_argument_errors():
    x = torch.randn(1, 8, 9, 9, device=DEV)
    wp = ops.pack_conv3x3_weight(torch.randn(8, 8, 3, 3, device=DEV), torch.float32)
    with pytest.raises(RuntimeError, match="PixelUnshuffle"):
        ops.conv3x3(x, wp, None, 8, ops.CONV_UNSHUFFLE2, out=torch.empty(1, 32, 4, 4, device=DEV))
    with pytest.raises(RuntimeError, match="PixelShuffle"):
        ops.conv3x3(x, ops.pack_conv3x3_weight(torch.randn(6, 8, 3, 3, device=DEV), torch.float32), None, 6, ops.CONV_SHUFFLE2,
                    out=torch.empty(1, 1, 18, 18, device=DEV))


def _both_paths(net, x):
    with torch.no_grad():
        unet.set_mode("torch")
        try:
            ref = net(x)
        finally:
            unet.set_mode("native")
        n0 = ops.launch_count()
        got = net(x)
        return got, ref, ops.launch_count() - n0


@pytest.mark.parametrize("name", ["sisr", "realsr", "mamber32"])
def test_native_unet_matches_torch_unet_fp32(name):
    """same OSS blocks, the non-OSS stages on vmb_conv3x3 / vmb_pixlin / in-place concatenation vs nn.Conv2d / torch.cat"""
    torch.manual_seed(21)
    cls = {"sisr": archs.MambaSISR6, "realsr": archs.MambaRealSR11, "mamber32": archs.Mamber32}[name]
    net = cls(dim=16, num_blocks=[2, 1, 1, 1], num_refinement_blocks=1).to(DEV).eval()
    x = torch.rand(2, 3, 32, 48, device=DEV)
    got, ref, _ = _both_paths(net, x)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-5)


def test_native_unet_vs_oracle_fp32_and_launch_census():
    """light SR net at a reduced width against the CPU oracle network; every non-OSS stage is a library launch:
    7 trunk convs + 2 reduce convs + conv_last on top of the blocks' kernels"""
    torch.manual_seed(22)
    net = archs.MambaSISR6(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(1, 3, 32, 32)
    ref = oss_ref.net_forward(sd, x, "sisr")
    net = net.to(DEV)
    got, tref, launches = _both_paths(net, x.to(DEV))
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-3, atol=1e-4)
    with torch.no_grad():
        unet.set_mode("torch")
        n0 = ops.launch_count()
        net(x.to(DEV))
        unet.set_mode("native")
    assert launches - (ops.launch_count() - n0) == 10


def test_native_unet_bf16_close_to_torch_unet_bf16():
    torch.manual_seed(23)
    net = archs.MambaSISR6(dim=16, num_blocks=[2, 1, 1, 1], num_refinement_blocks=2).to(DEV).eval().to(torch.bfloat16)
    for n, p in net.named_parameters():
        if n.rsplit(".", 1)[-1] in ("A_logs", "Ac_logs", "Ds", "Dsc", "dt_projs_bias", "dtc_projs_bias"):
            p.data = p.data.float()
    x = torch.rand(2, 3, 64, 64, device=DEV).to(torch.bfloat16)
    got, ref, _ = _both_paths(net, x)
    d = (got.float() - ref.float()).abs()
    assert d.mean() < 6e-3 and d.max() < 0.08, (d.mean().item(), d.max().item())
