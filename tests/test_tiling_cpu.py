"""Host logic of padded / tiled inference (vmambair_b200.tiling) against a line-by-line restatement of the reference loops
(RealSR/VmambaIR/utils.py:97-160 tile_process; SRGAN/VmambaIR/models/MambaSISR_model.py:87-118 pad_test)."""
import math

import pytest
import torch
import torch.nn.functional as F

from vmambair_b200 import tiling


def _net(scale):
    """shape-preserving stand-in with a receptive field (3x3 blur) and x`scale` upsampling: tile borders matter"""
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16
    def f(x):
        c = x.shape[1]
        y = F.conv2d(F.pad(x, (1, 1, 1, 1), "replicate"), k.view(1, 1, 3, 3).repeat(c, 1, 1, 1), groups=c)
        return F.interpolate(y, scale_factor=scale, mode="nearest")
    return f


def _reference_tile_loop(model, img, tile_size, tile_pad, scale):
    """the reference's loop, tile by tile (RealESRGANer.tile_process)"""
    batch, channel, height, width = img.shape
    output = img.new_zeros((batch, channel, height * scale, width * scale))
    tiles_x, tiles_y = math.ceil(width / tile_size), math.ceil(height / tile_size)
    for y in range(tiles_y):
        for x in range(tiles_x):
            ofs_x, ofs_y = x * tile_size, y * tile_size
            isx, iex = ofs_x, min(ofs_x + tile_size, width)
            isy, iey = ofs_y, min(ofs_y + tile_size, height)
            isxp, iexp = max(isx - tile_pad, 0), min(iex + tile_pad, width)
            isyp, ieyp = max(isy - tile_pad, 0), min(iey + tile_pad, height)
            tw, th = iex - isx, iey - isy
            out_tile = model(img[:, :, isyp:ieyp, isxp:iexp])
            osxt, osyt = (isx - isxp) * scale, (isy - isyp) * scale
            output[:, :, isy * scale:iey * scale, isx * scale:iex * scale] = out_tile[:, :, osyt:osyt + th * scale, osxt:osxt + tw * scale]
    return output


@pytest.mark.parametrize("hw,tile,pad,scale,mb", [((72, 88), 32, 8, 4, 3), ((64, 64), 64, 10, 4, 8), ((50, 97), 24, 5, 2, 2), ((40, 40), 16, 0, 1, 8)])
def test_tile_process_matches_reference_loop(hw, tile, pad, scale, mb):
    torch.manual_seed(0)
    img = torch.rand(2, 3, *hw)
    got = tiling.tile_process(_net(scale), img, tile, pad, scale, max_batch=mb)
    ref = _reference_tile_loop(_net(scale), img, tile, pad, scale)
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=0, atol=0)


@pytest.mark.parametrize("hw,m", [((61, 70), 8), ((64, 64), 8), ((33, 47), 16)])
def test_pad_to_multiple_is_the_reference_pad(hw, m):
    x = torch.rand(1, 3, *hw)
    xp, ph, pw = tiling.pad_to_multiple(x, m)
    h, w = hw
    mod_pad_h = m - h % m if h % m != 0 else 0   # MambaSISR_model.py:91-95
    mod_pad_w = m - w % m if w % m != 0 else 0
    assert (ph, pw) == (mod_pad_h, mod_pad_w)
    torch.testing.assert_close(xp, F.pad(x, (0, mod_pad_w, 0, mod_pad_h), "reflect"), rtol=0, atol=0)
    y = tiling.padded_inference(_net(4), x, m, 4)
    assert y.shape == (1, 3, h * 4, w * 4)
    torch.testing.assert_close(y, _net(4)(xp)[:, :, :h * 4, :w * 4], rtol=0, atol=0)
