"""optim.FlatAdam (vmb_fused_adam: gradient scale + global-norm clip + Adam / AdamW + EMA + gradient reset in one kernel) against
torch.optim.Adam / AdamW + clip_grad_norm_ + the reference's EMA update (SRGAN/VmambaIR/models/MambaSISR_model.py:141-147,
Deraining/basicsr/models/image_restoration_model.py:165-173, base_model.py:54-62)."""
import pytest
import torch

from vmambair_b200.optim import FlatAdam

pytestmark = pytest.mark.gpu



@pytest.fixture(autouse=True)
def _restore_grad_routing():
    yield
    from vmambair_b200 import fused_train
    fused_train.set_direct_grads(False)  # FlatAdam switches the direct .grad accumulation on; other test modules expect autograd's route


SHAPES = [(48, 48, 1, 1), (96,), (4, 35, 48), (192, 16), (7,), (254, 1, 3, 3), (1,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in SHAPES]


@pytest.mark.parametrize("wd,decoupled", [(0.0, False), (1e-2, False), (1e-2, True)])
@pytest.mark.parametrize("max_norm", [0.0, 0.01, 5.0])
def test_flat_adam_matches_torch(wd, decoupled, max_norm):
    ours, ref = _params(0), _params(0)
    ema_ref = [p.detach().clone() for p in ref]
    lr, betas, eps, decay, scale = 2e-4, (0.9, 0.99), 1e-8, 0.999, 0.5
    opt = FlatAdam(ours, lr=lr, betas=betas, eps=eps, weight_decay=wd, decoupled=decoupled, ema_decay=decay, max_grad_norm=max_norm)
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    topt = cls(ref, lr=lr, betas=betas, eps=eps, weight_decay=wd)
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        for po, pr in zip(ours, ref):
            gr = torch.randn(po.shape, generator=g).cuda() * (10.0 if step == 2 else 0.1)
            po.grad.copy_(gr)           # the .grad views of the flat buffer
            pr.grad = gr * scale        # grad_scale (1 / world after the all-reduce) applied before clipping
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_(ref, max_norm)
        topt.step()
        for e, p in zip(ema_ref, ref):
            e.mul_(decay).add_(p.detach(), alpha=1 - decay)
        opt.step(grad_scale=scale, zero_grad=True)
        assert float(opt.flat_grad.abs().max()) == 0.0  # the gradient reset is part of the kernel
    for po, pr in zip(ours, ref):
        torch.testing.assert_close(po.detach(), pr.detach(), rtol=2e-5, atol=2e-7)
    sd = opt.ema_state_dict(torch.nn.ParameterList(ours))
    for (k, e), er in zip(sd.items(), ema_ref):
        torch.testing.assert_close(e, er, rtol=2e-5, atol=2e-7)


def test_flat_adam_views_and_zero_grad_flag():
    ours = _params(3)
    before = [p.detach().clone() for p in ours]
    opt = FlatAdam(ours, lr=1e-3)
    for p, b in zip(ours, before):
        assert torch.equal(p.detach(), b)                      # re-pointing to the flat buffer keeps the values
        assert p.grad is not None and p.grad.data_ptr() >= opt.flat_grad.data_ptr()
        p.grad.fill_(1.0)
    opt.step(zero_grad=False)
    assert float(opt.flat_grad.min()) == 1.0                   # kept for gradient accumulation
    assert all(float((p.detach() - b).abs().max()) > 0 for p, b in zip(ours, before))
    ours[0].grad = None
    opt.check_views()
    assert ours[0].grad.data_ptr() == opt.flat_grad.data_ptr()
