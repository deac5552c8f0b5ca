"""ens_unet_input / ens_unet_input_bwd (event_net.assemble_input) against the oracle; needs a B200 (``-m gpu``)."""
import numpy as np
import pytest
import torch

import unet_input_oracle as uo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,scale,f64", [((680, 1200), 0.15, True), ((260, 346), 0.15, False), ((102, 180), 1.0, True), ((16, 20), 0.3, False)])
def test_assemble_input_matches_oracle_bitwise(shape, scale, f64):
    from evennicer_slam_b200.event_net import assemble_input
    rng = np.random.RandomState(5)
    a = rng.rand(*shape, 3) if f64 else rng.rand(*shape, 3).astype(np.float32)
    b = rng.rand(*shape, 3).astype(np.float32)
    tb = torch.from_numpy(b).to(DEV).requires_grad_(True)
    out = assemble_input(torch.from_numpy(a).to(DEV), tb, scale)
    want = uo.assemble_input(a, b, scale)
    assert out.dtype == torch.float32 and tuple(out.shape) == want.shape
    assert np.array_equal(out.detach().cpu().numpy(), want)
    g = rng.randn(*want.shape).astype(np.float32)
    (out * torch.from_numpy(g).to(DEV)).sum().backward()
    assert np.allclose(tb.grad.cpu().numpy(), uo.assemble_input_backward(g, *shape), atol=1e-6)


def test_inference_event_drop_in_runs_a_network():
    from evennicer_slam_b200.event_net import inference_event

    class Tiny(torch.nn.Module):                   # stands in for UNet_2heads: (events [1,2,h,w], masks [1,2,h,w])
        def __init__(self):
            super().__init__()
            self.c = torch.nn.Conv2d(6, 4, 3, padding=1)

        def forward(self, x):
            y = self.c(x)
            return y[:, :2], torch.sigmoid(y[:, 2:])
    net = Tiny().to(DEV)
    a = torch.rand(24, 30, 3, device=DEV, dtype=torch.float64)
    b = torch.rand(24, 30, 3, device=DEV, requires_grad=True)
    ev, mask = inference_event(net, a, b, DEV, scale_factor=1.0)
    assert tuple(ev.shape) == (24, 30, 2) and tuple(mask.shape) == (1, 2, 24, 30)
    ev.sum().backward()
    assert b.grad is not None and torch.isfinite(b.grad).all() and float(b.grad.abs().sum()) > 0
