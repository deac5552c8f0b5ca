"""oracle/unet_input_oracle.py against torchvision + torch themselves (the lines of event_net.py:74-87), CPU."""
import numpy as np
import pytest
import torch
from torchvision import transforms

import unet_input_oracle as uo


def _reference(img1, img2, scale_factor):
    img1 = img1.permute(2, 0, 1)
    img2 = img2.permute(2, 0, 1)
    if scale_factor != 1.0:
        c, h, w = img1.shape
        h_new, w_new = int(scale_factor * h), int(scale_factor * w)
        transform = transforms.Resize((h_new, w_new), interpolation=transforms.InterpolationMode.NEAREST)
        img1 = transform(img1)
        img2 = transform(img2)
    img_pair = torch.cat((img1, img2), dim=0)
    return img_pair.unsqueeze(0).to(dtype=torch.float32)


@pytest.mark.parametrize("shape,scale", [((680, 1200), 0.15), ((260, 346), 0.15), ((102, 180), 1.0), ((37, 53), 0.5), ((16, 20), 0.3)])
def test_oracle_matches_torchvision(shape, scale):
    rng = np.random.RandomState(3)
    a = rng.rand(*shape, 3)                       # the previous GT colour image is float64 (datasets.py:179)
    b = rng.rand(*shape, 3).astype(np.float32)
    ref = _reference(torch.from_numpy(a), torch.from_numpy(b).double(), scale).numpy()
    got = uo.assemble_input(a, b, scale)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    # backward: autograd of the reference lines
    t = torch.from_numpy(b).requires_grad_(True)
    out = _reference(torch.from_numpy(a).float(), t, scale)
    g = torch.from_numpy(rng.randn(*out.shape).astype(np.float32))
    (out * g).sum().backward()
    assert np.allclose(uo.assemble_input_backward(g.numpy(), *shape), t.grad.numpy(), atol=1e-6)
