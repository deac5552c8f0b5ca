"""Drop-in for ``src/event_net.py:inference_event`` (same name, arguments and returns): the tensor shuffling in front of the
UNet -- permute both images to CHW, optional nearest resize, concatenate, add the batch dimension, cast to float32 -- is
one CUDA launch (``ens_unet_input``), and so is its backward into the rendered colour image (``ens_unet_input_bwd``).
The UNet itself (cuDNN convolutions) is the caller's ``net`` and stays what it is (out of scope).
"""
from __future__ import annotations

import torch

from . import _lib


class _UNetInput(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, h, w):
        L = _lib.lib()
        dev = img2.device
        a = img1.detach()
        if a.dtype not in (torch.float32, torch.float64):
            a = a.float()
        a = a.contiguous()
        b = img2.detach()
        b = (b if b.dtype == torch.float32 else b.float()).contiguous()
        out = torch.empty((1, 6, h, w), dtype=torch.float32, device=dev)
        _lib.check(L.ens_unet_input(_lib.ptr(a), int(a.dtype == torch.float64), a.shape[0], a.shape[1], _lib.ptr(b), b.shape[0],
                                    b.shape[1], h, w, _lib.ptr(out), _lib.cur_stream(dev)), "ens_unet_input")
        ctx.shape2 = (b.shape[0], b.shape[1], h, w)
        ctx.dtype2 = img2.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        H2, W2, h, w = ctx.shape2
        gc = g.detach()
        gc = (gc if gc.dtype == torch.float32 else gc.float()).contiguous()
        g2 = torch.empty((H2, W2, 3), dtype=torch.float32, device=g.device)
        _lib.check(L.ens_unet_input_bwd(_lib.ptr(gc), H2, W2, h, w, _lib.ptr(g2), _lib.cur_stream(g.device)), "ens_unet_input_bwd")
        return None, g2.to(ctx.dtype2), None, None


def assemble_input(img1: torch.Tensor, img2: torch.Tensor, scale_factor: float = 1.0) -> torch.Tensor:
    """[1, 6, h, w] float32 network input from two HWC images (event_net.py:74-87); differentiable wrt img2."""
    assert img1.shape == img2.shape, 'The sizes of the two input images are not the same!'
    h, w = img1.shape[0], img1.shape[1]
    if scale_factor != 1.0:
        h, w = int(scale_factor * h), int(scale_factor * w)
        assert h > 0 and w > 0, 'Scale is too small, resized images would have no pixels'
    if not img2.is_cuda:
        raise RuntimeError("assemble_input needs CUDA tensors (there is no CPU fallback)")
    return _UNetInput.apply(img1.to(img2.device), img2, h, w)


def inference_event(net, img1, img2, device, scale_factor=1, out_threshold=0.5):
    """event_net.py:67-99 with the input assembly fused; returns (full_events [h,w,2], full_mask)."""
    net.eval()
    img_pair = assemble_input(img1, img2, scale_factor).to(device=device)
    events_pred, masks_pred = net(img_pair)
    mask_prob = masks_pred[:, 1][:, None, :, :]
    events_pred_roi = (events_pred * mask_prob)[0]
    full_mask = masks_pred
    full_events = events_pred_roi.squeeze().permute(1, 2, 0)
    return full_events, full_mask
