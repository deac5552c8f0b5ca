"""Fused loss glue around the renderer.

* ``mapper_loss`` / ``tracker_loss``: the RGB-D losses of Mapper.py:553-562 and
  Tracker.py:180-196 -- value and the gradients with respect to the rendered depth and colour in ONE launch instead of
  ~25 eager ones (abs, sqrt, median, boolean-mask gathers with their host syncs, sums, and their autograd transposes).
  No boolean indexing means no host sync: the whole iteration can be captured in a CUDA graph.
* ``event_loss``: the blurred-L2 event loss of Tracker.py:204-224 and Mapper.py:593-615 as ONE
  launch that returns the loss and its gradient with respect to the predicted event image.

    loss, parts = event_loss(gt_event, full_event, kernel_sizes=[9], kernel_weights=[1], balancer=0.025)
    loss.backward()                       # gradient flows into full_event (the UNet output) as in the reference
    parts                                  # double[2 + n]: balancer * total, unblurred sum, blurred sum per kernel
                                           # (what the reference logs as losses_event_list)
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib


def gaussian_kernel1d(kernel_size: int, sigma: float = None) -> np.ndarray:
    """torchvision's ``_get_gaussian_kernel1d`` in float32 (sigma default of ``gaussian_blur``: 0.15 ks + 0.35)."""
    if kernel_size < 1 or kernel_size % 2 == 0:
        raise ValueError(f"kernel_size should be odd and positive, got {kernel_size}")
    if sigma is None:
        sigma = kernel_size * 0.15 + 0.35
    half = (kernel_size - 1) * 0.5
    x = torch.linspace(-half, half, steps=kernel_size, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return (pdf / pdf.sum()).numpy()


class _EventLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, kernel_sizes, kernel_weights, balancer):
        if not (pred.is_cuda and gt.is_cuda):
            raise RuntimeError("event_loss needs CUDA tensors (there is no CPU fallback)")
        if pred.dim() != 3 or pred.shape != gt.shape:
            raise ValueError(f"gt and pred must both be [H,W,C], got {tuple(gt.shape)} and {tuple(pred.shape)}")
        L = _lib.lib()
        p = pred.detach().float().contiguous()
        g = gt.detach().float().contiguous()
        H, W, Cc = p.shape
        n = len(kernel_sizes)
        k1d = np.concatenate([gaussian_kernel1d(int(ks)) for ks in kernel_sizes]).astype(np.float32) if n else np.zeros(0, np.float32)
        ks_arr = (C.c_int * max(n, 1))(*[int(k) for k in kernel_sizes])
        k1_arr = (C.c_float * max(len(k1d), 1))(*k1d.tolist())
        w_arr = (C.c_float * max(n, 1))(*[float(w) for w in kernel_weights])
        parts = torch.empty(2 + n, dtype=torch.float64, device=p.device)
        need = ctx.needs_input_grad[0]
        g_pred = torch.empty_like(p) if need else None
        from .functional import TIMER
        TIMER.launches += 1
        _lib.check(L.ens_event_loss(_lib.ptr(g), _lib.ptr(p), H, W, Cc, ks_arr, k1_arr, w_arr, n, float(balancer),
                                    _lib.ptr(parts), _lib.ptr(g_pred), _lib.cur_stream(p.device)), "ens_event_loss")
        ctx.g_pred = g_pred
        ctx.pred_dtype = pred.dtype
        ctx.mark_non_differentiable(parts)
        return parts[0].to(torch.float32), parts

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        if ctx.g_pred is None:
            return None, None, None, None, None
        return (ctx.g_pred * g_loss.to(ctx.g_pred.dtype)).to(ctx.pred_dtype), None, None, None, None


def event_loss(gt_event: torch.Tensor, full_event: torch.Tensor, kernel_sizes: Sequence[int] = (9,),
               kernel_weights: Sequence[float] = (1.0,), balancer: float = 1.0, blur: bool = True
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``balancer * (sum (gt-pred)^2 + sum_k w_k sum (blur_k(gt) - blur_k(pred))^2)`` (float32 scalar, differentiable in
    ``full_event``) and the double vector of its parts.  ``blur=False`` is the reference's un-blurred branch."""
    if len(kernel_sizes) != len(kernel_weights):
        raise ValueError("kernel_sizes and kernel_weights must have the same length")
    if not blur:
        kernel_sizes, kernel_weights = (), ()
    return _EventLoss.apply(full_event, gt_event, tuple(kernel_sizes), tuple(kernel_weights), float(balancer))


class _RgbdLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, color, uncertainty, gt_depth, gt_color, tracker, use_color, w_color, handle_dynamic):
        if not depth.is_cuda:
            raise RuntimeError("the fused RGB-D losses need CUDA tensors (there is no CPU fallback)")
        L = _lib.lib()
        n = depth.shape[0]
        d = depth.detach().to(torch.float64).contiguous()
        gd = gt_depth.detach().to(torch.float32).contiguous()
        col = gcol = None
        f64 = 0
        if use_color:
            col = color.detach().to(torch.float32).contiguous()
            gcol = gt_color.detach().contiguous()
            if gcol.dtype == torch.float64:
                f64 = 1
            elif gcol.dtype != torch.float32:
                gcol = gcol.float()
            if tuple(col.shape) != (n, 3) or tuple(gcol.shape) != (n, 3):
                raise ValueError("color and gt_color must be [n,3]")
        var = uncertainty.detach().to(torch.float64).contiguous() if tracker else None
        work = torch.empty(n, dtype=torch.float64, device=d.device) if tracker else None
        loss = torch.empty(1, dtype=torch.float64, device=d.device)
        need_d, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and use_color
        g_d = torch.empty_like(d) if need_d else None
        g_c = torch.empty_like(col) if need_c else None
        from .functional import TIMER
        TIMER.launches += 1
        _lib.check(L.ens_rgbd_loss(1 if tracker else 0, _lib.ptr(d), _lib.ptr(var), _lib.ptr(col), _lib.ptr(gd),
                                   _lib.ptr(gcol), f64, n, 1 if use_color else 0, float(w_color),
                                   1 if handle_dynamic else 0, _lib.ptr(loss), _lib.ptr(g_d), _lib.ptr(g_c),
                                   _lib.ptr(work), _lib.cur_stream(d.device)), "ens_rgbd_loss")
        ctx.g = (g_d, g_c)
        ctx.dtypes = (depth.dtype, color.dtype if color is not None else None)
        return loss[0]

    @staticmethod
    def backward(ctx, g_loss):
        g_d, g_c = ctx.g
        gd = (g_d * g_loss).to(ctx.dtypes[0]) if g_d is not None else None
        gc = (g_c * g_loss.to(torch.float32)).to(ctx.dtypes[1]) if g_c is not None else None
        return gd, gc, None, None, None, None, None, None, None


def mapper_loss(batch_gt_depth, batch_gt_color, depth, color, w_color_loss: float = 0.2, use_color: bool = True):
    """Mapper.py:553-562: ``|gt_depth - depth|[gt_depth > 0].sum() (+ w * |gt_color - color|.sum())`` -> float64 scalar.
    ``use_color`` is the reference's ``(not self.nice) or (self.stage == 'color')``."""
    return _RgbdLoss.apply(depth, color if use_color else None, None, batch_gt_depth,
                           batch_gt_color if use_color else None, False, use_color, w_color_loss, False)


def tracker_loss(batch_gt_depth, batch_gt_color, depth, uncertainty, color, w_color_loss: float = 0.2,
                 use_color_in_tracking: bool = True, handle_dynamic: bool = True):
    """Tracker.py:180-196 (uncertainty is detached there, and here): -> float64 scalar."""
    return _RgbdLoss.apply(depth, color if use_color_in_tracking else None, uncertainty, batch_gt_depth,
                           batch_gt_color if use_color_in_tracking else None, True, use_color_in_tracking,
                           w_color_loss, handle_dynamic)
