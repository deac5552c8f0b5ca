"""Drop-ins for the hot-path functions of ``src/common.py`` (same names and signatures).

Pixel draws stay ``torch.randint`` on the caller's device/generator (bit-exact indices,
common.py:99,116); everything after the draw is one CUDA kernel.  Quaternion maths stays in
PyTorch: autograd finishes the pose gradient from ``d c2w``.
"""
from __future__ import annotations

import torch

from . import _ext
from .functional import TIMER, _LatticeRays, _SampleRays, _c2w_dev

_LIN_CACHE = {}


def _linspace_dev(lo, hi, steps, device):
    """torch.linspace on the CPU, moved to the device -- exactly as common.py:308,328 build it."""
    key = (float(lo), float(hi), int(steps), str(device))
    t = _LIN_CACHE.get(key)
    if t is None:
        t = torch.linspace(lo, hi, steps).to(device)
        _LIN_CACHE[key] = t
    return t


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device):
    """Get n rays from the image region H0..H1, W0..W1 (common.py:160-169)."""
    indices = torch.randint((H1 - H0) * (W1 - W0), (n,), device=device)     # common.py:99
    c2w = _c2w_dev(c2w, device)
    ext = _ext.module() if (_ext.ENABLED and not TIMER.enabled) else None
    if ext is not None:                                                     # autograd plumbing in C++ (csrc/ens_torch.cpp)
        return tuple(ext.sample_rays(c2w, indices, depth, color, [int(H0), int(H1), int(W0), int(W1)],
                                     [float(H), float(W), float(fx), float(fy), float(cx), float(cy)]))
    rays_o, rays_d, sample_depth, sample_color = _SampleRays.apply(
        c2w, indices, (int(H0), int(H1), int(W0), int(W1)),
        (int(H), int(W), float(fx), float(fy), float(cx), float(cy)), depth, color)
    return rays_o, rays_d, sample_depth, sample_color


def get_samples_event(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, event1, event2, device):
    """Get n rays from the image region H0..H1, W0..W1 together with the two event images' samples (common.py:178-187).
    One draw (common.py:116) serves rays, depth, colour and both event images, as in the reference."""
    indices = torch.randint((H1 - H0) * (W1 - W0), (n,), device=device)     # common.py:116 (the clamp at :117 is a no-op)
    c2w = _c2w_dev(c2w, device)
    ext = _ext.module() if (_ext.ENABLED and not TIMER.enabled) else None
    if ext is not None:
        rays_o, rays_d, sample_depth, sample_color = ext.sample_rays(
            c2w, indices, depth, color, [int(H0), int(H1), int(W0), int(W1)],
            [float(H), float(W), float(fx), float(fy), float(cx), float(cy)])
    else:
        rays_o, rays_d, sample_depth, sample_color = _SampleRays.apply(
            c2w, indices, (int(H0), int(H1), int(W0), int(W1)),
            (int(H), int(W), float(fx), float(fy), float(cx), float(cy)), depth, color)
    sample_event1 = event1[H0:H1, W0:W1].reshape(-1, 2)[indices]            # common.py:122-127
    sample_event2 = event2[H0:H1, W0:W1].reshape(-1, 2)[indices]
    return rays_o, rays_d, sample_depth, sample_color, sample_event1, sample_event2


def get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device):
    """Rays of arbitrary pixel coordinates (common.py:74-89): the lattice kernel in pairs mode."""
    c2w = _c2w_dev(c2w, device)
    i = i.reshape(-1).float().contiguous()
    j = j.reshape(-1).float().contiguous()
    from .functional import _PairRays
    return _PairRays.apply(c2w, i, j, (int(H), int(W), float(fx), float(fy), float(cx), float(cy)))


def get_rays(H, W, fx, fy, cx, cy, c2w, device):
    """Rays of the whole image, (H,W,3) each (common.py:300-318)."""
    c2w = _c2w_dev(c2w, device)
    lin_w = _linspace_dev(0, W - 1, W, device)
    lin_h = _linspace_dev(0, H - 1, H, device)
    return _LatticeRays.apply(c2w, lin_w, lin_h, (int(H), int(W), float(fx), float(fy), float(cx), float(cy)))


def get_rays_rescale(H, W, new_H, new_W, fx, fy, cx, cy, c2w, device):
    """Rays of the down-scaled image lattice (common.py:320-340)."""
    c2w = _c2w_dev(c2w, device)
    lin_w = _linspace_dev(0, W - 1, new_W, device)
    lin_h = _linspace_dev(0, H - 1, new_H, device)
    return _LatticeRays.apply(c2w, lin_w, lin_h, (int(H), int(W), float(fx), float(fy), float(cx), float(cy)))


def quad2rotation(quad):
    """Quaternion (batch,4) -> rotation (batch,3,3); PyTorch so autograd passes (common.py:189-212)."""
    bs = quad.shape[0]
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rot_mat = torch.zeros(bs, 3, 3, device=quad.device)
    rot_mat[:, 0, 0] = 1 - two_s * (qj ** 2 + qk ** 2)
    rot_mat[:, 0, 1] = two_s * (qi * qj - qk * qr)
    rot_mat[:, 0, 2] = two_s * (qi * qk + qj * qr)
    rot_mat[:, 1, 0] = two_s * (qi * qj + qk * qr)
    rot_mat[:, 1, 1] = 1 - two_s * (qi ** 2 + qk ** 2)
    rot_mat[:, 1, 2] = two_s * (qj * qk - qi * qr)
    rot_mat[:, 2, 0] = two_s * (qi * qk - qj * qr)
    rot_mat[:, 2, 1] = two_s * (qj * qk + qi * qr)
    rot_mat[:, 2, 2] = 1 - two_s * (qi ** 2 + qj ** 2)
    return rot_mat


def get_camera_from_tensor(inputs):
    """[quat, T] (7) -> (3,4) [R|t] (common.py:215-228).  CUDA float32 tensors take the fused kernel
    (one launch forward, one backward, instead of ~60 eager ops per pose)."""
    if inputs.is_cuda and inputs.dtype == torch.float32:
        ext = _ext.module() if (_ext.ENABLED and not TIMER.enabled) else None
        if ext is not None:
            return ext.pose_to_c2w(inputs.unsqueeze(0))[0] if inputs.dim() == 1 else ext.pose_to_c2w(inputs)
        from .functional import _PoseToC2W
        if inputs.dim() == 1:
            return _PoseToC2W.apply(inputs.unsqueeze(0))[0]
        return _PoseToC2W.apply(inputs)
    return get_camera_from_tensor_torch(inputs)


def get_camera_from_tensor_torch(inputs):
    """Eager-PyTorch form (CPU tensors / other dtypes)."""
    N = len(inputs.shape)
    if N == 1:
        inputs = inputs.unsqueeze(0)
    quad, T = inputs[:, :4], inputs[:, 4:]
    R = quad2rotation(quad)
    RT = torch.cat([R, T[:, :, None]], 2)
    if N == 1:
        RT = RT[0]
    return RT
