"""Frustum feature selection and keyframe overlap of the mapper on the GPU.

``FrustumSelector`` carries the two ``Mapper`` methods with their names, arguments and return values
(src/Mapper.py:115-186 ``get_mask_from_c2w``, :188-250 ``keyframe_selection_overlap``); the voxel / sample-point work runs
in ``ens_frustum_mask`` / ``ens_keyframe_overlap`` instead of numpy + cv2.remap.  What stays on the host is what the
reference also does there and what is a few hundred bytes: the 4x4 inverse of the pose (``np.linalg.inv``, so that the
result is the reference's bit for bit) and the final sort / ``np.random.permutation`` of the keyframe ids (the reference's
own random stream).

    sel = FrustumSelector(H, W, fx, fy, cx, cy, bound, device)
    mask = sel.voxel_mask(c2w, key, val.shape[2:], cur_gt_depth)      # bool [Z,Y,X] on the device -> FrustumGridAdam
    mask_np = sel.get_mask_from_c2w(c2w, key, val.shape[2:], depth_np)   # the reference's numpy bool [X,Y,Z]
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from .common import get_samples


class FrustumSelector:
    def __init__(self, H, W, fx, fy, cx, cy, bound, device):
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = int(H), int(W), float(fx), float(fy), float(cx), float(cy)
        self.bound = torch.as_tensor(bound).detach().cpu()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FrustumSelector needs a CUDA device (there is no CPU fallback)")
        self._axes: Dict[Tuple[int, int, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._cam6 = (C.c_double * 6)(self.H, self.W, self.fx, self.fy, self.cx, self.cy)

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _grid_axes(self, val_shape):
        """torch.linspace of the bound on the CPU exactly as Mapper.py:132-134 calls it, uploaded once per grid shape."""
        key = (int(val_shape[0]), int(val_shape[1]), int(val_shape[2]))
        if key not in self._axes:
            b = self.bound
            self._axes[key] = tuple(torch.linspace(b[a][0], b[a][1], n).to(self.device)
                                    for a, n in ((0, key[2]), (1, key[1]), (2, key[0])))
        return self._axes[key]

    def _depth_dev(self, depth):
        if isinstance(depth, np.ndarray):
            depth = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32))
        return depth.detach().to(device=self.device, dtype=torch.float32).contiguous()

    # -- Mapper.get_mask_from_c2w ----------------------------------------------------------------------------------
    def voxel_mask(self, c2w, key, val_shape, depth, zyx: bool = True) -> torch.Tensor:
        """bool tensor on the device: [Z,Y,X] (= ``torch.from_numpy(mask).permute(2,1,0)`` of the reference, the layout of
        the grids) or, with ``zyx=False``, the reference's [X,Y,Z]."""
        Z, Y, X = int(val_shape[0]), int(val_shape[1]), int(val_shape[2])
        shape = (Z, Y, X) if zyx else (X, Y, Z)
        if key == 'grid_coarse':                                        # :137-139
            return torch.ones(shape, dtype=torch.bool, device=self.device)
        L = _lib.lib()
        c2w_np = np.ascontiguousarray(torch.as_tensor(c2w).detach().cpu().numpy())   # :141 (float32 stays float32)
        w2c = np.ascontiguousarray(np.linalg.inv(c2w_np).astype(np.float32))         # :142
        centre = np.ascontiguousarray(c2w_np[:3, 3].astype(np.float32))
        xs, ys, zs = self._grid_axes((Z, Y, X))
        d = self._depth_dev(depth)
        if tuple(d.shape) != (self.H, self.W):
            raise ValueError(f"depth image must be [{self.H},{self.W}], got {tuple(d.shape)}")
        ws_bytes = L.ens_frustum_workspace_bytes(X, Y, Z)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        mask = torch.empty(shape, dtype=torch.uint8, device=self.device)
        _lib.check(L.ens_frustum_mask(w2c.ctypes.data_as(C.c_void_p), centre.ctypes.data_as(C.c_void_p),
                                      C.cast(self._cam6, C.c_void_p), _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(zs), X, Y, Z,
                                      _lib.ptr(d), self.H, self.W, int(zyx), _lib.ptr(mask), _lib.ptr(ws), ws_bytes,
                                      _lib.cur_stream(self.device)), "ens_frustum_mask")
        return mask.view(torch.bool)

    def get_mask_from_c2w(self, c2w, key, val_shape, depth_np):
        """The reference's signature and return value: numpy bool [val_shape[2], val_shape[1], val_shape[0]]."""
        return self.voxel_mask(c2w, key, val_shape, depth_np, zyx=False).cpu().numpy()

    # -- Mapper.keyframe_selection_overlap -------------------------------------------------------------------------
    def keyframe_overlap_counts(self, vertices: torch.Tensor, kf_c2w, edge: float = 20) -> np.ndarray:
        """vertices [N,3] float32 on the device; kf_c2w: list of 4x4 poses (tensors).  int32 counts per keyframe (host)."""
        L = _lib.lib()
        K = len(kf_c2w)
        if K == 0:
            return np.zeros(0, np.int32)
        c2ws = torch.stack([torch.as_tensor(c).detach().to(self.device) for c in kf_c2w]).cpu().numpy()   # ONE device->host copy
        w2cs = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(c2ws).astype(np.float32))).to(self.device)  # :225 per keyframe
        v = vertices.detach().to(device=self.device, dtype=torch.float32).contiguous()
        counts = torch.empty(K, dtype=torch.int32, device=self.device)
        _lib.check(L.ens_keyframe_overlap(_lib.ptr(w2cs), K, C.cast(self._cam6, C.c_void_p), float(edge), _lib.ptr(v),
                                          v.shape[0], _lib.ptr(counts), _lib.cur_stream(self.device)), "ens_keyframe_overlap")
        return counts.cpu().numpy()

    def keyframe_selection_overlap(self, gt_color, gt_depth, c2w, keyframe_dict, k, N_samples=16, pixels=100):
        """Mapper.py:188-250 with the per-keyframe projection on the GPU; same random draws (torch.randint inside
        get_samples, then np.random.permutation), same return value (list of keyframe ids)."""
        device = self.device
        H, W, fx, fy, cx, cy = self.H, self.W, self.fx, self.fy, self.cx, self.cy
        rays_o, rays_d, gt_depth, gt_color = get_samples(
            0, H, 0, W, pixels, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color, device)
        gt_depth = gt_depth.reshape(-1, 1)
        gt_depth = gt_depth.repeat(1, N_samples)
        t_vals = torch.linspace(0., 1., steps=N_samples).to(device)
        near = gt_depth * 0.8
        far = gt_depth + 0.5
        z_vals = near * (1. - t_vals) + far * (t_vals)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]      # [N_rays, N_samples, 3]
        vertices = pts.reshape(-1, 3).float()
        counts = self.keyframe_overlap_counts(vertices, [kf['est_c2w'] for kf in keyframe_dict])
        n = vertices.shape[0]
        list_keyframe = [{'id': i, 'percent_inside': int(c) / n} for i, c in enumerate(counts)]
        list_keyframe = sorted(list_keyframe, key=lambda i: i['percent_inside'], reverse=True)
        selected_keyframe_list = [dic['id'] for dic in list_keyframe if dic['percent_inside'] > 0.00]
        selected_keyframe_list = list(np.random.permutation(np.array(selected_keyframe_list))[:k])
        return selected_keyframe_list
