"""Drop-in ``Renderer`` with the call signatures of ``src/utils/Renderer.py`` of the reference.

``Tracker.py:150,175``, ``Mapper.py:548,591``, ``Mesher.py:548`` and ``Visualizer.py:79,244``
call these methods unchanged.  Each method is plumbing around the fused CUDA kernels:

  eval_points         Renderer.py:24-62    -> ens_eval_points
  render_batch_ray    Renderer.py:64-199   -> ens_render_fwd / ens_render_bwd (autograd.Function)
  render_img          Renderer.py:201-256  -> ens_lattice_rays + ens_render_fwd per 100k-ray batch
  render_img_rescale  Renderer.py:258-319  -> same, with gradient to c2w
  regulation          Renderer.py:322-360  -> iMAP only: unsupported (NICE configs never call it)

There is no PyTorch/CPU fallback: without ``libens_render.so`` or a CUDA device every method raises.
"""
from __future__ import annotations

import torch

from .common import get_rays, get_rays_rescale
from .functional import RenderSetup, eval_points as _eval_points, render_batch_ray as _render_batch_ray
from .scene import SceneCache


class Renderer(object):
    def __init__(self, cfg, args, slam, points_batch_size=500000, ray_batch_size=100000):
        self.ray_batch_size = ray_batch_size
        # Opt-in for the TRACKER's renderer (one line in Tracker.__init__: ``self.renderer.decoder_grads = False``): its
        # decoders are a deepcopy with requires_grad=True (Tracker.py:248-260) that no optimiser ever steps, so the reference
        # computes a weight-gradient pass nobody reads.  False = render with the decoder weights as constants: the pose-only
        # backward (no weight-gradient GEMMs, no activations kept), whatever requires_grad says.
        self.decoder_grads = True
        self.points_batch_size = points_batch_size      # kept for API parity; the fused kernel has no
                                                        # per-point HBM intermediates, so no point chunking
        self.lindisp = cfg['rendering']['lindisp']
        self.perturb = cfg['rendering']['perturb']
        self.N_samples = cfg['rendering']['N_samples']
        self.N_surface = cfg['rendering']['N_surface']
        self.N_importance = cfg['rendering']['N_importance']

        self.scale = cfg['scale']
        self.occupancy = cfg['occupancy']
        self.nice = slam.nice
        self.bound = slam.bound
        self.coarse_bound_enlarge = cfg['model']['coarse_bound_enlarge']

        self.H, self.W, self.fx, self.fy, self.cx, self.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy
        self._cache = SceneCache()
        self._tvals = {}
        if not self.nice:
            raise NotImplementedError("the fused renderer implements the NICE configuration (run.py --nice); "
                                      "iMAP mode (single MLP, density compositing, N_importance>0) is out of scope")

    # ------------------------------------------------------------------------------------------
    def _t_vals(self, device):
        key = str(device)
        tv = self._tvals.get(key)
        if tv is None:
            # exactly the reference's two linspace calls (Renderer.py:127-128 on CPU, 153 on device)
            t32 = torch.linspace(0., 1., steps=self.N_samples, device=device)
            t64 = torch.linspace(0., 1., steps=self.N_surface).double().to(device) if self.N_surface > 0 else None
            tv = (t32, t64)
            self._tvals[key] = tv
        return tv

    def _setup(self, stage, decoders, device):
        t32, t64 = self._t_vals(device)
        cbound = decoders.coarse_decoder.bound if hasattr(decoders, "coarse_decoder") else \
            self.bound * self.coarse_bound_enlarge
        return RenderSetup(stage, self.N_samples, self.N_surface, t32, t64, self.bound, cbound, self._cache,
                           n_importance=self.N_importance, lindisp=self.lindisp, perturb=self.perturb,
                           occupancy=self.occupancy)

    # ------------------------------------------------------------------------------------------
    def eval_points(self, p, decoders, c=None, stage='color', device='cuda:0'):
        """Occupancy / colour of points: (N,4) float32, occ = 100 outside the bound (Renderer.py:24-62)."""
        setup = self._setup(stage, decoders, p.device)
        return _eval_points(setup, c, decoders, p, apply_mask=True)

    def render_batch_ray(self, c, decoders, rays_d, rays_o, device, stage, gt_depth=None):
        """Render colour, depth and uncertainty of a batch of rays (Renderer.py:64-199).

        Returns depth (N,) float64, uncertainty (N,) float64, color (N,3) float32.
        """
        setup = self._setup(stage, decoders, rays_o.device)
        return _render_batch_ray(setup, c, decoders, rays_d, rays_o, gt_depth, freeze_params=not self.decoder_grads)

    def render_batch_ray_aux(self, c, decoders, rays_d, rays_o, device, stage, gt_depth=None):
        """render_batch_ray that also returns (raw, z_vals, weights) -- used by the parity tests."""
        setup = self._setup(stage, decoders, rays_o.device)
        return _render_batch_ray(setup, c, decoders, rays_d, rays_o, gt_depth, want_aux=True)

    def _render_rays_batched(self, c, decoders, rays_o, rays_d, device, stage, gt_depth):
        depth_list, uncertainty_list, color_list = [], [], []
        ray_batch_size = self.ray_batch_size
        for i in range(0, rays_d.shape[0], ray_batch_size):       # same batch boundaries as the reference:
            rays_d_batch = rays_d[i:i + ray_batch_size]           # the two depth maxima are per batch
            rays_o_batch = rays_o[i:i + ray_batch_size]
            gt_depth_batch = None if gt_depth is None else gt_depth[i:i + ray_batch_size]
            depth, uncertainty, color = self.render_batch_ray(
                c, decoders, rays_d_batch, rays_o_batch, device, stage, gt_depth=gt_depth_batch)
            depth_list.append(depth.double())
            uncertainty_list.append(uncertainty.double())
            color_list.append(color)
        return torch.cat(depth_list, dim=0), torch.cat(uncertainty_list, dim=0), torch.cat(color_list, dim=0)

    def render_img(self, c, decoders, c2w, device, stage, gt_depth=None):
        """Render depth, uncertainty and colour images, no grad (Renderer.py:201-256)."""
        with torch.no_grad():
            H, W = self.H, self.W
            rays_o, rays_d = get_rays(H, W, self.fx, self.fy, self.cx, self.cy, c2w, device)
            rays_o = rays_o.reshape(-1, 3)
            rays_d = rays_d.reshape(-1, 3)
            gt_depth = gt_depth.reshape(-1)
            depth, uncertainty, color = self._render_rays_batched(c, decoders, rays_o, rays_d, device, stage, gt_depth)
            return depth.reshape(H, W), uncertainty.reshape(H, W), color.reshape(H, W, 3)

    def render_img_rescale(self, c, decoders, c2w, device, stage, gt_depth=None, scale_factor=0.1):
        """Render down-scaled images WITH gradient (Renderer.py:258-319)."""
        from torchvision import transforms
        H, W = self.H, self.W
        new_H, new_W = int(H * scale_factor), int(W * scale_factor)
        rays_o, rays_d = get_rays_rescale(H, W, new_H, new_W, self.fx, self.fy, self.cx, self.cy, c2w, device)
        rays_o = rays_o.reshape(-1, 3)
        rays_d = rays_d.reshape(-1, 3)
        if gt_depth is not None:
            transform = transforms.Resize((new_H, new_W), interpolation=transforms.InterpolationMode.BILINEAR)
            gt_depth = transform(gt_depth.unsqueeze(0)).reshape(-1)
        depth, uncertainty, color = self._render_rays_batched(c, decoders, rays_o, rays_d, device, stage, gt_depth)
        return depth.reshape(new_H, new_W), uncertainty.reshape(new_H, new_W), color.reshape(new_H, new_W, 3)

    def regulation(self, c, decoders, rays_d, rays_o, gt_depth, device, stage='color'):
        raise NotImplementedError("Renderer.regulation is iMAP-only (Mapper.py:565-570 guards it with "
                                  "`not self.occupancy`); the fused renderer implements the NICE configuration")
