"""Build the drop-in objects (decoders, grids, Renderer) for a synthetic scene.

Mirrors what ``EvenNICER_SLAM.__init__`` does for the hot path (EvenNICER_SLAM.py:31-123):
``get_model`` -> ``load_bound`` -> ``grid_init`` -> ``Renderer``; used by tests, smoke() and bench.
"""
from __future__ import annotations

import types

import torch

from . import synthetic as syn


def default_cfg() -> dict:
    """The hot-path keys of configs/nice_slam.yaml (lines 1-5, 105-113)."""
    return {
        "coarse": True, "scale": 1, "occupancy": True,
        "grid_len": dict(syn.GRID_LEN, bound_divisible=syn.BOUND_DIVISIBLE),
        "rendering": {"N_samples": 32, "N_surface": 16, "N_importance": 0, "lindisp": False, "perturb": 0.0},
        "data": {"dim": 3},
        "model": {"c_dim": 32, "coarse_bound_enlarge": 2, "pos_embedding_method": "fourier"},
    }


def build(scene: syn.Scene, device="cuda:0", cfg=None, requires_grad: bool = True, native_layout: bool = False):
    """(decoders: NICE, c: dict of grids, renderer: Renderer, cfg)."""
    from .decoder import NICE
    from .renderer import Renderer
    cfg = cfg or default_cfg()
    gl = cfg["grid_len"]
    decoders = NICE(dim=3, c_dim=cfg["model"]["c_dim"], coarse=cfg["coarse"], coarse_grid_len=gl["coarse"],
                    middle_grid_len=gl["middle"], fine_grid_len=gl["fine"], color_grid_len=gl["color"],
                    pos_embedding_method=cfg["model"]["pos_embedding_method"])
    bound = torch.from_numpy(scene.bound.copy())                   # CPU float64, as load_bound leaves it
    decoders.bound = bound
    decoders.middle_decoder.bound = bound
    decoders.fine_decoder.bound = bound
    decoders.color_decoder.bound = bound
    decoders.coarse_decoder.bound = bound * cfg["model"]["coarse_bound_enlarge"]
    for lv in syn.LEVELS:
        sd = {k: torch.from_numpy(v.copy()) for k, v in scene.decoders[lv].items()}
        getattr(decoders, lv + "_decoder").load_state_dict(sd)
    decoders = decoders.to(device)
    for p in decoders.parameters():
        p.requires_grad_(requires_grad)
    c = {k: torch.from_numpy(v.copy()).to(device) for k, v in scene.grids.items()}
    if native_layout:       # grids allocated in the kernels' layout, exposed as [1,32,Z,Y,X] views (scene.as_native_layout)
        from .scene import as_native_layout
        c = {k: as_native_layout(v) for k, v in c.items()}
    cam = scene.cam
    slam = types.SimpleNamespace(nice=True, bound=bound, H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy)
    renderer = Renderer(cfg, None, slam)
    return decoders, c, renderer, cfg
