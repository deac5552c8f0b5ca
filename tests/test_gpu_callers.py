"""Tracker.py / Mapper.py keep working UNCHANGED on the drop-ins: the callers' loop bodies (tests/golden/caller_loops.py, a
line-for-line transcript of src/Mapper.py:498-578 and src/Tracker.py:159-201 -- per-keyframe get_samples, the inside_mask
pre-filter with boolean indexing, render_batch_ray, boolean-mask losses, backward) run on ``evennicer_slam_b200``'s
``get_samples`` / ``get_camera_from_tensor`` / ``Renderer`` and reproduce what the same lines gave on the reference itself
(tests/golden/caller_loops.npz, written by make_caller_golden.py).  Needs a B200 (``-m gpu``).

The pixel draws of the golden run (torch.randint on the CPU generator) are replayed: the CUDA generator draws another stream.
"""
import types

import numpy as np
import pytest
import torch

import caller_loops as cl
import cases
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_OUT, TOL_GRAD = 1e-4, 1e-3


class _ReplayRandint:
    def __init__(self, draws):
        self.draws, self.k, self.orig = draws, 0, torch.randint

    def __call__(self, high, size, device=None, **kw):
        r = torch.from_numpy(self.draws[self.k]).to(device)
        assert tuple(r.shape) == tuple(size) and int(r.max()) < high
        self.k += 1
        return r


@pytest.fixture(scope="module")
def world():
    from evennicer_slam_b200 import harness, common
    scene = cases.tiny_scene()
    decoders, c, renderer, cfg = harness.build(scene, DEV)
    cam = scene.cam
    ns = types.SimpleNamespace(get_samples=common.get_samples, get_camera_from_tensor=common.get_camera_from_tensor,
                               renderer=renderer, H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy)
    return dict(scene=scene, decoders=decoders, c=c, ns=ns, g=load_golden("caller_loops.npz"),
                bound=torch.from_numpy(scene.bound.copy()), frames=cl.caller_inputs())


@pytest.mark.parametrize("ext", [True, False])
@pytest.mark.parametrize("stage", ["middle", "fine", "color"])
def test_mapper_iteration_unchanged_on_the_dropins(world, stage, ext, monkeypatch):
    from evennicer_slam_b200 import common, _ext
    monkeypatch.setattr(_ext, "ENABLED", ext)          # the C++ autograd plumbing and the Python one
    g, frames, decoders = world["g"], world["frames"], world["decoders"]
    tag = "map." + stage
    for p in decoders.parameters():
        p.grad = None
    cg = {k: v.clone().requires_grad_(True) for k, v in world["c"].items()}
    cams = [torch.from_numpy(f[0].copy()).to(DEV).requires_grad_(True) for f in frames[1:]]
    kfs = [{"depth": torch.from_numpy(f[1]), "color": torch.from_numpy(f[2]),
            "est_c2w": common.get_camera_from_tensor(torch.from_numpy(f[0].copy()).to(DEV)).detach()} for f in frames[:3]]
    cur = {"depth": torch.from_numpy(frames[3][1]), "color": torch.from_numpy(frames[3][2])}
    replay = _ReplayRandint(g[tag + ".draws"])
    monkeypatch.setattr(torch, "randint", replay)
    loss, n_in, depth, color = cl.mapper_iteration(world["ns"], cg, decoders, kfs, cams, cur, world["bound"], stage, 12, DEV)
    monkeypatch.undo()
    assert replay.k == 4 and n_in == int(g[tag + ".n_inside"]) and n_in < 48          # the inside_mask filter dropped rays
    assert abs(loss - float(g[tag + ".loss"])) < TOL_OUT * abs(float(g[tag + ".loss"]))
    assert rel_err(depth.cpu().numpy(), g[tag + ".depth"]) < TOL_OUT and rel_err(color.cpu().numpy(), g[tag + ".color"]) < TOL_OUT
    for i, ct in enumerate(cams):
        assert rel_err(ct.grad.cpu().numpy(), g[f"{tag}.g_cam{i}"]) < TOL_GRAD, i
    for k, v in cg.items():
        key = f"{tag}.ggrid.{k}"
        if key in g.files:
            assert rel_err(v.grad.cpu().numpy(), g[key]) < TOL_GRAD, k
        else:
            assert v.grad is None, k                   # a level the stage does not touch gets no gradient, as in the reference
    for key in g.files:
        if key.startswith(tag + ".gdec."):
            _, _, _, lv, name = key.split(".", 4)
            got = dict(getattr(decoders, lv + "_decoder").named_parameters())[name].grad
            assert rel_err(got.cpu().numpy(), g[key]) < TOL_GRAD, key


@pytest.mark.parametrize("ext", [True, False])
def test_tracker_iteration_unchanged_on_the_dropins(world, ext, monkeypatch):
    from evennicer_slam_b200 import _ext
    monkeypatch.setattr(_ext, "ENABLED", ext)
    g, frames, decoders = world["g"], world["frames"], world["decoders"]
    ct = torch.from_numpy(frames[3][0].copy()).to(DEV).requires_grad_(True)
    cg = {k: v.clone() for k, v in world["c"].items()}
    replay = _ReplayRandint(g["track.draws"])
    monkeypatch.setattr(torch, "randint", replay)
    loss, n_in, n_mask = cl.tracker_iteration(world["ns"], cg, decoders, ct, torch.from_numpy(frames[3][1]).to(DEV),
                                              torch.from_numpy(frames[3][2]).to(DEV), world["bound"], 40, 2, 2, DEV)
    monkeypatch.undo()
    assert n_in == int(g["track.n_inside"]) and n_mask == int(g["track.n_mask"]) and n_in < 40
    assert abs(loss - float(g["track.loss"])) < TOL_OUT * abs(float(g["track.loss"]))
    assert rel_err(ct.grad.cpu().numpy(), g["track.g_cam"]) < TOL_GRAD
