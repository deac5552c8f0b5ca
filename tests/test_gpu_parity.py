"""Parity of the CUDA path (through the C ABI / the Python drop-ins) with the oracle and the
reference-generated goldens.  Needs a B200 (``-m gpu``).

Tolerances (BASELINE.json north_star): pixel indices and sample placement bit-exact; rendered
depth / colour 1e-4 relative; gradients 1e-3 relative (atomic ordering).
"""
import numpy as np
import pytest
import torch

import cases
import render_oracle as orc
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu

STAGES = ("coarse", "middle", "fine", "color")
DEV = "cuda:0"
TOL_OUT = 1e-4
TOL_GRAD = 1e-3
# room0-scale batches hold ~2e7 relu decisions, a few of them within float32 rounding of the kink.  The default path (tcgen05
# forward + backward) differs from the reference-CPU goldens in 4 decisions and meets 1e-3 outright (measured 1.9e-4 on the
# worst decoder tensor; the reference's own CUDA run differs from its CPU run by 2e-5: tests/test_gpu_reference.py).  Only the
# mma.sync A/B kernels (ENS_MAP_TC=0), whose chained-register forward rounds differently, need the wider bound (1.2e-3).
import os as _os
TOL_GRAD_KINK = 1e-3 if _os.environ.get("ENS_MAP_TC", "1") != "0" else 3e-3


@pytest.fixture(scope="module")
def tiny():
    from evennicer_slam_b200 import harness
    scene = cases.tiny_scene()
    decoders, c, renderer, cfg = harness.build(scene, DEV)
    cam_t, depth, color, event = cases.tiny_frame()
    return dict(scene=scene, decoders=decoders, c=c, renderer=renderer, cam_t=cam_t, depth=depth,
                color=color, g=load_golden("tiny_render.npz"), sc=orc.OracleScene.from_synthetic(scene))


def test_library_loads_and_reports_version():
    from evennicer_slam_b200 import _lib
    L = _lib.lib()
    assert L.ens_version() == _lib.ABI_VERSION
    assert L.ens_packed_decoder_floats(2) == 21028 + 22244 + 16800 + 37700 + 29216 and L.ens_decoder_grad_floats(3) == 15899


def test_grid_layout_roundtrip():
    from evennicer_slam_b200 import _lib
    L = _lib.lib()
    g = torch.randn(1, 32, 5, 7, 9, device=DEV)
    nat = torch.empty(5, 7, 9, 32, device=DEV)
    _lib.check(L.ens_grid_to_native(_lib.ptr(g), _lib.ptr(nat), 5 * 7 * 9, _lib.cur_stream(g.device)))
    assert torch.equal(nat, g[0].permute(1, 2, 3, 0).contiguous())
    back = torch.zeros_like(g)
    _lib.check(L.ens_grid_from_native(_lib.ptr(nat), _lib.ptr(back), 5 * 7 * 9, 0, _lib.cur_stream(g.device)))
    assert torch.equal(back, g)
    _lib.check(L.ens_grid_from_native(_lib.ptr(nat), _lib.ptr(back), 5 * 7 * 9, 1, _lib.cur_stream(g.device)))
    assert torch.equal(back, 2 * g)


def test_get_samples_bit_exact(tiny):
    """Pixel draw + gather + ray generation vs. the golden AND vs. the reference's own expression on the GPU."""
    from evennicer_slam_b200 import common
    scene, g = tiny["scene"], tiny["g"]
    cam = scene.cam
    c2w = torch.from_numpy(g["color.d.c2w"]).to(DEV)
    depth = torch.from_numpy(tiny["depth"]).to(DEV)
    color = torch.from_numpy(tiny["color"]).to(DEV)
    seed = int(g["color.d.seed"])
    torch.manual_seed(seed)
    gen_idx = torch.randint(cam.H * cam.W, (cases.N_TINY_RAYS,), device=DEV)
    torch.manual_seed(seed)
    ro, rd, sd, scol = common.get_samples(0, cam.H, 0, cam.W, cases.N_TINY_RAYS, cam.H, cam.W, cam.fx, cam.fy,
                                          cam.cx, cam.cy, c2w, depth, color, DEV)
    # same generator state -> same indices as an explicit randint on this device
    i, j, d_ref, c_ref = orc.select_pixels(gen_idx.cpu().numpy(), 0, cam.H, 0, cam.W, tiny["depth"], tiny["color"])
    assert np.array_equal(sd.cpu().numpy(), d_ref) and np.array_equal(scol.cpu().numpy(), c_ref)
    ro_ref, rd_ref = orc.rays_from_uv(i, j, g["color.d.c2w"], cam.fx, cam.fy, cam.cx, cam.cy)
    assert np.array_equal(rd.cpu().numpy(), rd_ref) and np.array_equal(ro.cpu().numpy(), ro_ref)
    # the reference's eager expression evaluated by torch on this GPU (common.py:80-88)
    ti, tj = torch.from_numpy(i).to(DEV), torch.from_numpy(j).to(DEV)
    dirs = torch.stack([(ti - cam.cx) / cam.fx, -(tj - cam.cy) / cam.fy, -torch.ones_like(ti)], -1).reshape(-1, 1, 3)
    rd_eager = torch.sum(dirs * c2w[:3, :3], -1)
    print("eager-CUDA rays identical:", bool(torch.equal(rd, rd_eager)), "max ulp-ish diff", float((rd - rd_eager).abs().max()))


def test_get_samples_event_shares_one_draw(tiny):
    """common.py:178-187: rays, depth, colour and both event images come from ONE torch.randint draw."""
    from evennicer_slam_b200 import common
    scene, g = tiny["scene"], tiny["g"]
    cam = scene.cam
    c2w = torch.from_numpy(g["color.d.c2w"]).to(DEV)
    depth = torch.from_numpy(tiny["depth"]).to(DEV)
    color = torch.from_numpy(tiny["color"]).to(DEV)
    gen = torch.Generator(device=DEV); gen.manual_seed(1)
    ev1 = torch.randn(cam.H, cam.W, 2, device=DEV, generator=gen)
    ev2 = torch.randn(cam.H, cam.W, 2, device=DEV, generator=gen)
    H0, H1, W0, W1 = 2, cam.H - 3, 1, cam.W - 2
    n = 77
    torch.manual_seed(9)
    idx = torch.randint((H1 - H0) * (W1 - W0), (n,), device=DEV).cpu().numpy()
    torch.manual_seed(9)
    ro, rd, sd, scol, se1, se2 = common.get_samples_event(H0, H1, W0, W1, n, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy,
                                                          c2w, depth, color, ev1, ev2, DEV)
    i, j, d_ref, c_ref = orc.select_pixels(idx, H0, H1, W0, W1, tiny["depth"], tiny["color"])
    ro_ref, rd_ref = orc.rays_from_uv(i, j, g["color.d.c2w"], cam.fx, cam.fy, cam.cx, cam.cy)
    assert np.array_equal(rd.cpu().numpy(), rd_ref) and np.array_equal(ro.cpu().numpy(), ro_ref)
    assert np.array_equal(sd.cpu().numpy(), d_ref) and np.array_equal(scol.cpu().numpy(), c_ref)
    e1 = ev1.cpu().numpy()[H0:H1, W0:W1].reshape(-1, 2)[idx]
    e2 = ev2.cpu().numpy()[H0:H1, W0:W1].reshape(-1, 2)[idx]
    assert np.array_equal(se1.cpu().numpy(), e1) and np.array_equal(se2.cpu().numpy(), e2)


def test_get_samples_indices_match_cpu_golden_when_generators_agree(tiny):
    """torch's CPU and CUDA generators differ, so golden indices (CPU) are checked by feeding them
    through the kernel rather than by re-drawing."""
    from evennicer_slam_b200 import _lib
    scene, g = tiny["scene"], tiny["g"]
    cam = scene.cam
    L = _lib.lib()
    idx = torch.from_numpy(g["color.d.indices"]).to(DEV)
    n = idx.numel()
    c2w = torch.from_numpy(g["color.d.c2w"]).to(DEV)
    depth = torch.from_numpy(tiny["depth"]).to(DEV)
    color = torch.from_numpy(tiny["color"]).to(DEV)
    ro = torch.empty(n, 3, device=DEV); rd = torch.empty(n, 3, device=DEV)
    sd = torch.empty(n, device=DEV); sc = torch.empty(n, 3, dtype=torch.float64, device=DEV)
    _lib.check(L.ens_sample_rays(_lib.ptr(idx), n, 0, cam.H, 0, cam.W, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy,
                                 _lib.ptr(c2w), c2w.stride(0), _lib.ptr(depth), _lib.ptr(color), 1, None, None,
                                 _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(sd), _lib.ptr(sc), _lib.cur_stream(idx.device)))
    assert np.array_equal(rd.cpu().numpy(), g["color.d.rays_d"])
    assert np.array_equal(ro.cpu().numpy(), g["color.d.rays_o"])
    assert np.array_equal(sd.cpu().numpy(), g["color.d.sample_depth"])
    assert np.array_equal(sc.cpu().numpy(), g["color.d.sample_color"])


@pytest.mark.parametrize("stage", STAGES)
def test_eval_points(tiny, stage):
    scene, renderer, decoders, c = tiny["scene"], tiny["renderer"], tiny["decoders"], tiny["c"]
    g = load_golden("tiny_eval_points.npz")
    pts = cases.eval_points_lattice(scene)
    out = renderer.eval_points(torch.from_numpy(pts).to(DEV), decoders, c, stage, DEV).cpu().numpy()
    ref = g[f"{stage}.f64"]
    assert np.array_equal(out[:, 3] == 100, ref[:, 3] == 100)
    assert rel_err(out, ref) < TOL_OUT
    oracle, _ = orc.eval_points(tiny["sc"], pts, stage)
    assert rel_err(out, oracle) < TOL_OUT
    out32 = renderer.eval_points(torch.from_numpy(pts.astype(np.float32)).to(DEV), decoders, c, stage, DEV).cpu().numpy()
    ref32 = g[f"{stage}.f32"]
    assert np.array_equal(out32[:, 3] == 100, ref32[:, 3] == 100)
    assert rel_err(out32, ref32) < TOL_OUT
    # NICE.forward: same decode without the bound rule (Mesher.py:308)
    bare = decoders(torch.from_numpy(pts).to(DEV)[None], c_grid=c, stage=stage).cpu().numpy()
    inside = ref[:, 3] != 100
    assert rel_err(bare[inside], ref[inside]) < TOL_OUT and not np.any(bare[:, 3] == 100)


@pytest.mark.parametrize("stage", ["middle", "fine", "color"])
def test_eval_points_variants(tiny, stage, monkeypatch):
    """Both forward-only decode variants against the reference goldens: the tcgen05/TMEM kernels (ens_decode_tc.cu,
    default) and the mma.sync kernels (ENS_EVAL_VARIANT=mma)."""
    for variant in ("tc", "mma"):
        monkeypatch.setenv("ENS_EVAL_VARIANT", variant)
        _check_eval_variant(tiny, stage)


def _check_eval_variant(tiny, stage):
    scene, renderer, decoders, c = tiny["scene"], tiny["renderer"], tiny["decoders"], tiny["c"]
    g = load_golden("tiny_eval_points.npz")
    pts = cases.eval_points_lattice(scene)
    for dt, key in ((np.float64, "f64"), (np.float32, "f32")):
        out = renderer.eval_points(torch.from_numpy(pts.astype(dt)).to(DEV), decoders, c, stage, DEV).cpu().numpy()
        ref = g[f"{stage}.{key}"]
        assert np.array_equal(out[:, 3] == 100, ref[:, 3] == 100)
        assert rel_err(out, ref) < TOL_OUT
    # ragged sizes: not a multiple of the 128-point tile, fewer tiles than SMs, one point
    for n in (1, 127, 129, 300):
        out = renderer.eval_points(torch.from_numpy(pts[:n]).to(DEV), decoders, c, stage, DEV).cpu().numpy()
        assert rel_err(out, g[f"{stage}.f64"][:n]) < TOL_OUT


def _run_case(tiny, stage, use_depth, with_params=True):
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    tag = f"{stage}.{'d' if use_depth else 'n'}"
    for p in decoders.parameters():
        p.grad = None
        p.requires_grad_(with_params)
    cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV).requires_grad_(True)
    rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV).requires_grad_(True)
    sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
    out = renderer.render_batch_ray_aux(cg, decoders, rd, ro, DEV, stage, gt_depth=sd if use_depth else None)
    return tag, cg, ro, rd, out


@pytest.mark.parametrize("stage", STAGES)
@pytest.mark.parametrize("use_depth", [True, False])
def test_render_forward(tiny, stage, use_depth):
    g = tiny["g"]
    tag, cg, ro, rd, (depth, var, color, raw, z, w) = _run_case(tiny, stage, use_depth)
    assert depth.dtype == torch.float64 and var.dtype == torch.float64 and color.dtype == torch.float32
    assert np.array_equal(z.cpu().numpy(), g[f"{tag}.z_vals"]), "sample placement must be bit-exact"
    assert rel_err(raw.cpu().numpy(), g[f"{tag}.raw"]) < TOL_OUT
    assert rel_err(depth.detach().cpu().numpy(), g[f"{tag}.depth"]) < TOL_OUT
    assert rel_err(var.detach().cpu().numpy(), g[f"{tag}.var"]) < TOL_OUT
    if stage == "color":
        assert rel_err(color.detach().cpu().numpy(), g[f"{tag}.color"]) < TOL_OUT
    else:
        assert float(color.abs().max()) == 0.0


@pytest.mark.parametrize("stage", STAGES)
@pytest.mark.parametrize("use_depth", [True, False])
def test_render_backward(tiny, stage, use_depth):
    g, decoders = tiny["g"], tiny["decoders"]
    tag, cg, ro, rd, (depth, var, color, raw, z, w) = _run_case(tiny, stage, use_depth)
    g_d, g_v, g_c = cases.upstream_grads(ro.shape[0])
    loss = (depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum() \
        + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()
    loss.backward()
    assert rel_err(ro.grad.cpu().numpy(), g[f"{tag}.g_rays_o"]) < TOL_GRAD
    assert rel_err(rd.grad.cpu().numpy(), g[f"{tag}.g_rays_d"]) < TOL_GRAD
    for name in orc.STAGE_DECODERS[stage]:
        gk = "grid_" + name
        assert rel_err(cg[gk].grad.cpu().numpy(), g[f"{tag}.ggrid.{gk}"]) < TOL_GRAD, gk
        for key, p in getattr(decoders, name + "_decoder").named_parameters():
            ref = g[f"{tag}.gdec.{name}.{key}"]
            got = p.grad.cpu().numpy()
            if np.abs(ref).max() == 0:
                assert np.abs(got).max() == 0, (name, key)
            else:
                assert rel_err(got, ref) < TOL_GRAD, (name, key)
    for lv in STAGES:   # untouched levels get no gradient
        if lv not in orc.STAGE_DECODERS[stage]:
            assert cg["grid_" + lv].grad is None


@pytest.mark.parametrize("stage,use_depth", [("color", True), ("fine", False), ("middle", True)])
def test_render_backward_recompute_path(tiny, stage, use_depth):
    """The backward kernels also run WITHOUT the forward's saved masks / activations (direct C-ABI callers that pass
    saved = NULL): they then recompute the forward in-kernel.  Same gradients, same tolerance."""
    from evennicer_slam_b200 import functional
    g, decoders = tiny["g"], tiny["decoders"]
    functional.SAVE_FORWARD = False
    try:
        tag, cg, ro, rd, (depth, var, color, raw, z, w) = _run_case(tiny, stage, use_depth)
        g_d, g_v, g_c = cases.upstream_grads(ro.shape[0])
        ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
         + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    finally:
        functional.SAVE_FORWARD = True
    assert rel_err(ro.grad.cpu().numpy(), g[f"{tag}.g_rays_o"]) < TOL_GRAD
    assert rel_err(rd.grad.cpu().numpy(), g[f"{tag}.g_rays_d"]) < TOL_GRAD
    for name in orc.STAGE_DECODERS[stage]:
        gk = "grid_" + name
        assert rel_err(cg[gk].grad.cpu().numpy(), g[f"{tag}.ggrid.{gk}"]) < TOL_GRAD, gk
        for key, p in getattr(decoders, name + "_decoder").named_parameters():
            ref = g[f"{tag}.gdec.{name}.{key}"]
            if np.abs(ref).max() > 0:
                assert rel_err(p.grad.cpu().numpy(), ref) < TOL_GRAD, (name, key)


@pytest.mark.parametrize("n_rays", [1, 5, 13, 37])
@pytest.mark.parametrize("stage,use_depth", [("color", True), ("fine", False)])
def test_ragged_ray_counts_vs_oracle(tiny, n_rays, stage, use_depth):
    """Ray counts that fill neither a forward CTA (8 rays) nor a backward CTA (4 rays) nor a 32-point tile: idle lanes
    must not contribute, and the forward's saved tiles must line up with the backward's.  Subsets of a kink-safe golden
    case are kink-safe, so gradients are held to 1e-3 against the oracle run on the same subset."""
    renderer, decoders, c, g, sc = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"], tiny["sc"]
    tag = f"{stage}.{'d' if use_depth else 'n'}"
    for p in decoders.parameters():
        p.grad = None
        p.requires_grad_(True)
    cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    ro_np, rd_np, sd_np = g[f"{tag}.rays_o"][:n_rays], g[f"{tag}.rays_d"][:n_rays], g[f"{tag}.sample_depth"][:n_rays]
    ro = torch.from_numpy(ro_np.copy()).to(DEV).requires_grad_(True)
    rd = torch.from_numpy(rd_np.copy()).to(DEV).requires_grad_(True)
    sd = torch.from_numpy(sd_np.copy()).to(DEV)
    depth, var, color, raw, z, w = renderer.render_batch_ray_aux(cg, decoders, rd, ro, DEV, stage,
                                                                gt_depth=sd if use_depth else None)
    g_d, g_v, g_c = cases.upstream_grads(n_rays)
    ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
     + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    t32 = torch.linspace(0., 1., 32).numpy()
    t64 = torch.linspace(0., 1., 16).double().numpy()
    od, ov, oc, cache = orc.render_batch_ray(sc, ro_np, rd_np, stage, sd_np if use_depth else None, t32, t64)
    og = orc.render_batch_ray_backward(sc, cache, g_d, g_v, g_c)
    assert np.array_equal(z.cpu().numpy(), cache["z"])
    assert rel_err(depth.detach().cpu().numpy(), od) < TOL_OUT and rel_err(var.detach().cpu().numpy(), ov) < TOL_OUT
    if stage == "color":
        assert rel_err(color.detach().cpu().numpy(), oc) < TOL_OUT
    assert rel_err(ro.grad.cpu().numpy(), og["rays_o"]) < TOL_GRAD
    assert rel_err(rd.grad.cpu().numpy(), og["rays_d"]) < TOL_GRAD
    for name in orc.STAGE_DECODERS[stage]:
        gk = "grid_" + name
        assert rel_err(cg[gk].grad.cpu().numpy(), og["grids"][gk]) < TOL_GRAD, gk
        for key, p in getattr(decoders, name + "_decoder").named_parameters():
            ref = og["decoders"][name][key]
            if np.abs(ref).max() > 0:
                assert rel_err(p.grad.cpu().numpy(), ref) < TOL_GRAD, (name, key)


def test_empty_batch_is_a_no_op(tiny):
    renderer, decoders, c = tiny["renderer"], tiny["decoders"], tiny["c"]
    ro = torch.empty(0, 3, device=DEV); rd = torch.empty(0, 3, device=DEV); sd = torch.empty(0, device=DEV)
    d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=None)
    assert d.shape == (0,) and u.shape == (0,) and col.shape == (0, 3)
    out = renderer.eval_points(torch.empty(0, 3, device=DEV, dtype=torch.float64), decoders, c, "fine", DEV)
    assert out.shape == (0, 4)


def test_forward_is_independent_of_batch_boundaries():
    """Size-independent property at full scale (room0, 20 000 rays): one call equals the concatenation of ragged
    chunks BIT FOR BIT when the chunks are given the whole batch's depth maxima (what sharding.py passes), and
    eval_points of 300 000 lattice points equals its chunked evaluation."""
    from evennicer_slam_b200 import harness, functional, common
    scene = cases.room0_scene()
    decoders, c, renderer, cfg = harness.build(scene, DEV, requires_grad=False)
    cam_t, depth, color, event = cases.room0_frame()
    cam = scene.cam
    torch.manual_seed(3)
    c2w = common.get_camera_from_tensor(torch.from_numpy(cam_t.copy()).to(DEV))
    ro, rd, sd, _ = common.get_samples(0, cam.H, 0, cam.W, 20000, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w,
                                       torch.from_numpy(depth).to(DEV), torch.from_numpy(color).to(DEV), DEV)
    setup = renderer._setup("color", decoders, DEV)
    dmax = functional.depth_batch_max(sd.contiguous())
    full = functional.render_batch_ray(setup, c, decoders, rd, ro, sd, depth_max=dmax)
    parts = []
    for lo, hi in ((0, 1), (1, 4098), (4098, 4099), (4099, 20000)):
        parts.append(functional.render_batch_ray(setup, c, decoders, rd[lo:hi], ro[lo:hi], sd[lo:hi], depth_max=dmax))
    for k in range(3):
        assert torch.equal(full[k], torch.cat([p[k] for p in parts])), k
    assert torch.isfinite(full[0]).all() and torch.isfinite(full[2]).all()
    b = scene.bound
    pts = torch.rand(300000, 3, device=DEV, dtype=torch.float64) * torch.tensor(b[:, 1] - b[:, 0] + 0.4, device=DEV) \
        + torch.tensor(b[:, 0] - 0.2, device=DEV)
    whole = renderer.eval_points(pts, decoders, c, "color", DEV)
    chunks = torch.cat([renderer.eval_points(pts[a:e], decoders, c, "color", DEV)
                        for a, e in ((0, 7), (7, 100001), (100001, 300000))])
    assert torch.equal(whole, chunks)
    assert int((whole[:, 3] == 100).sum()) > 0 and int((whole[:, 3] != 100).sum()) > 0


def test_backward_without_decoder_or_grid_grads_matches(tiny):
    """Tracker-style call: only the rays need gradient (decoder/grid grads skipped in the kernel)."""
    g = tiny["g"]
    renderer, decoders, c = tiny["renderer"], tiny["decoders"], tiny["c"]
    for p in decoders.parameters():
        p.requires_grad_(False)
    tag = "color.d"
    ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV).requires_grad_(True)
    rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV).requires_grad_(True)
    sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
    depth, var, color = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=sd)
    g_d, g_v, g_c = cases.upstream_grads(ro.shape[0])
    ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
     + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    assert rel_err(ro.grad.cpu().numpy(), g[f"{tag}.g_rays_o"]) < TOL_GRAD
    assert rel_err(rd.grad.cpu().numpy(), g[f"{tag}.g_rays_d"]) < TOL_GRAD
    for p in decoders.parameters():
        p.requires_grad_(True)


def test_pose_gradient_end_to_end(tiny):
    """camera_tensor -> get_camera_from_tensor -> get_samples -> render -> loss -> backward (Tracker.py:142-197)."""
    from evennicer_slam_b200 import common, _lib
    from evennicer_slam_b200.functional import _PairRays
    scene, g = tiny["scene"], tiny["g"]
    renderer, decoders, c = tiny["renderer"], tiny["decoders"], tiny["c"]
    cam = scene.cam
    ct = torch.from_numpy(tiny["cam_t"].copy()).to(DEV).requires_grad_(True)
    c2w = common.get_camera_from_tensor(ct)
    assert np.allclose(c2w.detach().cpu().numpy(), g["color.d.c2w"], atol=1e-6)
    idx = g["color.d.indices"]
    i, j, sd, _ = orc.select_pixels(idx, 0, cam.H, 0, cam.W, tiny["depth"], tiny["color"])
    ro, rd = _PairRays.apply(c2w, torch.from_numpy(i).to(DEV), torch.from_numpy(j).to(DEV),
                             (cam.H, cam.W, float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy)))
    depth, var, color = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=torch.from_numpy(sd).to(DEV))
    g_d, g_v, g_c = cases.upstream_grads(ro.shape[0])
    ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
     + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    assert rel_err(ct.grad.cpu().numpy(), g["color.d.g_cam"]) < TOL_GRAD


def test_full_frame_and_rescaled_frame(tiny):
    scene, renderer, decoders, c = tiny["scene"], tiny["renderer"], tiny["decoders"], tiny["c"]
    from evennicer_slam_b200 import common
    g = load_golden("tiny_frames.npz")
    cam = scene.cam
    ct = torch.from_numpy(tiny["cam_t"].copy()).to(DEV).requires_grad_(True)
    c2w = common.get_camera_from_tensor(ct)
    ro, rd = common.get_rays(cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w.detach(), DEV)
    assert np.array_equal(rd.cpu().numpy(), g["full.rays_d"]) and np.array_equal(ro.cpu().numpy(), g["full.rays_o"])
    nH, nW = int(cam.H * 0.5), int(cam.W * 0.5)
    ro2, rd2 = common.get_rays_rescale(cam.H, cam.W, nH, nW, cam.fx, cam.fy, cam.cx, cam.cy, c2w, DEV)
    assert np.array_equal(rd2.detach().cpu().numpy(), g["rescale.rays_d"])
    import evennicer_slam_b200.synthetic as syn
    gsum = torch.from_numpy(syn.det_uniform((nH, nW, 3), 555).astype(np.float32)).to(DEV)
    ((rd2 * gsum).sum() + (ro2 * gsum * 0.5).sum()).backward()
    assert rel_err(ct.grad.cpu().numpy(), g["rescale.g_cam"]) < 1e-5
    depth_img = torch.from_numpy(tiny["depth"]).to(DEV)
    renderer.ray_batch_size = 1000          # several batches -> per-batch depth maxima, as in the reference...
    d, u, col = renderer.render_img(c, decoders, c2w.detach(), DEV, "color", gt_depth=depth_img)
    renderer.ray_batch_size = 100000
    # ...but the golden used one 100k batch (3072 rays), so compare with a single batch too
    d1, u1, col1 = renderer.render_img(c, decoders, c2w.detach(), DEV, "color", gt_depth=depth_img)
    assert d1.dtype == torch.float64 and d1.shape == (cam.H, cam.W) and col1.shape == (cam.H, cam.W, 3)
    assert rel_err(d1.cpu().numpy(), g["render_img.depth"]) < TOL_OUT
    assert rel_err(u1.cpu().numpy(), g["render_img.var"]) < TOL_OUT
    assert rel_err(col1.cpu().numpy(), g["render_img.color"]) < TOL_OUT
    assert d.shape == d1.shape
    ct2 = torch.from_numpy(tiny["cam_t"].copy()).to(DEV).requires_grad_(True)
    c2w2 = common.get_camera_from_tensor(ct2)
    for p in decoders.parameters():
        p.requires_grad_(False)
    d, u, col = renderer.render_img_rescale(c, decoders, c2w2, DEV, "color", gt_depth=depth_img, scale_factor=0.5)
    assert rel_err(d.detach().cpu().numpy(), g["render_img_rescale.depth"]) < TOL_OUT
    assert rel_err(col.detach().cpu().numpy(), g["render_img_rescale.color"]) < TOL_OUT
    (col * gsum).sum().backward()
    assert rel_err(ct2.grad.cpu().numpy(), g["render_img_rescale.g_cam"]) < TOL_GRAD
    for p in decoders.parameters():
        p.requires_grad_(True)


def test_room0_mapping_batch_vs_golden_and_oracle():
    """Config C1/C3 shape: 1000 rays x 48 samples, colour stage, room0 grids (46 MiB)."""
    from evennicer_slam_b200 import harness
    scene = cases.room0_scene()
    decoders, c, renderer, cfg = harness.build(scene, DEV)
    g = load_golden("room0_color_1000.npz")
    cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    ro = torch.from_numpy(g["rays_o"]).to(DEV).requires_grad_(True)
    rd = torch.from_numpy(g["rays_d"]).to(DEV).requires_grad_(True)
    sd = torch.from_numpy(g["sample_depth"]).to(DEV)
    depth, var, color, raw, z, w = renderer.render_batch_ray_aux(cg, decoders, rd, ro, DEV, "color", gt_depth=sd)
    assert np.array_equal(z.cpu().numpy(), g["z_vals"])
    assert rel_err(depth.detach().cpu().numpy(), g["depth"]) < TOL_OUT
    assert rel_err(var.detach().cpu().numpy(), g["var"]) < TOL_OUT
    assert rel_err(color.detach().cpu().numpy(), g["color"]) < TOL_OUT
    g_d, g_v, g_c = cases.upstream_grads(cases.N_ROOM0_RAYS)
    ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
     + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    # Ray gradients.  The batch holds 24 M relu pre-activations, ~250 of them within 1e-5 of the kink, where the
    # reference's own CPU and GPU runs may pick different sides (cases.TINY_RELU_MARGIN); one flipped unit moves
    # ITS ray's gradient by up to a percent and nothing else.  What the callers consume is the pose gradient, the
    # sum over rays: that is held to 1e-3; per ray, all but a few rays must meet 1e-3 and none may be far off.
    cam = scene.cam
    i, j, _, _ = orc.select_pixels(g["indices"], 0, cam.H, 0, cam.W, *cases.room0_frame()[1:3])
    gc2w = orc.rays_from_uv_backward(i, j, cam.fx, cam.fy, cam.cx, cam.cy, ro.grad.cpu().numpy(), rd.grad.cpu().numpy())
    gc2w_ref = orc.rays_from_uv_backward(i, j, cam.fx, cam.fy, cam.cx, cam.cy, g["g_rays_o"], g["g_rays_d"])
    assert rel_err(gc2w, gc2w_ref) < TOL_GRAD
    for got, ref in ((ro.grad.cpu().numpy(), g["g_rays_o"]), (rd.grad.cpu().numpy(), g["g_rays_d"])):
        per_ray = np.abs(got - ref).max(axis=1) / np.abs(ref).max()
        assert (per_ray < TOL_GRAD).mean() >= 0.99, float((per_ray < TOL_GRAD).mean())
        assert per_ray.max() < 2e-2, float(per_ray.max())
    for name in ("fine", "color", "middle"):
        for key, p in getattr(decoders, name + "_decoder").named_parameters():
            ref = g[f"gdec.{name}.{key}"]
            if np.abs(ref).max() > 0:
                # gradient mass sits in the few samples per ray where alpha(1-alpha) is not ~0, so ONE relu decision
                # taken on the other side of the kink (|u| < 1e-6, within fp32 rounding of zero) shows as a few 1e-4
                # in every tensor below that layer: measured 1.9e-4 for the serial-fp32 FFMA kernels and 1.2e-3 for
                # the tensor-core kernels vs. the CPU reference (tools/diag_room0.py), all other tensors 1e-6.
                # Kink-free cases (tiny goldens, cases.TINY_RELU_MARGIN) are held to TOL_GRAD everywhere.
                assert rel_err(p.grad.cpu().numpy(), ref) < TOL_GRAD_KINK, (name, key)
        gk = "grid_" + name
        flat = cg[gk].grad.reshape(-1).cpu().numpy()
        assert rel_err(flat[g[f"ggrid.{gk}.probe_idx"]], g[f"ggrid.{gk}.probe_val"]) < TOL_GRAD
        assert abs(np.abs(flat.astype(np.float64)).sum() - g[f"ggrid.{gk}.l1"]) < TOL_GRAD * g[f"ggrid.{gk}.l1"]
        assert int((flat != 0).sum()) == int(g[f"ggrid.{gk}.nnz"])


def test_native_layout_grids_are_equivalent(tiny):
    """Grids re-allocated with scene.as_native_layout (a [1,32,Z,Y,X] VIEW of the kernels' [Z,Y,X,32] storage) give
    bit-identical outputs, the same gradients, need no layout conversion, and keep working with the callers' idioms
    (boolean-mask indexing and in-place masked writes, Mapper.py:343-361, 451-458)."""
    from evennicer_slam_b200.scene import as_native_layout, is_native_strided
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    tag = "color.d"
    for p in decoders.parameters():
        p.requires_grad_(True)
    outs = []
    for native in (False, True):
        for p in decoders.parameters():
            p.grad = None
        cg = {k: (as_native_layout(v.clone()) if native else v.clone()).requires_grad_(True) for k, v in c.items()}
        if native:
            assert all(is_native_strided(v) and v.shape == c[k].shape and torch.equal(v, c[k]) for k, v in cg.items())
        ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV).requires_grad_(True)
        rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV).requires_grad_(True)
        sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
        before = dict(renderer._cache.stats)
        depth, var, color = renderer.render_batch_ray(cg, decoders, rd, ro, DEV, "color", gt_depth=sd)
        if native:
            assert renderer._cache.stats["grid_convert"] == before["grid_convert"]      # nothing was re-laid-out
        g_d, g_v, g_c = cases.upstream_grads(ro.shape[0])
        ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
         + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
        outs.append((depth.detach(), color.detach(), {k: v.grad for k, v in cg.items()}, rd.grad))
    (d0, c0, gg0, r0), (d1, c1, gg1, r1) = outs
    assert torch.equal(d0, d1) and torch.equal(c0, c1)
    for k in ("grid_middle", "grid_fine", "grid_color"):
        assert gg1[k].shape == gg0[k].shape
        assert rel_err(gg1[k].cpu().numpy(), gg0[k].cpu().numpy()) < 1e-5
        assert rel_err(gg1[k].cpu().numpy(), g[f"{tag}.ggrid.{k}"]) < TOL_GRAD
    assert rel_err(r1.cpu().numpy(), r0.cpu().numpy()) < 1e-5
    # the Mapper's masked read / write idiom on a native-layout grid
    val = as_native_layout(c["grid_fine"].clone())
    mask = torch.rand(val.shape, device=DEV) > 0.5
    sel = val[mask].clone()
    assert torch.equal(sel, c["grid_fine"][mask])
    val[mask] = sel * 2
    assert torch.equal(val[mask], c["grid_fine"][mask] * 2) and is_native_strided(val)


def test_rpg_sized_grids_larger_than_l2():
    """RPG recording4 bounds with the 346 x 260 camera (configs/rpg/rpg.yaml:62-71): 197 MiB of grids, more than the
    126 MB L2.  A 200-ray colour-stage batch forward + backward against the oracle (per-ray quantities robustly: the
    case is not kink-screened), plus sanity on the dense grid gradient."""
    import evennicer_slam_b200.synthetic as syn
    from evennicer_slam_b200 import harness, common
    scene = syn.make_scene(syn.RPG4_BOUND, syn.RPG_CAM, seed=cases.SEED, name="rpg4", grid_std={"fine": 0.01})
    assert sum(v.nbytes for v in scene.grids.values()) > 190 << 20
    decoders, c, renderer, cfg = harness.build(scene, DEV)
    cam = scene.cam
    cam_t = syn.default_pose(syn.RPG4_BOUND, jitter_seed=3)
    depth, color, _ = syn.synthetic_frame(syn.RPG4_BOUND, cam, cam_t, seed=5, zero_frac=0.02)
    torch.manual_seed(11)
    c2w = common.get_camera_from_tensor(torch.from_numpy(cam_t.copy()).to(DEV))
    ro, rd, sd, _ = common.get_samples(0, cam.H, 0, cam.W, 200, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w,
                                       torch.from_numpy(depth).to(DEV), torch.from_numpy(color).to(DEV), DEV)
    ro = ro.detach().requires_grad_(True); rd = rd.detach().requires_grad_(True)
    cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    d, u, col, raw, z, w = renderer.render_batch_ray_aux(cg, decoders, rd, ro, DEV, "color", gt_depth=sd)
    g_d, g_v, g_c = cases.upstream_grads(200)
    ((d * torch.from_numpy(g_d).to(DEV)).sum() + (u * torch.from_numpy(g_v).to(DEV)).sum()
     + (col.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    sc = orc.OracleScene.from_synthetic(scene)
    t32 = torch.linspace(0., 1., 32).numpy(); t64 = torch.linspace(0., 1., 16).double().numpy()
    od, ov, oc, cache = orc.render_batch_ray(sc, ro.detach().cpu().numpy(), rd.detach().cpu().numpy(), "color",
                                             sd.cpu().numpy(), t32, t64)
    assert np.array_equal(z.cpu().numpy(), cache["z"])
    assert rel_err(d.detach().cpu().numpy(), od) < TOL_OUT and rel_err(col.detach().cpu().numpy(), oc) < TOL_OUT
    og = orc.render_batch_ray_backward(sc, cache, g_d, g_v, g_c)
    for got, ref in ((ro.grad.cpu().numpy(), og["rays_o"]), (rd.grad.cpu().numpy(), og["rays_d"])):
        per_ray = np.abs(got - ref).max(axis=1) / np.abs(ref).max()
        assert (per_ray < TOL_GRAD).mean() >= 0.98 and per_ray.max() < 2e-2
    for k in ("grid_middle", "grid_fine", "grid_color"):
        gg = cg[k].grad.cpu().numpy()
        # 200 rays, not kink-screened, against the ORACLE's relu decisions: one unit on the other side of the kink is 1/9600 of
        # the batch here (5x its weight in the 1000-ray case, which meets 1e-3 outright) -- the one place that keeps 3e-3
        assert rel_err(gg, og["grids"][k]) < 3e-3, k
        assert int((gg != 0).sum()) == int((og["grids"][k] != 0).sum()), k


def test_scene_cache_tracks_in_place_updates(tiny):
    """Mapper mutates grids in place every iteration (Mapper.py:451-458): results must follow."""
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    tag = "middle.d"
    ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV)
    rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV)
    sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
    c2 = {k: v.clone() for k, v in c.items()}
    with torch.no_grad():
        d0, _, _ = renderer.render_batch_ray(c2, decoders, rd, ro, DEV, "middle", gt_depth=sd)
        hits = renderer._cache.stats["grid_hit"]
        d1, _, _ = renderer.render_batch_ray(c2, decoders, rd, ro, DEV, "middle", gt_depth=sd)
        assert renderer._cache.stats["grid_hit"] == hits + 1 and torch.equal(d0, d1)
        c2["grid_middle"].mul_(0.5)
        d2, _, _ = renderer.render_batch_ray(c2, decoders, rd, ro, DEV, "middle", gt_depth=sd)
        assert not torch.equal(d0, d2)
        c2["grid_middle"].mul_(2.0)
        d3, _, _ = renderer.render_batch_ray(c2, decoders, rd, ro, DEV, "middle", gt_depth=sd)
        assert torch.equal(d0, d3)


def test_errors_are_reported_not_fatal(tiny):
    from evennicer_slam_b200 import _lib
    L = _lib.lib()
    assert L.ens_grid_to_native(None, None, 10, None) == -1
    assert b"unsupported" in L.ens_strerror(-5)
    renderer, decoders, c = tiny["renderer"], tiny["decoders"], tiny["c"]
    renderer.N_importance = 12
    try:
        with pytest.raises(RuntimeError, match="unsupported"):
            renderer.render_batch_ray(c, decoders, torch.zeros(4, 3, device=DEV), torch.zeros(4, 3, device=DEV), DEV, "color")
    finally:
        renderer.N_importance = 0


def test_fused_pose_kernel_matches_eager_quaternion_math():
    from evennicer_slam_b200 import common
    torch.manual_seed(3)
    cam = torch.randn(6, 7, device=DEV)
    cam[:, :4] += torch.tensor([1.5, 0, 0, 0], device=DEV)       # un-normalised quaternions
    a = cam.clone().requires_grad_(True)
    b = cam.clone().requires_grad_(True)
    fused = common.get_camera_from_tensor(a)
    eager = common.get_camera_from_tensor_torch(b)
    assert torch.equal(fused, eager), float((fused - eager).abs().max())
    g = torch.randn_like(fused)
    (fused * g).sum().backward()
    (eager * g).sum().backward()
    assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-5
    one = common.get_camera_from_tensor(cam[0])
    assert one.shape == (3, 4) and torch.equal(one, eager[0].detach())


def test_graphed_mapping_step_equals_eager(tiny):
    """A captured + replayed step produces the same gradients as the eager step."""
    from evennicer_slam_b200 import common
    from evennicer_slam_b200.graph import GraphedStep
    scene, renderer, decoders, c = tiny["scene"], tiny["renderer"], tiny["decoders"], tiny["c"]
    cam = scene.cam
    depth = torch.from_numpy(tiny["depth"]).to(DEV)
    color = torch.from_numpy(tiny["color"]).to(DEV)
    ct = torch.from_numpy(tiny["cam_t"].copy()).to(DEV).requires_grad_(True)
    cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    for p in decoders.parameters():
        p.requires_grad_(True)
    idx = torch.from_numpy(tiny["g"]["color.d.indices"]).to(DEV)

    def step():
        renderer._cache.invalidate()
        c2w = common.get_camera_from_tensor(ct)
        from evennicer_slam_b200.functional import _SampleRays
        ro, rd, sd, sc = _SampleRays.apply(c2w, idx, (0, cam.H, 0, cam.W),
                                           (cam.H, cam.W, float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy)),
                                           depth, color)
        d, u, col = renderer.render_batch_ray(cg, decoders, rd, ro, DEV, "color", gt_depth=sd)
        loss = torch.where(sd > 0, torch.abs(sd - d), 0.0).sum() + 0.2 * torch.abs(sc.float() - col).sum()
        loss.backward()
        return loss

    def grads():
        return [ct.grad.clone()] + [cg[k].grad.clone() for k in sorted(cg) if cg[k].grad is not None] + \
               [p.grad.clone() for p in decoders.parameters() if p.grad is not None]

    def zero():
        ct.grad = None
        for v in cg.values():
            v.grad = None
        for p in decoders.parameters():
            p.grad = None

    zero(); l0 = float(step()); torch.cuda.synchronize(); ref = grads()   # float(): do not keep the autograd graph alive
    zero()
    gs = GraphedStep(step, warmup=2, device=DEV)
    for _ in range(2):
        for t in [ct] + list(cg.values()) + list(decoders.parameters()):
            if t.grad is not None:
                t.grad.zero_()
        l1 = gs()
    torch.cuda.synchronize()
    got = grads()
    assert abs(l0 - float(l1)) < 1e-6 * abs(l0)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-4


def test_tcgen05_forward_feeds_the_pose_only_backward(tiny):
    """A backward without decoder gradients (tracking, the tracker's event render) keeps only relu masks; they then come
    from the tcgen05 decode path as one word per point (saved kind 2).  Same outputs and ray gradients as the mma.sync
    forward with lane-layout masks, for every stage that has the path, and the oracle's gradients."""
    from evennicer_slam_b200 import functional
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    req = [p.requires_grad for p in decoders.parameters()]
    for p in decoders.parameters():
        p.requires_grad_(False)
    try:
        for stage in ("middle", "fine", "color"):
            tag = f"{stage}.d"
            sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
            n = sd.shape[0]
            g_d, g_v, g_c = cases.upstream_grads(n)
            res = {}
            for use_tc in (True, False):
                functional.TC_POSE_FORWARD = use_tc
                ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV).requires_grad_(True)
                rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV).requires_grad_(True)
                d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, DEV, stage, gt_depth=sd)
                ((d * torch.from_numpy(g_d).to(DEV)).sum() + (u * torch.from_numpy(g_v).to(DEV)).sum()
                 + (col.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
                res[use_tc] = [x.detach().cpu().numpy() for x in (d, u, col, ro.grad, rd.grad)]
            for a, b in zip(res[True], res[False]):
                assert rel_err(a, b) < TOL_OUT
            assert rel_err(res[True][0], g[f"{tag}.depth"]) < TOL_OUT
            assert rel_err(res[True][3], g[f"{tag}.g_rays_o"]) < TOL_GRAD and rel_err(res[True][4], g[f"{tag}.g_rays_d"]) < TOL_GRAD
    finally:
        functional.TC_POSE_FORWARD = True
        for p, r in zip(decoders.parameters(), req):
            p.requires_grad_(r)


def test_decoder_parallel_backward_matches_sequential(tiny):
    """Small pose-only batches run the backward with one CTA per (ray group, decoder) and accumulate the ray gradients
    atomically (ENS_BWD_DECODER_PARALLEL, default on); same gradients as the one-CTA-walks-all-decoders form, also into
    the grids when they ask for gradient."""
    import os
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    req = [p.requires_grad for p in decoders.parameters()]
    for p in decoders.parameters():
        p.requires_grad_(False)
    try:
        for stage in ("fine", "color"):
            tag = f"{stage}.d"
            sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
            g_d, g_v, g_c = cases.upstream_grads(sd.shape[0])
            res = {}
            for mode in ("1", "0"):
                os.environ["ENS_BWD_DECODER_PARALLEL"] = mode
                ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV).requires_grad_(True)
                rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV).requires_grad_(True)
                cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
                d, u, col = renderer.render_batch_ray(cg, decoders, rd, ro, DEV, stage, gt_depth=sd)
                ((d * torch.from_numpy(g_d).to(DEV)).sum() + (u * torch.from_numpy(g_v).to(DEV)).sum()
                 + (col.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
                res[mode] = [ro.grad.cpu().numpy(), rd.grad.cpu().numpy()] + \
                            [cg[k].grad.cpu().numpy() for k in sorted(cg) if cg[k].grad is not None]
            assert len(res["1"]) == len(res["0"]) >= 3
            for a, b in zip(res["1"], res["0"]):
                assert rel_err(a, b) < 1e-5
            assert rel_err(res["1"][0], g[f"{tag}.g_rays_o"]) < TOL_GRAD
    finally:
        os.environ.pop("ENS_BWD_DECODER_PARALLEL", None)
        for p, r in zip(decoders.parameters(), req):
            p.requires_grad_(r)


def test_multi_decoder_tcgen05_launch_matches_sequential(tiny):
    """Small batches run all decoders of the stage in ONE tcgen05 launch (blockIdx.y = decoder, separate output planes
    combined by the compositing kernel; ENS_TC_MULTI, default on).  Bit-identical outputs, raw values and saved masks
    (same gradients) as one launch per decoder."""
    import os
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    req = [p.requires_grad for p in decoders.parameters()]
    for p in decoders.parameters():
        p.requires_grad_(False)
    try:
        for stage in ("fine", "color"):
            tag = f"{stage}.d"
            sd = torch.from_numpy(g[f"{tag}.sample_depth"]).to(DEV)
            res = {}
            for mode in ("1", "0"):
                os.environ["ENS_TC_MULTI"] = mode
                ro = torch.from_numpy(g[f"{tag}.rays_o"]).to(DEV).requires_grad_(True)
                rd = torch.from_numpy(g[f"{tag}.rays_d"]).to(DEV).requires_grad_(True)
                d, u, col, raw, z, w = renderer.render_batch_ray_aux(c, decoders, rd, ro, DEV, stage, gt_depth=sd)
                (d.sum() + col.double().sum()).backward()
                res[mode] = [x.detach().cpu().numpy() for x in (d, u, col, raw, w, ro.grad, rd.grad)]
                # forward only (no saved masks): the same path without a backward
                with torch.no_grad():
                    d2, u2, col2 = renderer.render_batch_ray(c, decoders, rd.detach(), ro.detach(), DEV, stage, gt_depth=sd)
                res[mode] += [d2.cpu().numpy(), col2.cpu().numpy()]
            for k, (a, b) in enumerate(zip(res["1"], res["0"])):
                if k in (5, 6):
                    assert rel_err(a, b) < 1e-5           # ray gradients: atomic accumulation order
                else:
                    assert np.array_equal(a, b), (stage, k)
            assert rel_err(res["1"][0], g[f"{tag}.depth"]) < TOL_OUT
    finally:
        os.environ.pop("ENS_TC_MULTI", None)
        for p, r in zip(decoders.parameters(), req):
            p.requires_grad_(r)


def test_small_batch_compositing_kernel_is_bit_identical(tiny):
    """Batches of up to 4096 rays composite with one warp per ray, larger ones with one thread per ray; the sequential
    products and sums run in the same order, so a ray renders to the same bits either way (no gt_depth: no batch-global
    depth maxima couple the rays)."""
    renderer, decoders, c, g = tiny["renderer"], tiny["decoders"], tiny["c"], tiny["g"]
    ro = torch.from_numpy(g["color.d.rays_o"]).to(DEV)
    rd = torch.from_numpy(g["color.d.rays_d"]).to(DEV)
    n = ro.shape[0]
    reps = 4200 // n + 1
    big_o, big_d = ro.repeat(reps, 1), rd.repeat(reps, 1)
    assert big_o.shape[0] > 4096 >= n
    with torch.no_grad():
        for stage in ("middle", "color"):
            d1, u1, c1 = renderer.render_batch_ray(c, decoders, rd, ro, DEV, stage, gt_depth=None)
            d2, u2, c2 = renderer.render_batch_ray(c, decoders, big_d, big_o, DEV, stage, gt_depth=None)
            assert torch.equal(d1, d2[:n]) and torch.equal(u1, u2[:n]) and torch.equal(c1, c2[:n])
            assert torch.equal(d2[:n], d2[-n:])


def _kernel_relu_masks(saved, kind, P, ndec):
    """relu decisions of the tcgen05 forward (saved kind 3: [relu outputs | mask words]) -> bool [ndec][5][P][32]"""
    assert kind == 3
    nt128 = (P + 127) // 128
    words = saved.view(torch.int32)[ndec * nt128 * 5 * 128 * 32:].cpu().numpy().view(np.uint32)
    nt32 = 4 * nt128
    words = words[:ndec * nt32 * 160].reshape(ndec, nt32, 5, 32)
    w = words.transpose(0, 2, 1, 3).reshape(ndec, 5, nt32 * 32)[:, :, :P]          # [dec][block][point]
    return ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)


def test_room0_gradients_meet_1e3_with_the_kernels_own_relu_decisions():
    """The headline batch (1000 rays x 48 samples, colour stage, room0) at the north-star tolerance WITHOUT a relaxation.

    24 M relu pre-activations: a few hundred lie within float32 rounding of zero, where two correct implementations (torch
    CPU, cuBLAS, these kernels) may take different sides; each such flip moves the gradients by the whole contribution of
    one unit at one point.  Here the oracle's backward is run with the relu decisions the KERNELS took (the mask words the
    tcgen05 forward saved): every decision that differs from the oracle's own is counted and must be a unit with
    |pre-activation| < 1e-5; with the same decisions, every gradient -- all 69 decoder tensors, the three grids, the
    rays -- meets 1e-3, as max-norm AND elementwise (tests/util.elem_err)."""
    from evennicer_slam_b200 import harness, functional
    from util import elem_err
    scene = cases.room0_scene()
    decoders, c, renderer, cfg = harness.build(scene, DEV)
    g = load_golden("room0_color_1000.npz")
    cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    ro = torch.from_numpy(g["rays_o"]).to(DEV).requires_grad_(True)
    rd = torch.from_numpy(g["rays_d"]).to(DEV).requires_grad_(True)
    sd = torch.from_numpy(g["sample_depth"]).to(DEV)
    functional._DEBUG["keep_saved"] = True
    try:
        depth, var, color = renderer.render_batch_ray(cg, decoders, rd, ro, DEV, "color", gt_depth=sd)
        saved, kind = functional._DEBUG.pop("saved")
    finally:
        functional._DEBUG.pop("keep_saved", None)
    if kind != 3:
        pytest.skip("the tcgen05 mapping path is switched off (ENS_MAP_TC=0)")
    g_d, g_v, g_c = cases.upstream_grads(cases.N_ROOM0_RAYS)
    ((depth * torch.from_numpy(g_d).to(DEV)).sum() + (var * torch.from_numpy(g_v).to(DEV)).sum()
     + (color.double() * torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    sc = orc.OracleScene.from_synthetic(scene)
    t32 = torch.linspace(0., 1., 32).numpy(); t64 = torch.linspace(0., 1., 16).double().numpy()
    od, ov, oc, cache = orc.render_batch_ray(sc, g["rays_o"], g["rays_d"], "color", g["sample_depth"], t32, t64)
    P = cache["R"] * cache["S"]
    masks = _kernel_relu_masks(saved, kind, P, 3)
    flips = 0
    for d, name in enumerate(("middle", "fine", "color")):
        acts = cache["caches"][name]["mlp"]["acts"]
        for i in range(5):
            x, u = acts[i]
            km = masks[d, i]                                        # [P][32]
            diff = km != (u > 0)
            flips += int(diff.sum())
            assert np.abs(u[diff]).max(initial=0.0) < 1e-5, (name, i, float(np.abs(u[diff]).max()))
            mag = np.maximum(np.abs(u), np.float32(1e-30))
            acts[i] = (x, np.where(km, mag, -mag).astype(np.float32))
    assert flips < 2000, flips                                      # of 23 M decisions
    og = orc.render_batch_ray_backward(sc, cache, g_d, g_v, g_c)
    worst = {}
    for name in ("middle", "fine", "color"):
        for key, p in getattr(decoders, name + "_decoder").named_parameters():
            ref = og["decoders"][name][key]
            if np.abs(ref).max() > 0:
                got = p.grad.cpu().numpy()
                worst[f"{name}.{key}"] = (rel_err(got, ref), elem_err(got, ref))
        gk = "grid_" + name
        got, ref = cg[gk].grad.cpu().numpy(), og["grids"][gk]
        worst[gk] = (rel_err(got, ref), elem_err(got, ref))
        assert int((got != 0).sum()) == int((ref != 0).sum()), gk
    worst["rays_o"] = (rel_err(ro.grad.cpu().numpy(), og["rays_o"]), elem_err(ro.grad.cpu().numpy(), og["rays_o"]))
    worst["rays_d"] = (rel_err(rd.grad.cpu().numpy(), og["rays_d"]), elem_err(rd.grad.cpu().numpy(), og["rays_d"]))
    bad = {k: v for k, v in worst.items() if v[0] >= TOL_GRAD or v[1] >= TOL_GRAD}
    print("relu decisions that differ from the oracle's:", flips, " worst max-norm / elementwise error:",
          max(v[0] for v in worst.values()), max(v[1] for v in worst.values()))
    assert not bad, bad


def test_split_tcgen05_backward_is_an_equivalent_ab_path(tiny, monkeypatch):
    """ENS_BWD_TC_SPLIT=1: data-gradient kernel (two tile groups, writes the g_u rows) + weight-gradient kernel, instead of the
    fused kernel.  Same raw sums -> same unfolded gradients (atomic order aside)."""
    scene, renderer, decoders, c = tiny["scene"], tiny["renderer"], tiny["decoders"], tiny["c"]
    g = tiny["g"]
    ro = torch.from_numpy(g["color.d.rays_o"]).to(DEV); rd = torch.from_numpy(g["color.d.rays_d"]).to(DEV)
    sd = torch.from_numpy(g["color.d.sample_depth"]).to(DEV)

    def run():
        for p in decoders.parameters():
            p.grad = None
        cg = {k: v.detach().clone().requires_grad_(True) for k, v in c.items()}
        ro_ = ro.clone().requires_grad_(True); rd_ = rd.clone().requires_grad_(True)
        d, u, col = renderer.render_batch_ray(cg, decoders, rd_, ro_, DEV, "color", gt_depth=sd)
        (d.sum() + u.sum() + col.sum()).backward()
        out = {"ro": ro_.grad.clone(), "rd": rd_.grad.clone()}
        out.update({k: v.grad.clone() for k, v in cg.items() if v.grad is not None})
        out.update({n: p.grad.clone() for n, p in decoders.named_parameters() if p.grad is not None})
        return out
    fused = run()
    monkeypatch.setenv("ENS_BWD_TC_SPLIT", "1")
    split = run()
    assert set(fused) == set(split)
    for k in fused:
        assert rel_err(split[k].cpu().numpy(), fused[k].cpu().numpy()) < 1e-4, k


@pytest.mark.parametrize("var", ["ENS_TC_PIPE", "ENS_TC_MULTI"])
def test_decode_ab_switches_are_bitwise_equivalent(tiny, monkeypatch, var):
    """The pipelined gather (ENS_TC_PIPE) and the one-launch-per-stage forward (ENS_TC_MULTI) only change WHEN work is done:
    outputs of eval_points and of a colour-stage render must be bit-identical with the switch off."""
    scene, renderer, decoders, c = tiny["scene"], tiny["renderer"], tiny["decoders"], tiny["c"]
    g = tiny["g"]
    pts = torch.from_numpy(cases.eval_points_lattice(scene)).to(DEV)
    ro = torch.from_numpy(g["color.d.rays_o"]).to(DEV); rd = torch.from_numpy(g["color.d.rays_d"]).to(DEV)
    sd = torch.from_numpy(g["color.d.sample_depth"]).to(DEV)

    def run():
        with torch.no_grad():
            ev = renderer.eval_points(pts, decoders, c, "color", DEV).clone()
            d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=sd)
        return ev, d.clone(), u.clone(), col.clone()
    on = run()
    monkeypatch.setenv(var, "0")
    off = run()
    for a, b in zip(on, off):
        assert torch.equal(a, b)

