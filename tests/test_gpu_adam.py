"""Parity of the fused frustum-masked Adam step (ens_grid_adam_step, through FrustumGridAdam) with the oracle
(oracle/adam_oracle.py, pinned against torch.optim.Adam in tests/test_adam_cpu.py).  Needs a B200 (``-m gpu``).

Tolerance: 2e-6 of the largest parameter change plus two ulps of the value (float32 Adam; fused multiply-adds and a
reciprocal multiply where torch's CPU kernels divide); unselected voxels bit-identical."""
import numpy as np
import pytest
import torch

import adam_oracle as ao

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(shapes, seed, mask_frac=0.6):
    from evennicer_slam_b200 import scene as scn
    rng = np.random.RandomState(seed)
    grids = {k: (rng.randn(*s) * 0.01).astype(np.float32) for k, s in shapes.items()}
    masks = {k: rng.rand(*s[2:]) < mask_frac for k, s in shapes.items()}
    c = {k: scn.as_native_layout(torch.from_numpy(v).to(DEV)).requires_grad_(True) for k, v in grids.items()}
    return rng, grids, masks, c


def _native_grad(g_np):
    """a gradient tensor shaped like the fused backward's: a [1,32,Z,Y,X] view of a contiguous [Z,Y,X,32] buffer"""
    from evennicer_slam_b200 import scene as scn
    return scn.as_native_layout(torch.from_numpy(g_np).to(DEV))


def _close(got, want, start):
    scale = np.abs(want - start).max()
    return np.all(np.abs(got - want) <= 2e-6 * scale + 2 * np.spacing(np.abs(want)))


@pytest.mark.parametrize("graph_safe", [False, True])
def test_masked_adam_matches_oracle_over_a_stage_schedule(graph_safe):
    from evennicer_slam_b200.optim import FrustumGridAdam
    shapes = {"grid_middle": (1, 32, 5, 6, 7), "grid_fine": (1, 32, 9, 4, 11), "grid_color": (1, 32, 9, 4, 11)}
    rng, grids, masks, c = _setup(shapes, 20)
    opt = FrustumGridAdam(c, {k: torch.from_numpy(m) for k, m in masks.items()}, graph_safe=graph_safe)
    want = {k: v.copy() for k, v in grids.items()}
    st = {k: ao.MaskedAdamState(s) for k, s in shapes.items()}
    # configs/nice_slam.yaml:45-76 -- a stage gives some levels lr 0 (moments still advance)
    sched = [{"grid_middle": 0.1, "grid_fine": 0.0, "grid_color": 0.0}] * 3 + \
            [{"grid_middle": 0.005, "grid_fine": 0.005, "grid_color": 0.0}] * 2 + \
            [{"grid_middle": 0.005, "grid_fine": 0.005, "grid_color": 0.005}] * 4
    for it, lrs in enumerate(sched):
        dense = {k: (rng.randn(*s) * 10 ** rng.uniform(-4, 0)).astype(np.float32) for k, s in shapes.items()}
        for k in shapes:
            c[k].grad = _native_grad(dense[k])
            ao.adam_step_masked(want[k], dense[k], masks[k], st[k], lrs[k], it + 1)
        if graph_safe:
            opt.set_dynamic(it + 1, lrs)
            opt.step({})
        else:
            opt.step(lrs)
    torch.cuda.synchronize()
    for k in shapes:
        got = c[k].detach().cpu().numpy()
        sel = np.broadcast_to(masks[k][None, None], shapes[k])
        assert np.array_equal(got[~sel], grids[k][~sel]), k
        assert _close(got, want[k], grids[k]), (k, np.abs(got - want[k]).max())
        m, v = opt.state[k]
        assert _close(m.permute(3, 0, 1, 2)[None].cpu().numpy(), st[k].m, np.zeros_like(st[k].m)), k
        assert _close(v.permute(3, 0, 1, 2)[None].cpu().numpy(), st[k].v, np.zeros_like(st[k].v)), k


def test_no_mask_clear_grad_and_ragged_sizes():
    """Every voxel selected; voxel counts that are not multiples of the CTA tile; gradient cleared for the next backward."""
    from evennicer_slam_b200.optim import FrustumGridAdam
    shapes = {"grid_fine": (1, 32, 3, 5, 7), "grid_color": (1, 32, 1, 1, 1)}
    rng, grids, masks, c = _setup(shapes, 4)
    opt = FrustumGridAdam(c, None)
    want = {k: v.copy() for k, v in grids.items()}
    st = {k: ao.MaskedAdamState(s) for k, s in shapes.items()}
    for it in range(3):
        dense = {k: rng.randn(*s).astype(np.float32) for k, s in shapes.items()}
        for k in shapes:
            c[k].grad = _native_grad(dense[k])
            ao.adam_step_masked(want[k], dense[k], None, st[k], 0.01, it + 1)
        opt.step({k: 0.01 for k in shapes}, clear_grad=True)
        for k in shapes:
            assert float(c[k].grad.abs().max()) == 0.0
    for k in shapes:
        assert _close(c[k].detach().cpu().numpy(), want[k], grids[k]), k


def test_rejects_reference_layout_grids_and_bad_masks():
    from evennicer_slam_b200.optim import FrustumGridAdam
    g = torch.zeros(1, 32, 3, 4, 5, device=DEV)
    with pytest.raises(ValueError):
        FrustumGridAdam({"grid_fine": g}, None)
    from evennicer_slam_b200 import scene as scn
    with pytest.raises(ValueError):
        FrustumGridAdam({"grid_fine": scn.as_native_layout(g)}, {"grid_fine": torch.ones(5, 4, 3, dtype=torch.bool)})


def test_mapping_iterations_with_fused_adam_follow_the_reference_sequence():
    """Three colour-stage mapping iterations on the tiny scene: render + loss + backward, then (a) the reference sequence
    -- gathered copy, index_put, torch.optim.Adam, write-back (Mapper.py:343-361, 451-458, 625, 633-641) -- and
    (b) FrustumGridAdam on the dense gradient.  Same renderer on both sides, so the grids must agree to Adam rounding."""
    import cases
    from evennicer_slam_b200 import harness, scene as scn
    from evennicer_slam_b200.optim import FrustumGridAdam
    from util import load_golden
    scene = cases.tiny_scene()
    decoders, c0, renderer, cfg = harness.build(scene, DEV, requires_grad=False)
    g = load_golden("tiny_render.npz")
    ro = torch.from_numpy(g["color.d.rays_o"]).to(DEV); rd = torch.from_numpy(g["color.d.rays_d"]).to(DEV)
    sd = torch.from_numpy(g["color.d.sample_depth"]).to(DEV)
    keys = ("grid_middle", "grid_fine", "grid_color")
    rng = np.random.RandomState(1)
    masks = {k: torch.from_numpy(rng.rand(*c0[k].shape[2:]) < 0.7).to(DEV) for k in keys}
    lrs = {"grid_middle": 0.005, "grid_fine": 0.005, "grid_color": 0.005}

    def loss_of(c):
        d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=sd)
        return torch.where(sd > 0, torch.abs(sd - d), 0.0).sum() + 0.2 * col.abs().sum()

    # (a) reference sequence
    ca = {k: v.clone() for k, v in c0.items()}
    masked = {}
    for k in keys:
        m = masks[k][None, None].repeat(1, 32, 1, 1, 1)
        masked[k] = (ca[k].detach()[m].clone().requires_grad_(True), m)
    opt = torch.optim.Adam([{"params": [masked[k][0]], "lr": lrs[k]} for k in keys])
    for it in range(3):
        for k in keys:
            val = ca[k].detach().clone(); val[masked[k][1]] = masked[k][0]; ca[k] = val
        opt.zero_grad(); loss_of(ca).backward(); opt.step()
        for k in keys:
            val = ca[k].detach(); val[masked[k][1]] = masked[k][0].clone().detach(); ca[k] = val
    # (b) fused
    cb = {k: (scn.as_native_layout(v.clone()).requires_grad_(True) if k in keys else v.clone()) for k, v in c0.items()}
    fopt = FrustumGridAdam(cb, {k: masks[k] for k in keys})
    for it in range(3):
        fopt.zero_grad(); loss_of(cb).backward(); fopt.step(lrs)
    torch.cuda.synchronize()
    for k in keys:
        a, b, s = ca[k].cpu().numpy(), cb[k].detach().cpu().numpy(), c0[k].cpu().numpy()
        assert np.abs(a - s).max() > 1e-3
        # the first steps of Adam move every touched feature by ~lr whatever its gradient: compare to the step size
        assert np.abs(a - b).max() < 1e-3 * np.abs(a - s).max(), (k, np.abs(a - b).max())


@pytest.mark.parametrize("graph_safe", [False, True])
def test_fused_adam_matches_torch_adam_on_decoder_shaped_groups(graph_safe):
    """FusedAdam (ens_tensors_adam_step) against torch.optim.Adam on the same CUDA tensors: decoder-shaped tensors in one
    group, camera tensors in another (different learning rates, changed mid-way as the stages do), one parameter that
    never receives a gradient (torch skips it; so must we, without shifting the other moments)."""
    from evennicer_slam_b200.optim import FusedAdam
    torch.manual_seed(5)
    shapes = [(32, 32), (32,), (3, 93), (32, 93), (32, 125), (4, 32), (4,), (1,)]
    pa = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    cams = [torch.randn(7, device=DEV).requires_grad_(True) for _ in range(4)]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    camsb = [p.detach().clone().requires_grad_(True) for p in cams]
    ref = torch.optim.Adam([{"params": pa, "lr": 0.005}, {"params": cams, "lr": 0.001}])
    opt = FusedAdam([{"params": pb, "lr": 0.005}, {"params": camsb, "lr": 0.001}], graph_safe=graph_safe)
    skip = 2                                                       # pa[2] / pb[2] never get a gradient
    for it in range(6):
        if it == 3:
            ref.param_groups[0]["lr"] = 0.0; opt.param_groups[0]["lr"] = 0.0      # a stage that freezes the decoders
        for k, (a, b) in enumerate(zip(pa + cams, pb + camsb)):
            if k == skip:
                continue
            gr = torch.randn_like(a) * 10 ** float(torch.empty(1).uniform_(-4, 0))
            a.grad = gr.clone(); b.grad = gr.clone()
        ref.step()
        if graph_safe:
            opt.set_dynamic(it + 1)
        opt.step()
    torch.cuda.synchronize()
    for k, (a, b) in enumerate(zip(pa + cams, pb + camsb)):
        a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
        assert np.all(np.abs(a - b) <= 2e-6 * 0.03 + 2 * np.spacing(np.abs(a))), (k, np.abs(a - b).max())


def test_render_after_fused_adam_step_uses_the_updated_weights():
    """FusedAdam / FrustumGridAdam write the parameters through raw pointers; they must bump the autograd versions so that
    SceneCache re-packs (a stale packed decoder was used silently after the first optimiser step)."""
    import cases
    from evennicer_slam_b200 import harness
    from evennicer_slam_b200.optim import FusedAdam
    scene = cases.tiny_scene()
    decoders, c, renderer, cfg = harness.build(scene, DEV)
    rng = np.random.RandomState(3)
    ro = torch.from_numpy(np.tile(np.array([[0.3, 0.2, 0.1]], np.float32), (16, 1))).to(DEV)
    rd = torch.from_numpy((rng.randn(16, 3) * 0.3 + np.array([0.1, 0.1, -1.0])).astype(np.float32)).to(DEV)
    gt = torch.full((16,), 1.5, device=DEV)
    params = list(decoders.parameters())
    opt = FusedAdam(params, lr=0.05)
    d0, _, col0 = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=gt)
    (d0.sum() + col0.sum().double()).backward()
    v_before = [p._version for p in params]
    opt.step()
    assert all(p._version > v for p, v in zip(params, v_before) if p.grad is not None)
    with torch.no_grad():
        d1, _, col1 = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=gt)
        renderer._cache.invalidate()                      # a render with freshly packed weights is the truth
        d2, _, col2 = renderer.render_batch_ray(c, decoders, rd, ro, DEV, "color", gt_depth=gt)
    torch.cuda.synchronize()
    assert torch.equal(d1, d2) and torch.equal(col1, col2)
    assert not torch.equal(col0.detach(), col1)           # and the step did change the output


def test_staged_schedule_skips_levels_without_gradient_like_torch_adam():
    """Mapper.py:462-473: the middle stage gives grid_fine / grid_color (and their decoders) no gradient; torch.optim.Adam
    skips such parameters and keeps a step count per parameter."""
    from evennicer_slam_b200 import scene as scn
    from evennicer_slam_b200.optim import FrustumGridAdam, FusedAdam
    rng = np.random.RandomState(5)
    shapes = {"grid_middle": (1, 32, 4, 5, 6), "grid_fine": (1, 32, 6, 5, 7), "grid_color": (1, 32, 6, 5, 7)}
    start = {k: (rng.randn(*s) * 0.01).astype(np.float32) for k, s in shapes.items()}
    c = {k: scn.as_native_layout(torch.from_numpy(v).to(DEV)).requires_grad_(True) for k, v in start.items()}
    ref = {k: torch.from_numpy(v.copy()).to(DEV).requires_grad_(True) for k, v in start.items()}
    small = [torch.from_numpy(rng.randn(n).astype(np.float32)).to(DEV).requires_grad_(True) for n in (7, 33, 128)]
    small_ref = [t.detach().clone().requires_grad_(True) for t in small]
    gopt = FrustumGridAdam(c, None)
    dopt = FusedAdam(small, lr=0.01)
    topt = torch.optim.Adam([{"params": [ref[k]], "lr": 0.0} for k in shapes] + [{"params": small_ref, "lr": 0.01}])
    stages = [("grid_middle",)] * 3 + [("grid_middle", "grid_fine")] * 2 + [("grid_middle", "grid_fine", "grid_color")] * 3
    for it, live in enumerate(stages):
        lrs = {k: 0.01 for k in live}
        for gi, k in enumerate(shapes):
            topt.param_groups[gi]["lr"] = lrs.get(k, 0.0)
            if k in live:
                g = (rng.randn(*shapes[k]) * 0.1).astype(np.float32)
                c[k].grad = scn.as_native_layout(torch.from_numpy(g).to(DEV))
                ref[k].grad = torch.from_numpy(g).to(DEV)
            else:
                c[k].grad = None
                ref[k].grad = None
        for j, (t, tr) in enumerate(zip(small, small_ref)):
            if j <= len(live) - 1:                         # tensor j first gets a gradient in stage j
                g = torch.from_numpy(rng.randn(t.numel()).astype(np.float32)).to(DEV)
                t.grad, tr.grad = g.clone(), g.clone()
            else:
                t.grad, tr.grad = None, None
        gopt.step(lrs)
        dopt.step()
        topt.step()
    torch.cuda.synchronize()
    assert gopt.steps == {"grid_middle": 8, "grid_fine": 5, "grid_color": 3}
    for k in shapes:
        got, want = c[k].detach().cpu().numpy(), ref[k].detach().cpu().numpy()
        assert _close(got, want, start[k]), k
    for t, tr in zip(small, small_ref):
        assert torch.allclose(t.detach(), tr.detach(), rtol=2e-6, atol=1e-7)
