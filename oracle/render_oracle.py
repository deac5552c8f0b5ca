"""CPU oracle for the EvenNICER-SLAM ray-rendering hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may use it, and only as the checker / CPU baseline.

Parity pinning: the reference ships no tests or golden vectors,
so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container from /root/reference with its two CPU shims
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks them on every run).

Each function cites the reference lines it follows.  dtype rules (float64 for
sample placement / points / depth sums, float32 for features / MLP / weights)
follow SURVEY.md 9.3 and are written out explicitly.

The backward functions are the analytic transposes of the forward (SURVEY.md
9.4); they are pinned against the reference's torch-autograd gradients in the
same golden files.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

F32 = np.float32
F64 = np.float64

STAGE_DECODERS = {          # decoder.py:312-342 -- which decoders run in which stage
    "coarse": ("coarse",),
    "middle": ("middle",),
    "fine": ("fine", "middle"),
    "color": ("fine", "color", "middle"),
}


# ----------------------------------------------------------------------------
# pixel selection and ray generation
# ----------------------------------------------------------------------------

def select_pixels(indices: np.ndarray, H0: int, H1: int, W0: int, W1: int,
                  depth: np.ndarray, color: np.ndarray):
    """common.py:92-142 (get_sample_uv/select_uv) given the drawn flat ``indices``.

    The reference builds a (H1-H0, W1-W0) meshgrid of linspace(W0,W1-1) x linspace(H0,H1-1)
    (exact integers in float32), flattens it row-major and indexes it with
    ``torch.randint`` output.  flat = row*Wc + col, i = W0+col, j = H0+row.
    """
    Wc = W1 - W0
    row = indices // Wc
    col = indices % Wc
    i = (W0 + col).astype(F32)
    j = (H0 + row).astype(F32)
    d = depth[H0:H1, W0:W1].reshape(-1)[indices]
    c = color[H0:H1, W0:W1].reshape(-1, 3)[indices]
    return i, j, d, c


def rays_from_uv(i: np.ndarray, j: np.ndarray, c2w: np.ndarray, fx, fy, cx, cy):
    """common.py:74-89.  float32; products rounded separately, summed k=0,1,2 in order."""
    i = i.astype(F32)
    j = j.astype(F32)
    c2w = c2w.astype(F32)
    dirs = np.stack([(i - F32(cx)) / F32(fx), -(j - F32(cy)) / F32(fy), -np.ones_like(i)], -1)
    prod = dirs[..., None, :] * c2w[:3, :3]                  # (..., 3r, 3k)
    rays_d = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    rays_o = np.broadcast_to(c2w[:3, 3], rays_d.shape).copy()
    return rays_o, rays_d


def rays_from_uv_backward(i, j, fx, fy, cx, cy, g_rays_o, g_rays_d):
    """Transpose of rays_from_uv wrt c2w[:3,:4] ('rays -> pose')."""
    i = i.astype(F32).reshape(-1)
    j = j.astype(F32).reshape(-1)
    dirs = np.stack([(i - F32(cx)) / F32(fx), -(j - F32(cy)) / F32(fy), -np.ones_like(i)], -1)
    g = np.zeros((3, 4), dtype=F64)
    g[:, :3] = g_rays_d.reshape(-1, 3).astype(F64).T @ dirs.astype(F64)
    g[:, 3] = g_rays_o.reshape(-1, 3).astype(F64).sum(0)
    return g.astype(F32)


def pixel_lattice(H: int, W: int, new_H: Optional[int] = None, new_W: Optional[int] = None,
                  lin_w: Optional[np.ndarray] = None, lin_h: Optional[np.ndarray] = None):
    """Pixel coordinates of get_rays / get_rays_rescale (common.py:300-340), row-major (H,W).

    ``lin_w``/``lin_h`` are the ``torch.linspace`` vectors the reference builds; for
    the full-resolution lattice they are exact integers.
    """
    nH = H if new_H is None else new_H
    nW = W if new_W is None else new_W
    if lin_w is None:
        assert nW == W, "rescaled lattices need the torch.linspace vector"
        lin_w = np.arange(W, dtype=F32)
    if lin_h is None:
        assert nH == H
        lin_h = np.arange(H, dtype=F32)
    i = np.broadcast_to(lin_w[None, :], (nH, nW)).astype(F32)
    j = np.broadcast_to(lin_h[:, None], (nH, nW)).astype(F32)
    return i, j


# ----------------------------------------------------------------------------
# trilinear feature gather (ATen grid_sampler_3d, bilinear/border/align_corners=True)
# ----------------------------------------------------------------------------

def normalize_3d_coordinate(p64: np.ndarray, bound: np.ndarray) -> np.ndarray:
    """common.py:342-357, in float64 (p's dtype), same op order."""
    p64 = p64.astype(F64)
    out = np.empty_like(p64)
    for k in range(3):
        out[:, k] = ((p64[:, k] - bound[k, 0]) / (bound[k, 1] - bound[k, 0])) * 2 - 1.0
    return out


def _grid_coords(pn32: np.ndarray, dims_zyx):
    """Un-normalise + border clip + floor; returns everything the gather and its backward need."""
    Z, Y, X = dims_zyx
    res = {}
    for name, col, size in (("x", 0, X), ("y", 1, Y), ("z", 2, Z)):
        raw = ((pn32[:, col] + F32(1.0)) / F32(2.0)) * F32(size - 1)
        clipped = np.minimum(F32(size - 1), np.maximum(raw, F32(0.0))).astype(F32)
        # NaN coordinates: ATen's std::max/min send NaN to 0 -- mirror that for robustness
        clipped = np.where(np.isnan(raw), F32(0.0), clipped).astype(F32)
        i0 = np.floor(clipped).astype(F32)
        res[name] = (raw, clipped, i0, size)
    return res


def trilinear_gather(grid: np.ndarray, pn32: np.ndarray):
    """decoder.py:168-175 / 254-260: F.grid_sample(c, vgrid, 'bilinear', border, align_corners=True).

    grid (1,C,Z,Y,X) float32; pn32 (N,3) float32 normalised coords (x->X, y->Y, z->Z).
    Returns (N,C) float32 and a cache for the backward.  Corner order and weight
    formulas follow ATen's grid_sampler_3d (tnw,tne,tsw,tse,bnw,bne,bsw,bse).
    """
    g = grid[0]
    C, Z, Y, X = g.shape
    cs = _grid_coords(pn32.astype(F32), (Z, Y, X))
    ix, iy, iz = cs["x"][1], cs["y"][1], cs["z"][1]
    x0, y0, z0 = cs["x"][2], cs["y"][2], cs["z"][2]
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    wx = (x1 - ix, ix - x0)
    wy = (y1 - iy, iy - y0)
    wz = (z1 - iz, iz - z0)
    N = pn32.shape[0]
    out = np.zeros((N, C), dtype=F32)
    corners = []
    gflat = g.reshape(C, -1)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi = (x0 + dx).astype(np.int64)
                yi = (y0 + dy).astype(np.int64)
                zi = (z0 + dz).astype(np.int64)
                w = (wx[dx] * wy[dy] * wz[dz]).astype(F32)
                ok = (xi >= 0) & (xi < X) & (yi >= 0) & (yi < Y) & (zi >= 0) & (zi < Z)
                lin = (np.clip(zi, 0, Z - 1) * Y + np.clip(yi, 0, Y - 1)) * X + np.clip(xi, 0, X - 1)
                vals = gflat[:, lin].T                       # (N,C)
                out += np.where(ok[:, None], vals * w[:, None], F32(0.0)).astype(F32)
                corners.append((dx, dy, dz, lin, ok, w, vals))
    cache = dict(cs=cs, corners=corners, shape=(C, Z, Y, X), wx=wx, wy=wy, wz=wz)
    return out, cache


def trilinear_gather_backward(cache, g_feat: np.ndarray, want_grid: bool = True,
                              want_coord: bool = True):
    """Transpose of trilinear_gather.

    g_feat (N,C).  Returns (g_grid (1,C,Z,Y,X) or None, g_pn (N,3) float32 or None).
    Coordinate gradient follows ATen's grid_sampler_3d_backward: d/dix of the eight
    weights, scaled by (size-1)/2, and ZERO where the un-clipped coordinate is
    <= 0 or >= size-1 (clip_coordinates_set_grad).
    """
    C, Z, Y, X = cache["shape"]
    g_grid = None
    if want_grid:
        g_grid = np.zeros((C, Z * Y * X), dtype=F32)
    N = g_feat.shape[0]
    gix = np.zeros(N, dtype=F32)
    giy = np.zeros(N, dtype=F32)
    giz = np.zeros(N, dtype=F32)
    wx, wy, wz = cache["wx"], cache["wy"], cache["wz"]
    for (dx, dy, dz, lin, ok, w, vals) in cache["corners"]:
        if want_grid:
            contrib = np.where(ok[:, None], g_feat * w[:, None], F32(0.0)).astype(F32)   # (N,C)
            for c in range(C):
                np.add.at(g_grid[c], lin, contrib[:, c])
        if want_coord:
            dot = np.where(ok, (vals * g_feat).sum(1), F32(0.0)).astype(F32)
            sx = F32(1.0) if dx else F32(-1.0)
            sy = F32(1.0) if dy else F32(-1.0)
            sz = F32(1.0) if dz else F32(-1.0)
            gix += sx * wy[dy] * wz[dz] * dot
            giy += sy * wx[dx] * wz[dz] * dot
            giz += sz * wx[dx] * wy[dy] * dot
    g_pn = None
    if want_coord:
        g_pn = np.zeros((N, 3), dtype=F32)
        for col, (name, gi) in enumerate((("x", gix), ("y", giy), ("z", giz))):
            raw, _, _, size = cache["cs"][name]
            inside = (raw > 0) & (raw < F32(size - 1))
            g_pn[:, col] = np.where(inside, gi * F32((size - 1) / 2.0), F32(0.0))
    if want_grid:
        g_grid = g_grid.reshape(1, C, Z, Y, X)
    return g_grid, g_pn


# ----------------------------------------------------------------------------
# decoders
# ----------------------------------------------------------------------------

def _relu(x):
    return np.maximum(x, F32(0.0))


def mlp_forward(params: Dict[str, np.ndarray], p64: np.ndarray, c_feat: np.ndarray):
    """MLP.forward after the gather (decoder.py:189-203).  c_feat (N,c_dim) float32."""
    pe = p64.astype(F32)
    B = params["embedder._B"]
    q = (pe @ B).astype(F32)
    emb = np.sin(q).astype(F32)
    x = emb
    acts = []
    for i in range(5):
        W = params[f"pts_linears.{i}.weight"]
        b = params[f"pts_linears.{i}.bias"]
        Wc = params[f"fc_c.{i}.weight"]
        bc = params[f"fc_c.{i}.bias"]
        u = (x @ W.T + b).astype(F32)
        h = (_relu(u) + (c_feat @ Wc.T + bc)).astype(F32)
        acts.append((x, u))
        x = np.concatenate([emb, h], -1) if i == 2 else h
    out = (x @ params["output_linear.weight"].T + params["output_linear.bias"]).astype(F32)
    cache = dict(pe=pe, q=q, emb=emb, acts=acts, h_last=x, c=c_feat)
    return out, cache


def mlp_backward(params, cache, g_out: np.ndarray, want_params: bool = True):
    """Transpose of mlp_forward (SURVEY.md 9.4 'MLP block' / 'Embed').

    Returns (g_params dict, g_c_feat (N,c_dim), g_pe (N,3) float32).
    """
    gp: Dict[str, np.ndarray] = {}
    Wo = params["output_linear.weight"]
    g_h = (g_out @ Wo).astype(F32)
    if want_params:
        gp["output_linear.weight"] = (g_out.T @ cache["h_last"]).astype(F32)
        gp["output_linear.bias"] = g_out.sum(0).astype(F32)
    c = cache["c"]
    g_c = np.zeros_like(c)
    g_emb = np.zeros_like(cache["emb"])
    E = cache["emb"].shape[1]
    for i in (4, 3, 2, 1, 0):
        x, u = cache["acts"][i]
        W = params[f"pts_linears.{i}.weight"]
        Wc = params[f"fc_c.{i}.weight"]
        if want_params:
            gp[f"fc_c.{i}.weight"] = (g_h.T @ c).astype(F32)
            gp[f"fc_c.{i}.bias"] = g_h.sum(0).astype(F32)
        g_c += (g_h @ Wc).astype(F32)
        g_u = (g_h * (u > 0)).astype(F32)
        if want_params:
            gp[f"pts_linears.{i}.weight"] = (g_u.T @ x).astype(F32)
            gp[f"pts_linears.{i}.bias"] = g_u.sum(0).astype(F32)
        g_x = (g_u @ W).astype(F32)
        if i == 3:
            g_emb += g_x[:, :E]
            g_h = g_x[:, E:]
        elif i == 0:
            g_emb += g_x
        else:
            g_h = g_x
    g_q = (g_emb * np.cos(cache["q"])).astype(F32)
    if want_params:
        gp["embedder._B"] = (cache["pe"].T @ g_q).astype(F32)
    g_pe = (g_q @ params["embedder._B"].T).astype(F32)
    return gp, g_c, g_pe


def mlp_no_xyz_forward(params, c_feat: np.ndarray):
    """MLP_no_xyz.forward after the gather (decoder.py:262-274)."""
    x = c_feat
    acts = []
    for i in range(5):
        W = params[f"pts_linears.{i}.weight"]
        b = params[f"pts_linears.{i}.bias"]
        u = (x @ W.T + b).astype(F32)
        h = _relu(u)
        acts.append((x, u))
        x = np.concatenate([c_feat, h], -1) if i == 2 else h
    out = (x @ params["output_linear.weight"].T + params["output_linear.bias"]).astype(F32)
    return out, dict(acts=acts, h_last=x, c=c_feat)


def mlp_no_xyz_backward(params, cache, g_out, want_params=True):
    gp: Dict[str, np.ndarray] = {}
    g_h = (g_out @ params["output_linear.weight"]).astype(F32)
    if want_params:
        gp["output_linear.weight"] = (g_out.T @ cache["h_last"]).astype(F32)
        gp["output_linear.bias"] = g_out.sum(0).astype(F32)
    c = cache["c"]
    Cd = c.shape[1]
    g_c = np.zeros_like(c)
    for i in (4, 3, 2, 1, 0):
        x, u = cache["acts"][i]
        W = params[f"pts_linears.{i}.weight"]
        g_u = (g_h * (u > 0)).astype(F32)
        if want_params:
            gp[f"pts_linears.{i}.weight"] = (g_u.T @ x).astype(F32)
            gp[f"pts_linears.{i}.bias"] = g_u.sum(0).astype(F32)
        g_x = (g_u @ W).astype(F32)
        if i == 3:
            g_c += g_x[:, :Cd]
            g_h = g_x[:, Cd:]
        elif i == 0:
            g_c += g_x
        else:
            g_h = g_x
    return gp, g_c


class OracleScene:
    """Plain container: grids (reference layout), decoder state dicts, bounds."""

    def __init__(self, bound, coarse_bound, grids, decoders):
        self.bound = np.asarray(bound, dtype=F64)
        self.coarse_bound = np.asarray(coarse_bound, dtype=F64)
        self.grids = grids
        self.decoders = decoders

    @classmethod
    def from_synthetic(cls, scene):
        return cls(scene.bound, scene.coarse_bound, scene.grids, scene.decoders)

    def decoder_bound(self, name):
        return self.coarse_bound if name == "coarse" else self.bound


def _decoder_forward(sc: OracleScene, name: str, p64: np.ndarray):
    """One decoder incl. its gathers.  decoder.py:177-203 (MLP), 262-274 (MLP_no_xyz)."""
    bound = sc.decoder_bound(name)
    pn32 = normalize_3d_coordinate(p64, bound).astype(F32)
    feat, gc = trilinear_gather(sc.grids["grid_" + name], pn32)
    cache = dict(name=name, gather=gc, bound=bound)
    if name == "coarse":
        out, mc = mlp_no_xyz_forward(sc.decoders[name], feat)
        cache["mlp"] = mc
        return out[:, 0], cache
    if name == "fine":
        # concat_feature: middle features gathered under no_grad (decoder.py:182-187)
        feat_m, _ = trilinear_gather(sc.grids["grid_middle"], pn32)
        feat = np.concatenate([feat, feat_m], 1)
    out, mc = mlp_forward(sc.decoders[name], p64, feat)
    cache["mlp"] = mc
    return (out if name == "color" else out[:, 0]), cache


def nice_forward(sc: OracleScene, p64: np.ndarray, stage: str):
    """NICE.forward (decoder.py:312-342) -> raw (N,4) float32 [r,g,b,occ]."""
    N = p64.shape[0]
    raw = np.zeros((N, 4), dtype=F32)
    caches = {}
    if stage == "coarse":
        occ, caches["coarse"] = _decoder_forward(sc, "coarse", p64)
        raw[:, 3] = occ
    elif stage == "middle":
        occ, caches["middle"] = _decoder_forward(sc, "middle", p64)
        raw[:, 3] = occ
    elif stage == "fine":
        f, caches["fine"] = _decoder_forward(sc, "fine", p64)
        m, caches["middle"] = _decoder_forward(sc, "middle", p64)
        raw[:, 3] = f + m
    elif stage == "color":
        f, caches["fine"] = _decoder_forward(sc, "fine", p64)
        col, caches["color"] = _decoder_forward(sc, "color", p64)
        m, caches["middle"] = _decoder_forward(sc, "middle", p64)
        raw[:, :] = col
        raw[:, 3] = f + m
    else:
        raise ValueError(stage)
    return raw, caches


def inside_mask(p64: np.ndarray, bound: np.ndarray) -> np.ndarray:
    """Renderer.py:43-47 -- strict inequalities on the (float64) points, un-enlarged bound."""
    m = np.ones(p64.shape[0], dtype=bool)
    for k in range(3):
        m &= (p64[:, k] < bound[k, 1]) & (p64[:, k] > bound[k, 0])
    return m


def eval_points(sc: OracleScene, p: np.ndarray, stage: str):
    """Renderer.eval_points (Renderer.py:24-62); the 500k split has no numerical effect."""
    p64 = p.astype(F64) if p.dtype != F64 else p
    if p.dtype == F32:
        # float32 points (Mesher path): the reference normalises in float32 -- see eval_points_f32
        return eval_points_f32(sc, p, stage)
    raw, caches = nice_forward(sc, p64, stage)
    mask = inside_mask(p64, sc.bound)
    raw[~mask, 3] = F32(100.0)
    return raw, dict(caches=caches, mask=mask)


def eval_points_f32(sc: OracleScene, p32: np.ndarray, stage: str):
    """eval_points for float32 input (Mesher.py:281-319 feeds float32 lattice points).

    With float32 ``p`` the reference's normalisation ``(p - lo)/(hi - lo)*2 - 1``
    promotes against the float64 *0-dim* bound entries, which under torch's type
    promotion keeps float32 (0-dim tensors do not promote a dimensioned tensor of
    the same category).  The bound scalars are therefore rounded to float32 first.
    """
    p32 = p32.astype(F32)
    N = p32.shape[0]

    def norm32(bound):
        out = np.empty_like(p32)
        for k in range(3):
            lo = F32(bound[k, 0])
            span = F32(bound[k, 1] - bound[k, 0])     # 0-dim f64 - 0-dim f64 = f64, then cast on use
            out[:, k] = ((p32[:, k] - lo) / span) * F32(2) - F32(1.0)
        return out

    raw = np.zeros((N, 4), dtype=F32)
    outs = {}
    for name in STAGE_DECODERS[stage]:
        pn32 = norm32(sc.decoder_bound(name))
        feat, _ = trilinear_gather(sc.grids["grid_" + name], pn32)
        if name == "coarse":
            o, _ = mlp_no_xyz_forward(sc.decoders[name], feat)
            outs[name] = o[:, 0]
            continue
        if name == "fine":
            fm, _ = trilinear_gather(sc.grids["grid_middle"], pn32)
            feat = np.concatenate([feat, fm], 1)
        o, _ = mlp_forward(sc.decoders[name], p32.astype(F64), feat)
        outs[name] = o if name == "color" else o[:, 0]
    if stage == "color":
        raw[:, :] = outs["color"]
        raw[:, 3] = outs["fine"] + outs["middle"]
    elif stage == "fine":
        raw[:, 3] = outs["fine"] + outs["middle"]
    else:
        raw[:, 3] = outs[stage]
    m = np.ones(N, dtype=bool)
    for k in range(3):
        # mask compares float32 points with 0-dim float64 bounds -> comparison in float32
        m &= (p32[:, k] < F32(sc.bound[k, 1])) & (p32[:, k] > F32(sc.bound[k, 0]))
    raw[~m, 3] = F32(100.0)
    return raw, dict(mask=m)


# ----------------------------------------------------------------------------
# sample placement (float64, bit-exact contract)
# ----------------------------------------------------------------------------

def depth_batch_max(gt_depth: np.ndarray) -> Tuple[float, float]:
    """The two batch-global scalars of Renderer.py:110,145 as float64 values.

    max(gt_depth*1.2) is a float32 product (tensor * python float); multiplication by a
    positive constant is monotone under rounding, so it equals max(gt_depth)*1.2f.
    """
    d = gt_depth.astype(F32).reshape(-1)
    far_clip = F64((d * F32(1.2)).max())
    surf_far = F64(d.max())
    return far_clip, surf_far


def ray_box_far(rays_o: np.ndarray, rays_d: np.ndarray, bound: np.ndarray) -> np.ndarray:
    """Renderer.py:99-106: far_bb (R,) float64 = min_axis(max_pair((bound - o)/d)) + 0.01."""
    o = rays_o.astype(F32)[:, :, None].astype(F64)
    d = rays_d.astype(F32)[:, :, None].astype(F64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (bound[None].astype(F64) - o) / d
    far = np.min(np.max(t, axis=2), axis=1)
    return far + 0.01


def place_samples(rays_o, rays_d, gt_depth, bound, t_vals: np.ndarray, t_surf: Optional[np.ndarray],
                  n_samples: int, n_surface: int, far_clip=None, surf_far=None) -> np.ndarray:
    """z_vals (R,S) float64 -- Renderer.py:83-171 with lindisp=False, perturb=0.

    t_vals: float32 ``torch.linspace(0,1,N_samples)``; t_surf: float64
    ``torch.linspace(0,1,N_surface).double()`` (passed in, not restated, so the bits
    are those of the torch build that also drives the device under test).
    far_clip / surf_far override the batch maxima (used when a ray batch is sharded).
    """
    R = rays_o.shape[0]
    t32 = t_vals.astype(F32)
    far_bb = ray_box_far(rays_o, rays_d, bound)[:, None]                 # (R,1) f64
    if gt_depth is None:
        near_term = (F32(0.01) * (F32(1.0) - t32)).astype(F32)             # (S,) f32
        z = near_term.astype(F64)[None, :] + far_bb * t32.astype(F64)[None, :]
        return z
    d = gt_depth.astype(F32).reshape(-1, 1)
    if far_clip is None or surf_far is None:
        fc, sf = depth_batch_max(d)
        far_clip = fc if far_clip is None else far_clip
        surf_far = sf if surf_far is None else surf_far
    far = np.minimum(np.maximum(far_bb, 0.0), F64(far_clip))             # clamp(far_bb, 0, max)
    far = np.where(np.isnan(far_bb), far_bb, far)
    near = (d * F32(0.01)).astype(F32)                                     # (R,1) f32
    near_term = (near * (F32(1.0) - t32)[None, :]).astype(F32)            # f32 product
    z = near_term.astype(F64) + far * t32.astype(F64)[None, :]
    if n_surface > 0:
        ts = t_surf.astype(F64)[None, :]
        lo = (F32(0.95) * d).astype(F32).astype(F64)
        hi = (F32(1.05) * d).astype(F32).astype(F64)
        z_nz = lo * (1.0 - ts) + hi * ts
        z_zero = 0.001 * (1.0 - ts) + F64(surf_far) * ts
        z_surf = np.where(d > 0, z_nz, z_zero)
        z = np.sort(np.concatenate([z, z_surf], -1), -1)
    return z


# ----------------------------------------------------------------------------
# compositing
# ----------------------------------------------------------------------------

def composite(raw: np.ndarray, z_vals: np.ndarray):
    """raw2outputs_nerf_color, occupancy=True (common.py:256-297).

    raw (R,S,4) float32 (occ in [...,3]); z_vals (R,S) float64.
    Returns depth f64, var f64, rgb f32, weights f32, and a cache.
    """
    occ = raw[..., 3].astype(F32)
    with np.errstate(over="ignore"):
        alpha = (F32(1.0) / (F32(1.0) + np.exp(-(F32(10.0) * occ)))).astype(F32)
    one_minus = ((F32(1.0) - alpha) + F32(1e-10)).astype(F32)
    R, S = alpha.shape
    T = np.ones((R, S), dtype=F32)
    for s in range(1, S):                                   # sequential cumprod, float32
        T[:, s] = T[:, s - 1] * one_minus[:, s - 1]
    w = (alpha * T).astype(F32)
    rgb = (w[..., None] * raw[..., :3]).astype(F32).sum(-2, dtype=F32)
    depth = (w.astype(F64) * z_vals).sum(-1)
    tmp = z_vals - depth[:, None]
    var = (w.astype(F64) * tmp * tmp).sum(1)
    return depth, var, rgb, w, dict(alpha=alpha, T=T, one_minus=one_minus, w=w, z=z_vals,
                                    depth=depth, rgb_pts=raw[..., :3])


def composite_backward(cache, g_depth, g_var, g_color):
    """SURVEY.md 9.4 'Compositing'.  Returns (g_occ (R,S) f32, g_rgb_pts (R,S,3) f32)."""
    a, T, om, w, z = cache["alpha"], cache["T"], cache["one_minus"], cache["w"], cache["z"]
    depth = cache["depth"]
    R, S = a.shape
    g_depth = np.zeros(R, F64) if g_depth is None else g_depth.astype(F64)
    g_var = np.zeros(R, F64) if g_var is None else g_var.astype(F64)
    g_color = np.zeros((R, 3), F32) if g_color is None else g_color.astype(F32)
    dz = z - depth[:, None]
    # depth appears in var through tmp = z - depth: d var / d depth = -2 sum w (z - depth)
    g_depth_total = g_depth + g_var * (-2.0 * (w.astype(F64) * dz).sum(1))
    g_w = (cache["rgb_pts"].astype(F64) * g_color[:, None, :].astype(F64)).sum(-1) \
        + g_depth_total[:, None] * z + g_var[:, None] * dz * dz
    g_w = g_w.astype(F32)
    g_rgb = (w[..., None] * g_color[:, None, :]).astype(F32)
    # suffix sums of w_k g_w_k for k > s
    wg = (w * g_w).astype(F32)
    suffix = np.zeros_like(wg)
    acc = np.zeros(R, dtype=F32)
    for s in range(S - 1, -1, -1):
        suffix[:, s] = acc
        acc = acc + wg[:, s]
    g_alpha = (T * g_w - suffix / om).astype(F32)
    g_occ = (F32(10.0) * a * (F32(1.0) - a) * g_alpha).astype(F32)
    return g_occ, g_rgb


# ----------------------------------------------------------------------------
# render_batch_ray forward / backward
# ----------------------------------------------------------------------------

def render_batch_ray(sc: OracleScene, rays_o, rays_d, stage: str, gt_depth, t_vals, t_surf,
                     n_samples=32, n_surface=16, far_clip=None, surf_far=None):
    """Renderer.render_batch_ray (Renderer.py:64-199), N_importance=0, perturb=0, occupancy=True.

    Returns depth (R,) f64, uncertainty (R,) f64, color (R,3) f32 and a cache.
    """
    if stage == "coarse":
        gt_depth = None
    ns = n_surface if gt_depth is not None else 0
    rays_o = rays_o.astype(F32)
    rays_d = rays_d.astype(F32)
    z = place_samples(rays_o, rays_d, gt_depth, sc.bound, t_vals, t_surf, n_samples, ns,
                      far_clip, surf_far)
    R, S = z.shape
    pts = rays_o.astype(F64)[:, None, :] + rays_d.astype(F64)[:, None, :] * z[:, :, None]
    p64 = pts.reshape(-1, 3)
    raw, caches = nice_forward(sc, p64, stage)
    mask = inside_mask(p64, sc.bound)
    raw[~mask, 3] = F32(100.0)
    raw = raw.reshape(R, S, 4)
    depth, var, rgb, w, cc = composite(raw, z)
    cache = dict(stage=stage, z=z, p64=p64, caches=caches, mask=mask, comp=cc, R=R, S=S,
                 raw=raw, weights=w)
    return depth, var, rgb, cache


def render_batch_ray_backward(sc: OracleScene, cache, g_depth, g_var, g_color,
                              want_params=True, want_grids=True, want_rays=True):
    """Backward of render_batch_ray into grids / decoder params / rays.

    Returns dict(grids={'grid_x': ...}, decoders={name: {key: grad}}, rays_o, rays_d).
    """
    stage, R, S = cache["stage"], cache["R"], cache["S"]
    g_occ, g_rgb = composite_backward(cache["comp"], g_depth, g_var, g_color)
    g_occ = g_occ.reshape(-1)
    g_occ = np.where(cache["mask"], g_occ, F32(0.0)).astype(F32)      # raw[~mask,3]=100 cuts the graph
    g_rgb = g_rgb.reshape(-1, 3)
    N = R * S
    g_p = np.zeros((N, 3), dtype=F64)
    out = dict(grids={}, decoders={})
    for name in STAGE_DECODERS[stage]:
        dc = cache["caches"][name]
        params = sc.decoders[name]
        if name == "color":
            g_out = np.concatenate([g_rgb, np.zeros((N, 1), F32)], 1)   # output 3 is overwritten
        else:
            g_out = g_occ[:, None]
        if name == "coarse":
            gp, g_c = mlp_no_xyz_backward(params, dc["mlp"], g_out, want_params)
            g_pe = None
        else:
            gp, g_c, g_pe = mlp_backward(params, dc["mlp"], g_out, want_params)
        if name == "fine":
            g_c = g_c[:, :32]                                           # middle half is no_grad
        g_grid, g_pn = trilinear_gather_backward(dc["gather"], np.ascontiguousarray(g_c),
                                                 want_grids, want_rays)
        if want_grids:
            key = "grid_" + name
            out["grids"][key] = out["grids"].get(key, 0) + g_grid
        if want_params:
            out["decoders"][name] = gp
        if want_rays:
            b = dc["bound"]
            for k in range(3):
                g_p[:, k] += (g_pn[:, k].astype(F64) * 2) / (b[k, 1] - b[k, 0])
            if g_pe is not None:
                g_p += g_pe.astype(F64)
    if want_rays:
        g_p = g_p.reshape(R, S, 3)
        out["rays_o"] = g_p.sum(1).astype(F32)
        out["rays_d"] = (g_p * cache["z"][:, :, None]).sum(1).astype(F32)
    return out
