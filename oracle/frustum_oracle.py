"""CPU restatement (numpy) of the frustum feature selection and the keyframe overlap test -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/Mapper.py:115-186 (get_mask_from_c2w) and :188-250 (keyframe_selection_overlap), with every
library call the reference makes spelled out in a FIXED arithmetic order so that the CUDA kernels can be held to it bit for
bit:

  * ``w2c @ homo_vertices`` (float32, numpy matmul): products rounded one by one, summed pairwise (p0+p1)+(p2+p3) -- the
    order numpy's float32 matmul inner loop produces on the build container (checked bitwise by
    tests/golden/make_frustum_golden.py); another BLAS build may round the last bit differently, which can only move a
    voxel that sits within one float32 ulp of a threshold;
  * ``K @ cam_cord`` (float64): (fx*X + 0*Y) + cx*Z;
  * ``cv2.remap(depth, u, v, INTER_LINEAR)`` (OpenCV 4.x, float32 image, float32 maps, BORDER_CONSTANT 0): coordinates are
    rounded to 1/32 pixel (``cvRound(x * 32)``, round-half-even, then saturated to int16 pixels), the four weights come
    from OpenCV's 32 x 32 float table ((1-fy)*(1-fx), (1-fy)*fx, fy*(1-fx), fy*fx with fx, fy multiples of 1/32 in float32)
    and the sum is v0*w0 + v1*w1 + v2*w2 + v3*w3 left to right in float32 (imgwarp.cpp: remapBilinear).
Pinned against the reference's own function (run with cv2) by tests/golden/frustum.npz.
"""
import numpy as np

f32 = np.float32


def world_to_camera_f32(w2c, pts):
    """rows 0..2 of (w2c @ [x y z 1]^T) in float32, pairwise order.  pts [N,3] float32 -> [N,3] float32."""
    w2c = np.asarray(w2c, f32)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    out = np.empty((pts.shape[0], 3), f32)
    for r in range(3):
        a = w2c[r]
        out[:, r] = (a[0] * x + a[1] * y) + (a[2] * z + a[3] * f32(1))
    return out


def project(cam_cord, fx, fy, cx, cy):
    """Mapper.py:146-151: flips x, applies K in float64, divides by z + 1e-5.  Returns (uv float32 [N,2], z float64 [N])."""
    X = -cam_cord[:, 0].astype(np.float64)
    Y = cam_cord[:, 1].astype(np.float64)
    Z = cam_cord[:, 2].astype(np.float64)
    u = (fx * X + 0.0 * Y) + cx * Z
    v = (0.0 * X + fy * Y) + cy * Z
    z = ((0.0 * X + 0.0 * Y) + 1.0 * Z) + 1e-5
    uv = np.stack([u / z, v / z], -1).astype(f32)
    return uv, z


def remap_linear(img, u, v):
    """cv2.remap(img, u, v, INTER_LINEAR) for a float32 single-channel image and float32 coordinate lists."""
    img = np.asarray(img, f32)
    H, W = img.shape
    with np.errstate(invalid="ignore", over="ignore"):
        sx = np.rint(u.astype(f32) * f32(32)).astype(np.float64)       # cvRound: round half to even
        sy = np.rint(v.astype(f32) * f32(32)).astype(np.float64)
    # cvRound of NaN / out-of-range floats gives INT_MIN on x86 (cvtss2si)
    bad = ~np.isfinite(sx) | (np.abs(sx) >= 2 ** 31)
    sx = np.where(bad, -2.0 ** 31, sx).astype(np.int64)
    bad = ~np.isfinite(sy) | (np.abs(sy) >= 2 ** 31)
    sy = np.where(bad, -2.0 ** 31, sy).astype(np.int64)
    fxi, fyi = sx & 31, sy & 31
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    fx1 = fxi.astype(f32) * f32(1 / 32)
    fy1 = fyi.astype(f32) * f32(1 / 32)
    fx0, fy0 = f32(1) - fx1, f32(1) - fy1
    w = [fy0 * fx0, fy0 * fx1, fy1 * fx0, fy1 * fx1]

    def px(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0)).astype(f32)
    v0, v1, v2, v3 = px(iy, ix), px(iy, ix + 1), px(iy + 1, ix), px(iy + 1, ix + 1)
    return ((v0 * w[0] + v1 * w[1]) + v2 * w[2]) + v3 * w[3]


def voxel_centres(bound, val_shape, linspace):
    """Mapper.py:132-136.  bound [3][2]; val_shape = grid.shape[2:] = (Z, Y, X); linspace(a, b, n) -> float32 array
    (the reference calls torch.linspace on the CPU; tests pass exactly that).  Returns [X*Y*Z, 3] float32, X slowest."""
    xs = linspace(bound[0][0], bound[0][1], val_shape[2])
    ys = linspace(bound[1][0], bound[1][1], val_shape[1])
    zs = linspace(bound[2][0], bound[2][1], val_shape[0])
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    return np.stack([X, Y, Z], -1).reshape(-1, 3).astype(f32)


def get_mask_from_c2w(c2w, val_shape, depth, bound, cam, linspace, w2c=None):
    """Mapper.py:115-186 for the middle / fine / colour grids.  cam = (H, W, fx, fy, cx, cy); c2w float32 [4,4].
    Returns bool [X, Y, Z]."""
    H, W, fx, fy, cx, cy = cam
    c2w = np.asarray(c2w, f32)
    if w2c is None:
        w2c = np.linalg.inv(c2w)
    pts = voxel_centres(bound, val_shape, linspace)
    uv, z = project(world_to_camera_f32(w2c, pts), fx, fy, cx, cy)
    depths = remap_linear(np.asarray(depth, f32), uv[:, 0], uv[:, 1])
    mask = (uv[:, 0] < W) & (uv[:, 0] > 0) & (uv[:, 1] < H) & (uv[:, 1] > 0)
    dmax = depths.max()
    depths = np.where(depths == 0, dmax, depths)
    mask = mask & (0 <= -z) & (-z <= (depths + f32(0.5)).astype(np.float64))
    d = pts - c2w[:3, 3][None, :]                                   # torch float32: sum(dist*dist, axis=1)
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    mask = mask | (d2 < f32(0.25))
    return mask.reshape(val_shape[2], val_shape[1], val_shape[0])


def keyframe_overlap(vertices, w2cs, cam, edge=20):
    """Mapper.py:222-241: per keyframe, the number of the current frame's sample points that project inside its image
    (``percent_inside`` = count / len(vertices)).  vertices [N,3] float32, w2cs [K,4,4] float32 -> int64 [K]."""
    H, W, fx, fy, cx, cy = cam
    out = np.zeros(len(w2cs), np.int64)
    for k, w2c in enumerate(w2cs):
        uv, z = project(world_to_camera_f32(w2c, vertices), fx, fy, cx, cy)
        m = (uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge) & (z < 0)
        out[k] = int(m.sum())
    return out


def overlap_sample_points(rays_o, rays_d, gt_depth, t_vals):
    """Mapper.py:210-217 in float32: near = 0.8 d, far = d + 0.5, z = near (1-t) + far t, pts = o + d z."""
    d = np.asarray(gt_depth, f32).reshape(-1, 1)
    t = np.asarray(t_vals, f32)[None, :]
    near, far = d * f32(0.8), d + f32(0.5)
    z = near * (f32(1.) - t) + far * t
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
    return pts.reshape(-1, 3).astype(f32)
