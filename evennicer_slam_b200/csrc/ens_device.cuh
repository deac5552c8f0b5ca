// Device-side building blocks shared by the render kernel variants (fma / mma).
#pragma once
#include "ens_common.cuh"

namespace ens {

// =============================================================================================
// trilinear coordinates (ATen grid_sampler_3d: bilinear, border, align_corners=True)
// =============================================================================================
struct Vox {
  int x0, y0, z0;
  float fx, fy, fz;   // weight of the +1 corner: ix - floor(ix)
  float gx, gy, gz;   // weight of the  0 corner: (floor(ix)+1) - ix
  float sx, sy, sz;   // d(ix)/d(pn) if the un-clipped coordinate is strictly inside, else 0
};

__device__ __forceinline__ void axis_coord(float pn, int size, int &i0, float &f, float &g, float &s) {
  const float lim = (float)(size - 1);
  const float raw = __fmul_rn(__fmul_rn(__fadd_rn(pn, 1.f), 0.5f), lim);   // ((x+1)/2)*(size-1)
  float c = fminf(lim, fmaxf(raw, 0.f));                                   // border clip (NaN -> 0)
  const float fl = floorf(c);
  i0 = (int)fl;
  f = __fsub_rn(c, fl);
  g = __fsub_rn(__fadd_rn(fl, 1.f), c);
  s = (raw > 0.f && raw < lim) ? 0.5f * lim : 0.f;                         // clip_coordinates_set_grad
}

__device__ __forceinline__ Vox make_vox(const float pn[3], const int dims[3]) {
  Vox v;
  axis_coord(pn[0], dims[2], v.x0, v.fx, v.gx, v.sx);
  axis_coord(pn[1], dims[1], v.y0, v.fy, v.gy, v.sy);
  axis_coord(pn[2], dims[0], v.z0, v.fz, v.gz, v.sz);
  return v;
}

// pn = ((p - lo)/(hi - lo))*2 - 1 in float64, then .float()      (common.py:342-357, decoder.py:171)
__device__ __forceinline__ void normalize64(const double p[3], const double lo[3], const double hi[3], float pn[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double q = __ddiv_rn(__dsub_rn(p[k], lo[k]), __dsub_rn(hi[k], lo[k]));
    pn[k] = __double2float_rn(__dsub_rn(__dmul_rn(q, 2.0), 1.0));
  }
}
// float32 points (Mesher path): the float64 0-dim bounds are cast to float32 by type promotion
__device__ __forceinline__ void normalize32(const float p[3], const double lo[3], const double hi[3], float pn[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float l = __double2float_rn(lo[k]);
    const float span = __double2float_rn(__dsub_rn(hi[k], lo[k]));
    pn[k] = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p[k], l), span), 2.f), 1.f);
  }
}

// corner c: dx = c&1, dy = (c>>1)&1, dz = c>>2  (ATen order tnw,tne,tsw,tse,bnw,bne,bsw,bse)
__device__ __forceinline__ void corner(const Vox &v, const int dims[3], int c, int64_t &lin, float &w) {
  const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
  int x = v.x0 + dx, y = v.y0 + dy, z = v.z0 + dz;
  const bool ok = (x < dims[2]) && (y < dims[1]) && (z < dims[0]);
  w = __fmul_rn(__fmul_rn(dx ? v.fx : v.gx, dy ? v.fy : v.gy), dz ? v.fz : v.gz);
  if (!ok) { w = 0.f; x = min(x, dims[2] - 1); y = min(y, dims[1] - 1); z = min(z, dims[0] - 1); }
  lin = ((int64_t)(z * dims[1] + y) * dims[2] + x) * C;
}

// 32-channel trilinear feature of one point -> row[0..31] (shared, float4-aligned)
__device__ __forceinline__ void gather32(const float *__restrict__ grid, const int dims[3], const Vox &v,
                                         float *__restrict__ row) {
  float acc[C];
#pragma unroll
  for (int j = 0; j < C; ++j) acc[j] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int64_t lin; float w;
    corner(v, dims, c, lin, w);
    const float4 *src = reinterpret_cast<const float4 *>(grid + lin);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = __ldg(src + q);
      acc[4 * q + 0] = fmaf(a.x, w, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(a.y, w, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(a.z, w, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(a.w, w, acc[4 * q + 3]);
    }
  }
  float4 *dst = reinterpret_cast<float4 *>(row);
#pragma unroll
  for (int q = 0; q < 8; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}

// backward of gather32: scatter g (32) into the native-layout gradient grid and/or accumulate the
// gradient wrt the normalised coordinates.
__device__ __forceinline__ void gather32_bwd(const float *__restrict__ grid, float *__restrict__ ggrid,
                                             const int dims[3], const Vox &v, const float (&g)[C], bool want_coord,
                                             float gpn[3]) {
  float gix = 0.f, giy = 0.f, giz = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int64_t lin; float w;
    corner(v, dims, c, lin, w);
    if (ggrid != nullptr && w != 0.f) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        red_add_v4(ggrid + lin + 4 * q, w * g[4 * q], w * g[4 * q + 1], w * g[4 * q + 2], w * g[4 * q + 3]);
    }
    if (want_coord) {
      const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
      const bool ok = (v.x0 + dx < dims[2]) && (v.y0 + dy < dims[1]) && (v.z0 + dz < dims[0]);
      if (ok) {
        const float4 *src = reinterpret_cast<const float4 *>(grid + lin);
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 a = __ldg(src + q);
          dot = fmaf(a.x, g[4 * q], dot); dot = fmaf(a.y, g[4 * q + 1], dot);
          dot = fmaf(a.z, g[4 * q + 2], dot); dot = fmaf(a.w, g[4 * q + 3], dot);
        }
        const float wx = dx ? v.fx : v.gx, wy = dy ? v.fy : v.gy, wz = dz ? v.fz : v.gz;
        gix += (dx ? 1.f : -1.f) * wy * wz * dot;
        giy += (dy ? 1.f : -1.f) * wx * wz * dot;
        giz += (dz ? 1.f : -1.f) * wx * wy * dot;
      }
    }
  }
  gpn[0] = gix * v.sx; gpn[1] = giy * v.sy; gpn[2] = giz * v.sz;
}

// =============================================================================================
// sample placement (Renderer.py:83-171), float64 with the reference's op order.  CTA-collective.
// Thread (rl, s) gets the s-th smallest z of its ray.  zc / zs: shared double[NT].
// =============================================================================================
struct RayArgs {
  const float *rays_o, *rays_d, *gt_depth;
  const double *depth_max;       // [2] or null
  const float *t_vals;           // [n_samples]
  const double *t_surf;          // [n_surface]
  int n_samples, n_surface, S, rpc;
  int64_t R;
};

__device__ __forceinline__ double place_sample(const RayArgs &ra, const DevScene &sc, bool valid, int64_t ray, int rl,
                                               int s, const float o[3], const float d[3], double *zc, double *zs) {
  double cand = 0.0;
  if (valid) {
    double far_bb = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double ok = (double)o[k], dk = (double)d[k];
      const double t0 = __ddiv_rn(__dsub_rn(sc.lo[k], ok), dk);
      const double t1 = __ddiv_rn(__dsub_rn(sc.hi[k], ok), dk);
      far_bb = tmin_nan(far_bb, tmax_nan(t0, t1));
    }
    far_bb = __dadd_rn(far_bb, 0.01);
    if (ra.gt_depth != nullptr) {
      const float dep = ra.gt_depth[ray];
      if (s < ra.n_samples) {
        double far = far_bb;
        if (far == far) far = fmin(fmax(far_bb, 0.0), ra.depth_max[0]);       // clamp(far_bb, 0, max(gt*1.2))
        const float t = ra.t_vals[s];
        const float near = __fmul_rn(dep, 0.01f);
        const float nt = __fmul_rn(near, __fsub_rn(1.0f, t));
        cand = __dadd_rn((double)nt, __dmul_rn(far, (double)t));
      } else {
        const double ts = ra.t_surf[s - ra.n_samples];
        const double omt = __dsub_rn(1.0, ts);
        if (dep > 0.f) {
          const double lo = (double)__fmul_rn(0.95f, dep), hi = (double)__fmul_rn(1.05f, dep);
          cand = __dadd_rn(__dmul_rn(lo, omt), __dmul_rn(hi, ts));
        } else {
          cand = __dadd_rn(__dmul_rn(0.001, omt), __dmul_rn(ra.depth_max[1], ts));
        }
      }
    } else {
      const float t = ra.t_vals[s];
      const float nt = __fmul_rn(0.01f, __fsub_rn(1.0f, t));
      cand = __dadd_rn((double)nt, __dmul_rn(far_bb, (double)t));
    }
  }
  if (ra.gt_depth == nullptr || ra.n_surface == 0) return cand;   // already ordered as the reference leaves it
  // torch.sort of the S candidates: rank by counting (ties broken by index; equal values are identical)
  zc[threadIdx.x] = cand;
  __syncthreads();
  if (valid) {
    int rank = 0;
    const double *row = zc + rl * ra.S;
    for (int j = 0; j < ra.S; ++j) {
      const double v = row[j];
      rank += (v < cand) || (v == cand && j < s);
    }
    zs[rl * ra.S + rank] = cand;
  }
  __syncthreads();
  return valid ? zs[threadIdx.x] : 0.0;
}

// =============================================================================================
// render forward
// =============================================================================================
struct FwdArgs {
  DevScene sc;
  RayArgs ra;
  double *depth, *var;
  float *color;
  double *z_out;
  float *w_out, *raw_out;
  // saved-for-backward (optional): relu masks [3 decoders][n_tiles][5][32] and activation tiles
  // [3 decoders][n_tiles][5][1024]; a tile = 32 consecutive sample points
  uint32_t *save_masks;
  float *save_h;
  int64_t n_tiles;
  float *save_r = nullptr;      // tcgen05 forward, saved kind 3: relu outputs [3 decoders][ceil(P/128)][5][128][32]
};

struct BwdArgs {
  DevScene sc;
  RayArgs ra;
  const uint32_t *save_masks;   // from the forward (null: recompute)
  const float *save_h;          // from the forward (null: recompute when decoder grads are wanted)
  int64_t n_tiles;
  int mask_fmt;                 // 0: words in the mma kernels' lane layout; 1: one word per point (tcgen05 forward)
  int dec_par;                  // 1: blockIdx.y selects ONE decoder of the stage (small pose-only batches: 3x the CTAs,
                                //    a third of the per-CTA latency); ray gradients are then accumulated atomically
  const float *raw;
  const double *g_depth, *g_var;
  const float *g_color;
  float *ggrid[4];
  float *gdec[4];
  float *g_rays_o, *g_rays_d;
  float *hscratch;          // activation scratch when decoder grads are requested (ens_bwd_workspace_bytes)
  // split mapping backward: scratch between the data-gradient kernel and the weight-gradient kernel
  float *split_gh;          // [3 decoders][n_tiles][5][1024]  g_h tiles
  uint32_t *split_mw;       // [3 decoders][n_tiles][5][32]    relu mask words per point
  float *split_pts;         // [n_tiles][32][8]                p.float(), valid, normalised coordinates
  const float *save_r = nullptr;   // tcgen05 forward, saved kind 3 (see FwdArgs)
};

constexpr int NT_RENDER = 192;   // fma variant: 4 rays x 48 samples (6 x 32-sample rays)
constexpr int NT_MMA = 384;      // mma variant: 12 warps x 32 points = 8 rays x 48 samples

inline int check_scene(const EnsScene *sc, int stage) {
  if (!sc) return ENS_EINVAL;
  if (stage < 0 || stage > 3) return ENS_EINVAL;
  const int need[4][3] = {{ENS_LEVEL_COARSE, -1, -1}, {ENS_LEVEL_MIDDLE, -1, -1}, {ENS_LEVEL_MIDDLE, ENS_LEVEL_FINE, -1},
                          {ENS_LEVEL_MIDDLE, ENS_LEVEL_FINE, ENS_LEVEL_COLOR}};
  for (int k = 0; k < 3; ++k) {
    const int l = need[stage][k];
    if (l < 0) continue;
    if (!sc->grid[l] || !sc->weights[l]) return ENS_EINVAL;
    for (int q = 0; q < 3; ++q) if (sc->dims[l][q] < 1) return ENS_ESHAPE;
  }
  return ENS_OK;
}

inline int check_cfg(const EnsRenderCfg *cfg, bool has_depth, int stage, int &S, int &ns) {
  if (!cfg || !cfg->t_vals) return ENS_EINVAL;
  if (cfg->n_importance != 0 || cfg->lindisp != 0 || cfg->perturb != 0.f || cfg->occupancy != 1) return ENS_EUNSUPPORTED;
  if (cfg->n_samples < 1 || cfg->n_surface < 0) return ENS_ESHAPE;
  ns = (has_depth && stage != ENS_STAGE_COARSE) ? cfg->n_surface : 0;
  if (ns > 0 && !cfg->t_vals_surface) return ENS_EINVAL;
  S = cfg->n_samples + ns;
  if (S > ENS_MAX_SAMPLES || S > NT_RENDER) return ENS_ESHAPE;
  return ENS_OK;
}

inline void fill_ray_args(RayArgs &ra, const EnsRenderCfg *cfg, int stage, const float *rays_o, const float *rays_d,
                          const float *gt_depth, const double *depth_max, int64_t R, int S, int ns) {
  ra.rays_o = rays_o; ra.rays_d = rays_d;
  ra.gt_depth = (stage == ENS_STAGE_COARSE) ? nullptr : gt_depth;
  ra.depth_max = depth_max;
  ra.t_vals = cfg->t_vals; ra.t_surf = cfg->t_vals_surface;
  ra.n_samples = cfg->n_samples; ra.n_surface = ns; ra.S = S; ra.rpc = 1; ra.R = R;
}

// mma-variant launchers (ens_render_mma.cu)
int mma_eval_points(const DevScene &sc, int stage, const void *pts, int pts_is_f64, int64_t n, int apply_mask,
                    float *out4, cudaStream_t s);
int mma_render_fwd(FwdArgs &a, int stage, cudaStream_t s);
// tcgen05 variant (ens_decode_tc.cu)
int tc_eval_points(const DevScene &sc, int stage, const void *pts, int pts_is_f64, int64_t n, int apply_mask,
                   float *out4, cudaStream_t s, uint32_t *msave = nullptr, int64_t mstride = 0, float *rsave = nullptr,
                   int64_t rstride = 0);
constexpr int64_t TC_RSAVE_TILE_FLOATS = 5 * 128 * 32;     // r_0..r_4 of one 128-point tile of one decoder
// bytes of the saved-for-backward buffer of the tcgen05 forward with decoder gradients (saved kind 3)
inline int64_t tc_saved_r_bytes(int64_t n_rays, int S, int stage) {
  if (stage == ENS_STAGE_COARSE || n_rays <= 0 || S < 1) return 0;
  const int ndec = stage == ENS_STAGE_MIDDLE ? 1 : (stage == ENS_STAGE_FINE ? 2 : 3);
  return (int64_t)ndec * ((n_rays * S + 127) / 128) * TC_RSAVE_TILE_FLOATS * (int64_t)sizeof(float);
}
// saved kind 3 = [relu outputs | relu mask words]: masks as in kind 2, [decoder][n_tiles32][5][32], n_tiles32 = 4 ceil(P / 128)
inline int64_t tc_saved3_bytes(int64_t n_rays, int S, int stage, int64_t *r_bytes, int64_t *n_tiles32) {
  const int64_t rb = tc_saved_r_bytes(n_rays, S, stage);
  if (rb <= 0) return 0;
  const int ndec = stage == ENS_STAGE_MIDDLE ? 1 : (stage == ENS_STAGE_FINE ? 2 : 3);
  const int64_t nt = 4 * ((n_rays * S + 127) / 128);
  if (r_bytes) *r_bytes = rb;
  if (n_tiles32) *n_tiles32 = nt;
  return rb + (int64_t)ndec * nt * 160 * 4;
}
// tcgen05 backward (ens_bwd_tc.cu)
int tc_render_bwd(BwdArgs &a, int stage, bool wg, void *workspace, int64_t workspace_bytes, cudaStream_t s);
int64_t tc_bwd_workspace_bytes(int64_t n_rays, int S, int stage, bool wg);
int tc_render_fwd(FwdArgs &a, int stage, void *scratch, int64_t scratch_bytes, cudaStream_t s);
int64_t tc_fwd_scratch_bytes(int64_t n_rays, int S, int stage);
int64_t mma_fwd_saved_bytes(int64_t n_rays, int S, int stage, int want_h, int64_t *n_tiles, int64_t *h_offset);
int mma_render_bwd(BwdArgs &a, int stage, bool wg, cudaStream_t s);
int64_t mma_bwd_workspace_bytes(int64_t n_rays, int S);

}  // namespace ens
