// Fused render kernels, variant "mma": the 32-wide decoder layers run on the tensor pipe.
//
// ncu on the fma variant (profiles/r01_fma_*) shows the kernels contraction-bound (FMA pipe the busiest
// unit at 19-29 %, DRAM < 0.2 %, one 6-warp CTA per SM), so -- as BASELINE.json's north_star prescribes --
// the layer GEMMs move to mma.sync.m16n8k8 TF32 with the 3xTF32 error-compensated split
// (a = a_hi + a_lo, b = b_hi + b_lo;  a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi), which keeps float32-level
// accuracy (outputs 1e-4, gradients 1e-3 are the contract; plain TF32 would not meet it).
//
// Mapping.  A warp owns 32 consecutive sample points = two m16 row tiles.  Lane L is the "owner" of point
// L (sample placement, mask, trilinear coordinates, final raw value); for the GEMMs the same lane, as
// (g = L/4, t = L%4), holds rows {g, g+8, g+16, g+24} of the accumulator fragments.  Activations never
// leave registers between layers: an accumulator fragment (row g, cols 8n+2t, 8n+2t+1) is reused directly
// as the A fragment of k-tile n of the next layer, with the k order inside a tile permuted to
// (2t, 2t+1); the B fragments are loaded with the same permutation, which in the [out][in] weight layout
// is one 8-byte shared load per lane (XOR-swizzled, conflict-free; see MlpPackV2).
// Grid features are gathered by 8 lanes per point (one 128-byte voxel line per 8 lanes, LDG.128) into a
// warp-private swizzled [32][CD] shared tile and read back as A fragments for the fc_c GEMMs.
#include "ens_mma.cuh"

namespace ens {

// NICE.forward for the owner lane's point (decoder.py:312-342).  CTA-collective.
// sw: weight region; sfeat: the CTA's [NT][RS] feature tile.  SAVE / msave / hsave: see mlp_mma; the pointers address
// this warp's tile of decoder 0 (middle); decoder d lives dstride_m words / dstride_h floats further.
template <int STAGE, int SAVE = 0>
__device__ __forceinline__ float4 decode_stage_mma(const DevScene &sc, float *__restrict__ sw, float *__restrict__ sfeat,
                                                   const float pn[3], const float p32[3], uint32_t *msave = nullptr,
                                                   float *hsave = nullptr, int64_t dstride_m = 0, int64_t dstride_h = 0) {
  constexpr int RS = MmaStage<STAGE>::RS;
  float *crow = sfeat + (threadIdx.x >> 5) * 32 * RS;             // this warp's 32 rows
  float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
  {   // middle decoder; in the fine/colour stages its features also form the fine decoder's concat half
    constexpr int C0 = (RS == 64) ? 32 : 0;
    stage_blob(sw, sc.w[ENS_LEVEL_MIDDLE] + MlpPack<32>::total(), MlpPackV2<32>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_MIDDLE]);
    gather_warp<RS>(sc.grid[ENS_LEVEL_MIDDLE], sc.dims[ENS_LEVEL_MIDDLE], v, crow, C0);
    stage_blob_wait();
    __syncthreads();
    float o[1];
    mlp_mma<32, RS, 1, SAVE>(sw, crow, C0, p32[0], p32[1], p32[2], o, msave, hsave);
    raw.w = o[0];
  }
  if (STAGE == ENS_STAGE_FINE || STAGE == ENS_STAGE_COLOR) {
    __syncthreads();
    stage_blob(sw, sc.w[ENS_LEVEL_FINE] + MlpPack<64>::total(), MlpPackV2<64>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_FINE]);
    gather_warp<RS>(sc.grid[ENS_LEVEL_FINE], sc.dims[ENS_LEVEL_FINE], v, crow, 0);
    stage_blob_wait();
    __syncthreads();
    float o[1];
    mlp_mma<64, RS, 1, SAVE>(sw, crow, 0, p32[0], p32[1], p32[2], o, SAVE ? msave + dstride_m : nullptr,
                             SAVE == 2 ? hsave + dstride_h : nullptr);
    raw.w = __fadd_rn(o[0], raw.w);                               // fine_occ + middle_occ
  }
  if (STAGE == ENS_STAGE_COLOR) {
    __syncthreads();
    stage_blob(sw, sc.w[ENS_LEVEL_COLOR] + MlpPack<32>::total(), MlpPackV2<32>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_COLOR]);
    gather_warp<RS>(sc.grid[ENS_LEVEL_COLOR], sc.dims[ENS_LEVEL_COLOR], v, crow, 0);
    stage_blob_wait();
    __syncthreads();
    float o[4];
    mlp_mma<32, RS, 4, SAVE>(sw, crow, 0, p32[0], p32[1], p32[2], o, SAVE ? msave + 2 * dstride_m : nullptr,
                             SAVE == 2 ? hsave + 2 * dstride_h : nullptr);
    raw.x = o[0]; raw.y = o[1]; raw.z = o[2];
  }
  return raw;
}

// ---------------------------------------------------------------------------------------------
// eval_points (mma)
// ---------------------------------------------------------------------------------------------
template <int STAGE, bool F64>
__global__ void __launch_bounds__(NT_MMA) eval_points_mma_kernel(DevScene sc, const void *__restrict__ pts, int64_t n,
                                                                 int apply_mask, float *__restrict__ out4) {
  extern __shared__ __align__(16) float smem[];
  float *sw = smem;
  float *sfeat = smem + MmaStage<STAGE>::WMAX;
  const int64_t t = (int64_t)blockIdx.x * NT_MMA + threadIdx.x;
  const bool valid = t < n;
  float pn[3], p32[3];
  bool inside = true;
  if (F64) {
    double p[3] = {0.0, 0.0, 0.0};
    if (valid) { const double *pp = (const double *)pts + t * 3; p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2]; }
    normalize64(p, sc.lo, sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k) { p32[k] = __double2float_rn(p[k]); inside &= (p[k] < sc.hi[k]) && (p[k] > sc.lo[k]); }
  } else {
    if (valid) { const float *pp = (const float *)pts + t * 3; p32[0] = pp[0]; p32[1] = pp[1]; p32[2] = pp[2]; }
    else { p32[0] = p32[1] = p32[2] = 0.f; }
    normalize32(p32, sc.lo, sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      inside &= (p32[k] < __double2float_rn(sc.hi[k])) && (p32[k] > __double2float_rn(sc.lo[k]));
  }
  float4 raw = decode_stage_mma<STAGE>(sc, sw, sfeat, pn, p32);
  if (apply_mask && !inside) raw.w = 100.f;
  if (valid) reinterpret_cast<float4 *>(out4)[t] = raw;
}

// ---------------------------------------------------------------------------------------------
// render forward (mma): same structure as the fma variant, different decode
// ---------------------------------------------------------------------------------------------
template <int STAGE, int SAVE>
__global__ void __launch_bounds__(NT_MMA) render_fwd_mma_kernel(FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NT = NT_MMA;
  constexpr int RS = MmaStage<STAGE>::RS;
  float *sw = smem;
  float *sfeat = smem + MmaStage<STAGE>::WMAX;
  double *zc = reinterpret_cast<double *>(sfeat + NT * RS);
  double *zs = zc + NT;
  float4 *sraw = reinterpret_cast<float4 *>(zs + NT);
  float *salpha = reinterpret_cast<float *>(sraw + NT);
  float *sT = salpha + NT;

  const RayArgs &ra = a.ra;
  const int S = ra.S;
  const int rl = threadIdx.x / S, s = threadIdx.x % S;
  const int64_t ray = (int64_t)blockIdx.x * ra.rpc + rl;
  const bool valid = (rl < ra.rpc) && (ray < ra.R);
  float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
  if (valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ra.rays_o[ray * 3 + k]; d[k] = ra.rays_d[ray * 3 + k]; }
  }
  const double z = place_sample(ra, a.sc, valid, ray, rl, s, o, d, zc, zs);
  double p[3];
  float pn[3], p32[3];
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[k] = __dadd_rn((double)o[k], __dmul_rn((double)d[k], z));
    p32[k] = __double2float_rn(p[k]);
    inside &= (p[k] < a.sc.hi[k]) && (p[k] > a.sc.lo[k]);
  }
  normalize64(p, a.sc.lo, a.sc.hi, pn);
  // saved-for-backward tile of this warp: 32 consecutive points starting at blockIdx.x * NT + warp * 32
  // (the launcher only enables SAVE when NT is a multiple of S, so CTA b owns points [b*NT, (b+1)*NT))
  const int64_t tile = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  uint32_t *msave = SAVE ? a.save_masks + tile * 160 + (threadIdx.x & 31) : nullptr;
  float *hsave = SAVE == 2 ? a.save_h + tile * 5120 : nullptr;
  float4 raw = decode_stage_mma<STAGE, SAVE>(a.sc, sw, sfeat, pn, p32, msave, hsave, a.n_tiles * 160, a.n_tiles * 5120);
  if (!inside) raw.w = 100.f;
  // ---- compositing (common.py:285-296) ----
  const float alpha = 1.f / (1.f + expf(-(10.f * raw.w)));
  __syncthreads();
  zs[threadIdx.x] = z;
  sraw[threadIdx.x] = raw;
  salpha[threadIdx.x] = alpha;
  __syncthreads();
  if (valid && s == 0) {
    float T = 1.f;
    for (int k = 0; k < S; ++k) {
      sT[threadIdx.x + k] = T;
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, salpha[threadIdx.x + k]), 1e-10f));
    }
  }
  __syncthreads();
  const float w = __fmul_rn(alpha, sT[threadIdx.x]);
  __syncthreads();
  salpha[threadIdx.x] = w;
  __syncthreads();
  if (valid && s == 0) {
    double dep = 0.0;
    float cr = 0.f, cg = 0.f, cb = 0.f;
    for (int k = 0; k < S; ++k) {
      const float wk = salpha[threadIdx.x + k];
      const float4 rk = sraw[threadIdx.x + k];
      dep = __dadd_rn(dep, __dmul_rn((double)wk, zs[threadIdx.x + k]));
      cr = __fadd_rn(cr, __fmul_rn(wk, rk.x)); cg = __fadd_rn(cg, __fmul_rn(wk, rk.y)); cb = __fadd_rn(cb, __fmul_rn(wk, rk.z));
    }
    double var = 0.0;
    for (int k = 0; k < S; ++k) {
      const double tmp = __dsub_rn(zs[threadIdx.x + k], dep);
      var = __dadd_rn(var, __dmul_rn(__dmul_rn((double)salpha[threadIdx.x + k], tmp), tmp));
    }
    a.depth[ray] = dep;
    a.var[ray] = var;
    a.color[ray * 3 + 0] = cr; a.color[ray * 3 + 1] = cg; a.color[ray * 3 + 2] = cb;
  }
  if (valid) {
    const int64_t pi = ray * S + s;
    if (a.z_out) a.z_out[pi] = z;
    if (a.w_out) a.w_out[pi] = w;
    if (a.raw_out) reinterpret_cast<float4 *>(a.raw_out)[pi] = raw;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int STAGE>
static int launch_eval_mma(const DevScene &sc, const void *pts, int f64, int64_t n, int am, float *out4, cudaStream_t s) {
  const size_t smem = (size_t)(MmaStage<STAGE>::WMAX + NT_MMA * MmaStage<STAGE>::RS) * 4;
  const unsigned g = (unsigned)((n + NT_MMA - 1) / NT_MMA);
  if (f64) {
    ENS_CUDA_CALL(cudaFuncSetAttribute(eval_points_mma_kernel<STAGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_points_mma_kernel<STAGE, true><<<g, NT_MMA, smem, s>>>(sc, pts, n, am, out4);
  } else {
    ENS_CUDA_CALL(cudaFuncSetAttribute(eval_points_mma_kernel<STAGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_points_mma_kernel<STAGE, false><<<g, NT_MMA, smem, s>>>(sc, pts, n, am, out4);
  }
  ENS_CHECK_CUDA();
  return ENS_OK;
}

int mma_eval_points(const DevScene &sc, int stage, const void *pts, int pts_is_f64, int64_t n, int apply_mask,
                    float *out4, cudaStream_t s) {
  switch (stage) {
    case ENS_STAGE_MIDDLE: return launch_eval_mma<ENS_STAGE_MIDDLE>(sc, pts, pts_is_f64, n, apply_mask, out4, s);
    case ENS_STAGE_FINE: return launch_eval_mma<ENS_STAGE_FINE>(sc, pts, pts_is_f64, n, apply_mask, out4, s);
    case ENS_STAGE_COLOR: return launch_eval_mma<ENS_STAGE_COLOR>(sc, pts, pts_is_f64, n, apply_mask, out4, s);
    default: return ENS_EUNSUPPORTED;
  }
}

template <int STAGE, int SAVE>
static int launch_fwd_mma(FwdArgs &a, cudaStream_t s) {
  constexpr int NT = NT_MMA;
  a.ra.rpc = NT / a.ra.S;
  const size_t smem = (size_t)(MmaStage<STAGE>::WMAX + NT * MmaStage<STAGE>::RS) * 4 + (size_t)NT * (8 + 8 + 16 + 4 + 4);
  ENS_CUDA_CALL(cudaFuncSetAttribute(render_fwd_mma_kernel<STAGE, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned g = (unsigned)((a.ra.R + a.ra.rpc - 1) / a.ra.rpc);
  render_fwd_mma_kernel<STAGE, SAVE><<<g, NT, smem, s>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

template <int STAGE>
static int launch_fwd_mma_save(FwdArgs &a, cudaStream_t s) {
  const int save = (a.save_masks != nullptr && (192 % a.ra.S) == 0) ? (a.save_h != nullptr ? 2 : 1) : 0;
  switch (save) {
    case 2: return launch_fwd_mma<STAGE, 2>(a, s);
    case 1: return launch_fwd_mma<STAGE, 1>(a, s);
    default: return launch_fwd_mma<STAGE, 0>(a, s);
  }
}

int mma_render_fwd(FwdArgs &a, int stage, cudaStream_t s) {
  switch (stage) {
    case ENS_STAGE_MIDDLE: return launch_fwd_mma_save<ENS_STAGE_MIDDLE>(a, s);
    case ENS_STAGE_FINE: return launch_fwd_mma_save<ENS_STAGE_FINE>(a, s);
    case ENS_STAGE_COLOR: return launch_fwd_mma_save<ENS_STAGE_COLOR>(a, s);
    default: return ENS_EUNSUPPORTED;
  }
}

// Bytes of the saved-for-backward buffer of the mma forward, 0 if the fast path does not apply (coarse stage, or a
// samples-per-ray count that does not tile the CTAs).  Layout: masks u32 [3][n_tiles][5][32], then (want_h)
// activations f32 [3][n_tiles][5][1024] at *h_offset bytes.
int64_t mma_fwd_saved_bytes(int64_t n_rays, int S, int stage, int want_h, int64_t *n_tiles, int64_t *h_offset) {
  if (stage == ENS_STAGE_COARSE || S < 1 || (192 % S) != 0 || n_rays <= 0) return 0;
  const int rpc = NT_MMA / S;
  const int64_t tiles = ((n_rays + rpc - 1) / rpc) * (NT_MMA / 32);
  const int64_t mbytes = 3 * tiles * 160 * 4;
  if (n_tiles) *n_tiles = tiles;
  if (h_offset) *h_offset = mbytes;
  return mbytes + (want_h ? 3 * tiles * 5120 * 4 : 0);
}

}  // namespace ens
