// NICE.forward on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in TMEM) -- variant "tc".
//
// One launch decodes ONE decoder (middle, fine or colour) for N points; the stage dispatch of NICE.forward
// (decoder.py:312-342) is a sequence of launches that read-modify-write the (N,4) output:
//     middle: out = (0,0,0,occ_m)      fine: out.w = occ_f + out.w      colour: out.xyz = rgb
// and the last launch applies Renderer.eval_points' bound rule (out.w = 100 outside slam.bound, Renderer.py:43-58).
//
// Mapping.  A persistent CTA (one per SM) stages the decoder's weights once -- every [32][K] matrix twice, as the
// value and as its TF32 remainder, in the canonical K-major UMMA layout (MlpPackTC) -- and then walks 128-point tiles.
// Thread r of the CTA owns point r of the tile = TMEM lane r (thread-per-point epilogue):
//   * trilinear gather of its 32 (64) features, split hi/lo, tcgen05.st into the feature columns (A operand in TMEM);
//   * Fourier embedding in three 32-column chunks: sin -> hi/lo -> tcgen05.st; one elected thread issues
//     D  += e W0^T  and  D3 += e W3e^T  (3xTF32: lo.hi, hi.lo, hi.hi; M = 128 points, N = 32, K = 8 per instruction);
//   * per block i: tcgen05.ld of the accumulator; r = relu(D + b'_i) in registers; hi/lo -> x columns; the next
//     block's hidden GEMM and its folded feature GEMM (M_i = W_{i+1} Wc_i, see MlpPackTC) accumulate into D (or D3 for
//     the skip block);
//   * output layer on the FMA pipe (Wo r_4 + Mo c + bo'), read-modify-write of the output row.
// Two tiles are in flight per CTA (two groups of four warps with their own TMEM columns and barriers).
// Completion of each MMA batch is tracked with tcgen05.commit -> mbarrier; thread sync around TMEM stores uses
// tcgen05.wait::st + tcgen05.fence + bar.sync.  Verified first in isolation by tools/tc_probe.cu.
#include "ens_tc.cuh"

namespace ens {

struct TcArgs {
  DevScene sc;
  const void *pts;        // [n][3] f64 or f32
  int64_t n;
  float *out4;            // [n][4]
  int apply_mask;         // last launch of the stage and the caller wants the bound rule
  int separate;           // 1: write this decoder's own plane (no read-modify-write of a shared raw array): middle
                          // (inside ? 1 : 0, 0, 0, occ), fine (0, 0, 0, occ), colour (r, g, b, 0) -- combined by composite_kernel
  uint32_t *msave;        // optional: this decoder's saved relu masks, one word per point and block --
                          // [n_tiles32][5][32], word (tile, i, j) = point 32 tile + j, bit k = unit k of block i active
  float *rsave;           // optional: this decoder's relu outputs r_0..r_4 = relu(u_i), [n_tiles128][5][128][32] -- what the
                          // tcgen05 backward (ens_bwd_tc.cu) needs for the weight gradients (saved kind 3)
  int nctas;              // CTAs that walk this decoder's tiles (0: gridDim.x)
  int nopipe;             // ENS_TC_PIPE=0: gather every tile up front instead of during the previous tile's round trips (A/B)
};

// Two 128-point tiles are in flight per CTA: tile group 0 = warps 0-3, group 1 = warps 4-7; each owns 256 TMEM columns,
// an mbarrier and a named barrier, and walks its own tile sequence, so one group's epilogue overlaps the other's MMAs.
// TMEM columns of a group: accumulators D and D3 (the skip block's, which starts with W3e e), the A operands
// e-chunk / x (hi, lo) and the features (hi, lo; CD columns each).
constexpr int TC_D = 0, TC_D3 = 32, TC_XH = 64, TC_XL = 96, TC_FH = 128;
constexpr int TC_GROUP_COLS = 256, TC_COLS = 512;

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory"); }

template <int LEVEL, int CD, int NO, bool F64>
__device__ __forceinline__ void decode_tc_body(const TcArgs &a, float *smem) {
  using P = MlpPackTC<CD>;
  constexpr int TC_FL = TC_FH + CD;
  static_assert(TC_FL + CD <= TC_GROUP_COLS, "TMEM budget of a tile group");
  static_assert(P::off_W3e() == P::off_W0() + 32 * EMBP && TC_D3 == TC_D + 32, "[W0; W3e] and [D | D3] must be contiguous");
  float *sw = smem;                                     // the tc blob
  float *stile = smem + P::total() + (threadIdx.x >> 5) * 1024;   // this warp's [32][32] gather staging tile
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, grp = tid >> 7, gt = tid & 127, gwarp = (tid >> 5) & 3;

  {   // stage the blob once per CTA
    const float *gw = a.sc.w[LEVEL] + off_tc<CD>();
    const uint32_t s0 = smem_u32(sw);
    for (int i = tid; i < P::total() / 4; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + i * 16), "l"(gw + i * 4) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged weights -> visible to the UMMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb0 = tmem_base_s + (uint32_t)(TC_GROUP_COLS * grp);   // MMA operands address lane 0
  const uint32_t tb = tb0 + ((uint32_t)(32 * gwarp) << 16);             // this warp's lane quarter
  const uint32_t swb = smem_u32(sw);
  uint64_t *bar = &bars[grp];
  uint32_t parity = 0;

  const int64_t n_tiles = (a.n + 127) / 128;
  bool have_pref = false;                             // the staging tile already holds this tile's features (level LEVEL)
  const int64_t tile_stride = (int64_t)(a.nctas > 0 ? a.nctas : (int)gridDim.x) * 2;
  for (int64_t tile = (int64_t)blockIdx.x * 2 + grp; tile < n_tiles; tile += tile_stride) {
    const int64_t pt = tile * 128 + gt;
    const bool valid = pt < a.n;
    float pn[3], p32[3];
    bool inside = true;
    if (F64) {
      double p[3] = {0.0, 0.0, 0.0};
      if (valid) { const double *pp = (const double *)a.pts + pt * 3; p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2]; }
      normalize64(p, a.sc.lo, a.sc.hi, pn);
#pragma unroll
      for (int k = 0; k < 3; ++k) { p32[k] = __double2float_rn(p[k]); inside &= (p[k] < a.sc.hi[k]) && (p[k] > a.sc.lo[k]); }
    } else {
      if (valid) { const float *pp = (const float *)a.pts + pt * 3; p32[0] = pp[0]; p32[1] = pp[1]; p32[2] = pp[2]; }
      else { p32[0] = p32[1] = p32[2] = 0.f; }
      normalize32(p32, a.sc.lo, a.sc.hi, pn);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        inside &= (p32[k] < __double2float_rn(a.sc.hi[k])) && (p32[k] > __double2float_rn(a.sc.lo[k]));
    }

    // ---- features -> TMEM (A operand of the folded feature GEMMs; read back for the output layer) ----
    // Warp-cooperative gather (8 lanes per 128-byte voxel line, as in the mma kernels: a thread-per-point gather costs
    // 8x the L1 wavefronts and was the kernel's top stall) into the warp's swizzled [32][32] staging tile; each thread
    // then reads back its own row.
    // The first half (the decoder's own level) of every tile but a CTA's first was gathered during the previous tile's
    // MMA round trips, one group of four points per round trip (gather_warp_step): its L2 latency hides behind the waits.
#pragma unroll
    for (int half = 0; half < CD / 32; ++half) {      // fine decoder: [fine | middle] concat (decoder.py:182-187)
      const int lv = (half == 0) ? LEVEL : ENS_LEVEL_MIDDLE;
      __syncwarp();
      if (half > 0 || !have_pref) {
        const Vox v = make_vox(pn, a.sc.dims[lv]);
        gather_warp<32>(a.sc.grid[lv], a.sc.dims[lv], v, stile, 0);
      }
      float f[32];
      const int lane = tid & 31;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = *reinterpret_cast<const float4 *>(stile + lane * 32 + ((4 * q) ^ ((lane & 3) << 3)));
        f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
      }
      tmem_st32_split(tb + TC_FH + 32 * half, tb + TC_FL + 32 * half, f);
    }
    // ---- the next tile of this group: its voxel, for the gather steps below ----
    const int64_t tile_n = tile + tile_stride;
    const bool has_next = tile_n < n_tiles && !a.nopipe;
    Vox vn = make_vox(pn, a.sc.dims[LEVEL]);
    if (has_next) {
      const int64_t ptn = tile_n * 128 + gt;
      float pnn[3];
      if (F64) {
        double p[3] = {0.0, 0.0, 0.0};
        if (ptn < a.n) { const double *pp = (const double *)a.pts + ptn * 3; p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2]; }
        normalize64(p, a.sc.lo, a.sc.hi, pnn);
      } else {
        float q[3] = {0.f, 0.f, 0.f};
        if (ptn < a.n) { const float *pp = (const float *)a.pts + ptn * 3; q[0] = pp[0]; q[1] = pp[1]; q[2] = pp[2]; }
        normalize32(q, a.sc.lo, a.sc.hi, pnn);
      }
      vn = make_vox(pnn, a.sc.dims[LEVEL]);
    }
    int gstep = 0;
    __syncwarp();                                      // every lane has read its row of the staging tile
    // ---- embedding chunks: D += e W0^T, D3 += e W3e^T ----
#pragma unroll 1
    for (int jc = 0; jc < 3; ++jc) {
      float e[32];
      const float *B = sw + P::off_B() + 32 * jc;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float q = fmaf(p32[2], B[2 * EMBP + k], fmaf(p32[1], B[EMBP + k], p32[0] * B[k]));
        e[k] = fast_sin(q);
      }
      if (jc > 0) { mbar_wait(bar, parity); parity ^= 1; tc_fence_after(); }     // previous chunk consumed
      tmem_st32_split(tb + TC_XH, tb + TC_XL, e);
      tmem_st_done();
      tc_fence_before();
      group_sync(grp);
      if (gwarp == 0) {
        tc_fence_after();
        // chunk jc = columns 32jc..32jc+31 of the matrices = +8 core matrices = +1024 B.  W3e follows W0 in the blob at
        // exactly four 8-row groups, so [W0; W3e] is ONE canonical [64][96] matrix and [D | D3] (adjacent TMEM
        // columns) one N = 64 accumulator: a single GEMM feeds both.
        issue_gemm<32, EMBP, 64>(tb0 + TC_D, tb0 + TC_XH, tb0 + TC_XL, swb + jc * 1024, P::off_W0(), P::TOT(), jc > 0);
        umma_commit(bar);
        __syncwarp();
      }
      if (has_next) { gather_warp_step<32>(a.sc.grid[LEVEL], a.sc.dims[LEVEL], vn, stile, 0, gstep); ++gstep; }
    }
    // ---- blocks 0..4:  r_i = relu(u_i + b'_i);  u_{i+1} = W_{i+1} r_i + M_i c ----
    float r[32];
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
      mbar_wait(bar, parity); parity ^= 1;
      tc_fence_after();
      tmem_ld32(tb + (i == 3 ? TC_D3 : TC_D), r);
      const float *bi = sw + P::off_b(0) + 32 * i;
#pragma unroll
      for (int k = 0; k < 32; ++k) r[k] = fmaxf(r[k] + bi[k], 0.f);
      if (a.msave != nullptr && valid) {          // saved for a pose-only backward (render_bwd_mma_kernel, mask_fmt 1)
        uint32_t mw = 0u;
#pragma unroll
        for (int k = 0; k < 32; ++k) mw |= (r[k] > 0.f ? 1u : 0u) << k;
        a.msave[(pt >> 5) * 160 + i * 32 + (pt & 31)] = mw;
      }
      if (a.rsave != nullptr) {                   // saved for the backward with decoder gradients (every lane: padded tile)
        float4 *dst = reinterpret_cast<float4 *>(a.rsave + ((tile * 5 + i) * 128 + gt) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q)              // streaming store: read once by the backward, keep the grids in L2
          __stcs(dst + q, valid ? make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]) : make_float4(0.f, 0.f, 0.f, 0.f));
      }
      if (i < 4) {
        tmem_st32_split(tb + TC_XH, tb + TC_XL, r);
        tmem_st_done();
        tc_fence_before();
        group_sync(grp);
        if (gwarp == 0) {
          tc_fence_after();
          const uint32_t dst = tb0 + ((i + 1 == 3) ? TC_D3 : TC_D);      // the skip block continues D3 (= W3e e)
          issue_gemm<32, 32>(dst, tb0 + TC_XH, tb0 + TC_XL, swb, P::off_Wh(1) + i * 1024, P::TOT(), (i + 1 == 3) ? 1u : 0u);
          issue_gemm<CD, CD>(dst, tb0 + TC_FH, tb0 + TC_FL, swb, P::off_M(0) + i * 32 * CD, P::TOT(), 1u);
          umma_commit(bar);
          __syncwarp();
        }
        if (has_next) { gather_warp_step<32>(a.sc.grid[LEVEL], a.sc.dims[LEVEL], vn, stile, 0, gstep); ++gstep; }
      }
    }
    if (has_next) {                                    // seven round trips, eight groups: the last one here
      gather_warp_step<32>(a.sc.grid[LEVEL], a.sc.dims[LEVEL], vn, stile, 0, gstep);
      __syncwarp();
    }
    have_pref = has_next;
    // ---- output layer (FMA pipe): out = Wo r_4 + Mo c + bo' ----
    float o[NO];
#pragma unroll
    for (int q = 0; q < NO; ++q) {
      const float *Wo = sw + P::off_Wo() + 32 * q;
      float s = sw[P::off_bo() + q];
#pragma unroll
      for (int k = 0; k < 32; ++k) s = fmaf(Wo[k], r[k], s);
      o[q] = s;
    }
#pragma unroll
    for (int half = 0; half < CD / 32; ++half) {
      float c[32];
      tmem_ld32(tb + TC_FH + 32 * half, c);                 // the hi columns hold the features themselves
#pragma unroll
      for (int q = 0; q < NO; ++q) {
        const float *Mo = sw + P::off_Mo() + CD * q + 32 * half;
        float s = o[q];
#pragma unroll
        for (int k = 0; k < 32; ++k) s = fmaf(Mo[k], c[k], s);
        o[q] = s;
      }
    }
    if (valid) {
      float4 *dst = reinterpret_cast<float4 *>(a.out4) + pt;
      float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.separate) {
        if (LEVEL == ENS_LEVEL_COLOR) { v4.x = o[0]; v4.y = o[1]; v4.z = o[2]; }
        else v4.w = o[0];
        if (LEVEL == ENS_LEVEL_MIDDLE) v4.x = inside ? 1.f : 0.f;
      } else {
        if (LEVEL == ENS_LEVEL_MIDDLE) v4.w = o[0];
        else if (LEVEL == ENS_LEVEL_FINE) { v4 = *dst; v4.w = __fadd_rn(o[0], v4.w); }
        else { v4 = *dst; v4.x = o[0]; v4.y = o[1]; v4.z = o[2]; }
        if (a.apply_mask && !inside) v4.w = 100.f;
      }
      *dst = v4;
    }
    // the next tile's tcgen05.st must not overtake this tile's TMEM loads
    tc_fence_before();
    group_sync(grp);
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"((uint32_t)TC_COLS) : "memory");
}

template <int LEVEL, int CD, int NO, bool F64>
__global__ void __launch_bounds__(256, 1) decode_tc_kernel(TcArgs a) {
  extern __shared__ __align__(128) float smem[];
  decode_tc_body<LEVEL, CD, NO, F64>(a, smem);
}

// All decoders of a stage in ONE launch, blockIdx.y = decoder, each writing its own output plane (TcArgs::separate).
// Small batches (tracking: 9 600 points = 38 CTAs per decoder): three back-to-back launches are three times the latency of
// one tile chain; side by side they still fit on the 148 SMs.  Larger batches (mapping: 375 tiles per decoder): the SMs are
// dealt to the decoders in proportion to their cost and every CTA walks its decoder's tiles -- 1125 tiles over 147 x 2 tile
// groups is 4 rounds, where three launches of 375 tiles over 296 groups are 3 x 2.
struct TcMultiArgs {
  TcArgs base;
  float *out[3];
  uint32_t *msave[3];
  float *rsave[3];
  int ctas[3];
};

template <int STAGE>
__global__ void __launch_bounds__(256, 1) decode_tc_multi_kernel(TcMultiArgs m) {
  extern __shared__ __align__(128) float smem[];
  TcArgs a = m.base;
  const int d = blockIdx.y;
  if ((int)blockIdx.x >= m.ctas[d]) return;
  a.nctas = m.ctas[d];
  a.out4 = m.out[d];
  a.msave = m.msave[d];
  a.rsave = m.rsave[d];
  a.separate = 1;
  if (d == 0) decode_tc_body<ENS_LEVEL_MIDDLE, 32, 1, true>(a, smem);
  else if (d == 1) decode_tc_body<ENS_LEVEL_FINE, 64, 1, true>(a, smem);
  else if (STAGE == ENS_STAGE_COLOR) decode_tc_body<ENS_LEVEL_COLOR, 32, 4, true>(a, smem);
}

static int tc_nopipe() {
  const char *v = std::getenv("ENS_TC_PIPE");
  return (v && v[0] == '0') ? 1 : 0;
}

template <int LEVEL, int CD, int NO>
static int launch_decode_tc(const DevScene &sc, const void *pts, int f64, int64_t n, int apply_mask, float *out4, cudaStream_t s,
                            uint32_t *msave = nullptr, float *rsave = nullptr) {
  TcArgs a;
  a.sc = sc; a.pts = pts; a.n = n; a.out4 = out4; a.apply_mask = apply_mask; a.msave = msave; a.rsave = rsave; a.separate = 0; a.nctas = 0; a.nopipe = tc_nopipe();
  const size_t smem = (size_t)(MlpPackTC<CD>::total() + 8 * 1024) * 4;      // blob + one staging tile per warp
  const int sms = sm_count();
  const int64_t pairs = ((n + 127) / 128 + 1) / 2;
  const unsigned g = (unsigned)(pairs < sms ? pairs : sms);
  if (f64) {
    ENS_CUDA_CALL(cudaFuncSetAttribute(decode_tc_kernel<LEVEL, CD, NO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decode_tc_kernel<LEVEL, CD, NO, true><<<g, 256, smem, s>>>(a);
  } else {
    ENS_CUDA_CALL(cudaFuncSetAttribute(decode_tc_kernel<LEVEL, CD, NO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decode_tc_kernel<LEVEL, CD, NO, false><<<g, 256, smem, s>>>(a);
  }
  ENS_CHECK_CUDA();
  return ENS_OK;
}

// NICE.forward / Renderer.eval_points for stages middle, fine, color: one launch per decoder of the stage.
// msave / mstride: optional saved-mask buffer of the stage (decoder d = middle, fine, colour at msave + d * mstride words).
int tc_eval_points(const DevScene &sc, int stage, const void *pts, int pts_is_f64, int64_t n, int apply_mask,
                   float *out4, cudaStream_t s, uint32_t *msave, int64_t mstride, float *rsave, int64_t rstride) {
  if (stage == ENS_STAGE_COARSE) return ENS_EUNSUPPORTED;
  int rc = launch_decode_tc<ENS_LEVEL_MIDDLE, 32, 1>(sc, pts, pts_is_f64, n, stage == ENS_STAGE_MIDDLE ? apply_mask : 0, out4, s, msave, rsave);
  if (rc != ENS_OK || stage == ENS_STAGE_MIDDLE) return rc;
  rc = launch_decode_tc<ENS_LEVEL_FINE, 64, 1>(sc, pts, pts_is_f64, n, stage == ENS_STAGE_FINE ? apply_mask : 0, out4, s,
                                               msave ? msave + mstride : nullptr, rsave ? rsave + rstride : nullptr);
  if (rc != ENS_OK || stage == ENS_STAGE_FINE) return rc;
  return launch_decode_tc<ENS_LEVEL_COLOR, 32, 4>(sc, pts, pts_is_f64, n, apply_mask, out4, s, msave ? msave + 2 * mstride : nullptr,
                                                  rsave ? rsave + 2 * rstride : nullptr);
}

// ---------------------------------------------------------------------------------------------
// Renderer.render_batch_ray forward WITHOUT a backward (render_img, visualisation, mesh colours) on the tcgen05 decode:
//   place_kernel      sample placement (float64, same code as the fused kernels) -> z [R][S], points [R*S][3] f64
//   decode_tc_kernel  one launch per decoder, bound rule included               -> raw [R][S][4]
//   composite_kernel  raw2outputs_nerf_color (common.py:256-297), one thread per ray, the reference's sequential order
// The three per-point arrays live in a caller-provided scratch (ens_fwd_scratch_bytes).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT_MMA) place_kernel(DevScene sc, RayArgs ra, double *__restrict__ z_out,
                                                       double *__restrict__ pts) {
  __shared__ double zc[NT_MMA], zs[NT_MMA];
  const int S = ra.S;
  const int rl = threadIdx.x / S, s = threadIdx.x % S;
  const int64_t ray = (int64_t)blockIdx.x * ra.rpc + rl;
  const bool valid = (rl < ra.rpc) && (ray < ra.R);
  float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
  if (valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ra.rays_o[ray * 3 + k]; d[k] = ra.rays_d[ray * 3 + k]; }
  }
  const double z = place_sample(ra, sc, valid, ray, rl, s, o, d, zc, zs);
  if (valid) {
    const int64_t pi = ray * S + s;
    z_out[pi] = z;
#pragma unroll
    for (int k = 0; k < 3; ++k) pts[pi * 3 + k] = __dadd_rn((double)o[k], __dmul_rn((double)d[k], z));   // Renderer.py:173-174
  }
}

__global__ void __launch_bounds__(128) composite_kernel(float4 *__restrict__ raw, const double *__restrict__ z,
                                                        int64_t R, int S, double *__restrict__ depth,
                                                        double *__restrict__ var, float *__restrict__ color,
                                                        float *__restrict__ w_out, const float4 *__restrict__ planes,
                                                        int n_planes) {
  const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= R) return;
  float4 *rr = raw + ray * S;
  const double *zz = z + ray * S;
  if (planes != nullptr) {
    // the decoders wrote separate planes (decode_tc_multi_kernel): raw = (rgb of the colour decoder, fine_occ + middle_occ),
    // occupancy 100 outside the bound (Renderer.py:58) -- the same values the sequential launches leave in `raw`
    const int64_t P = R * (int64_t)S;
    for (int k = 0; k < S; ++k) {
      const int64_t pi = ray * S + k;
      const float4 m = planes[pi];
      float4 r4 = make_float4(0.f, 0.f, 0.f, m.w);
      if (n_planes > 1) r4.w = __fadd_rn(planes[P + pi].w, m.w);
      if (n_planes > 2) { const float4 c4 = planes[2 * P + pi]; r4.x = c4.x; r4.y = c4.y; r4.z = c4.z; }
      if (m.x == 0.f) r4.w = 100.f;
      rr[k] = r4;
    }
  }
  float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
  double dep = 0.0;
  for (int k = 0; k < S; ++k) {                       // sequential cumprod / sums, exactly the fused kernels' order
    const float4 rk = rr[k];
    const float alpha = 1.f / (1.f + expf(-(10.f * rk.w)));
    const float wk = __fmul_rn(alpha, T);
    T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
    dep = __dadd_rn(dep, __dmul_rn((double)wk, zz[k]));
    cr = __fadd_rn(cr, __fmul_rn(wk, rk.x)); cg = __fadd_rn(cg, __fmul_rn(wk, rk.y)); cb = __fadd_rn(cb, __fmul_rn(wk, rk.z));
    if (w_out) w_out[ray * S + k] = wk;
  }
  double v = 0.0;
  T = 1.f;
  for (int k = 0; k < S; ++k) {
    const float alpha = 1.f / (1.f + expf(-(10.f * rr[k].w)));
    const float wk = __fmul_rn(alpha, T);
    T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
    const double tmp = __dsub_rn(zz[k], dep);
    v = __dadd_rn(v, __dmul_rn(__dmul_rn((double)wk, tmp), tmp));
  }
  depth[ray] = dep;
  var[ray] = v;
  color[ray * 3 + 0] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
}

// The same compositing for SMALL batches (tracking: 200 rays would be two CTAs of the kernel above, every thread walking
// 48 dependent global loads three times): one warp per ray.  The lanes fetch and combine the samples and evaluate the
// sigmoids in parallel into shared memory; lane 0 then runs the sequential products and sums in exactly the order of
// the kernel above (bit-identical results); the lanes write the weights back.
__global__ void __launch_bounds__(128) composite_warp_kernel(float4 *__restrict__ raw, const double *__restrict__ z,
                                                             int64_t R, int S, double *__restrict__ depth,
                                                             double *__restrict__ var, float *__restrict__ color,
                                                             float *__restrict__ w_out, const float4 *__restrict__ planes,
                                                             int n_planes) {
  __shared__ float4 s_raw[4][ENS_MAX_SAMPLES];
  __shared__ double s_z[4][ENS_MAX_SAMPLES];
  __shared__ float s_a[4][ENS_MAX_SAMPLES];
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + wi;
  if (ray >= R) return;
  const int64_t P = R * (int64_t)S;
  for (int k = lane; k < S; k += 32) {
    const int64_t pi = ray * S + k;
    float4 r4;
    if (planes != nullptr) {
      const float4 m = planes[pi];
      r4 = make_float4(0.f, 0.f, 0.f, m.w);
      if (n_planes > 1) r4.w = __fadd_rn(planes[P + pi].w, m.w);
      if (n_planes > 2) { const float4 c4 = planes[2 * P + pi]; r4.x = c4.x; r4.y = c4.y; r4.z = c4.z; }
      if (m.x == 0.f) r4.w = 100.f;
      raw[pi] = r4;
    } else {
      r4 = raw[pi];
    }
    s_raw[wi][k] = r4;
    s_z[wi][k] = z[pi];
    s_a[wi][k] = 1.f / (1.f + expf(-(10.f * r4.w)));
  }
  __syncwarp();
  if (lane == 0) {
    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
    double dep = 0.0;
    for (int k = 0; k < S; ++k) {
      const float4 rk = s_raw[wi][k];
      const float alpha = s_a[wi][k];
      const float wk = __fmul_rn(alpha, T);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
      dep = __dadd_rn(dep, __dmul_rn((double)wk, s_z[wi][k]));
      cr = __fadd_rn(cr, __fmul_rn(wk, rk.x)); cg = __fadd_rn(cg, __fmul_rn(wk, rk.y)); cb = __fadd_rn(cb, __fmul_rn(wk, rk.z));
      s_a[wi][k] = wk;                                  // alpha is not needed again: keep the weight
    }
    double v = 0.0;
    for (int k = 0; k < S; ++k) {
      const double tmp = __dsub_rn(s_z[wi][k], dep);
      v = __dadd_rn(v, __dmul_rn(__dmul_rn((double)s_a[wi][k], tmp), tmp));
    }
    depth[ray] = dep;
    var[ray] = v;
    color[ray * 3 + 0] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
  }
  __syncwarp();
  if (w_out)
    for (int k = lane; k < S; k += 32) w_out[ray * S + k] = s_a[wi][k];
}

int64_t tc_fwd_scratch_bytes(int64_t n_rays, int S, int stage) {
  if (stage == ENS_STAGE_COARSE || n_rays <= 0 || S < 1) return 0;
  return n_rays * (int64_t)S * (24 + 8 + 16 + 3 * 16);      // points, z, raw, one output plane per decoder
}

static bool use_tc_multi() {
  const char *v = std::getenv("ENS_TC_MULTI");
  return !(v && v[0] == '0');
}

template <int STAGE>
static int launch_decode_tc_multi(const FwdArgs &a, const double *pts, int64_t P, float *planes, cudaStream_t s, int sms) {
  constexpr int ndec = (STAGE == ENS_STAGE_FINE) ? 2 : 3;
  TcMultiArgs m;
  m.base.sc = a.sc; m.base.pts = pts; m.base.n = P; m.base.out4 = nullptr; m.base.apply_mask = 0; m.base.msave = nullptr;
  m.base.rsave = nullptr;
  m.base.separate = 1;
  m.base.nctas = 0;
  m.base.nopipe = tc_nopipe();
  const int64_t rstride = ((P + 127) / 128) * TC_RSAVE_TILE_FLOATS;
  for (int d = 0; d < 3; ++d) {
    m.out[d] = planes + (int64_t)d * P * 4;
    m.msave[d] = a.save_masks ? a.save_masks + (int64_t)d * a.n_tiles * 160 : nullptr;
    m.rsave[d] = a.save_r ? a.save_r + (int64_t)d * rstride : nullptr;
  }
  const size_t smem = (size_t)(MlpPackTC<64>::total() + 8 * 1024) * 4;       // the fine decoder's blob is the largest
  ENS_CUDA_CALL(cudaFuncSetAttribute(decode_tc_multi_kernel<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t pairs = ((P + 127) / 128 + 1) / 2;
  int gx = 0;
  if (pairs * ndec <= sms) {
    for (int d = 0; d < 3; ++d) m.ctas[d] = d < ndec ? (int)pairs : 0;
    gx = (int)pairs;
  } else {
    // persistent: SMs in proportion to the decoders' cost per tile (ncu, 1000-ray colour-stage forward: 38.8 / 50.8 / 41.5 us)
    const double cost[3] = {1.0, 1.31, 1.07};
    double tot = 0.0;
    for (int d = 0; d < ndec; ++d) tot += cost[d];
    for (int d = 0; d < 3; ++d) {
      int64_t n = d < ndec ? (int64_t)(sms * cost[d] / tot) : 0;
      if (d < ndec && n < 1) n = 1;
      if (n > pairs) n = pairs;
      m.ctas[d] = (int)n;
      if ((int)n > gx) gx = (int)n;
    }
  }
  decode_tc_multi_kernel<STAGE><<<dim3((unsigned)gx, ndec), 256, smem, s>>>(m);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

int tc_render_fwd(FwdArgs &a, int stage, void *scratch, int64_t scratch_bytes, cudaStream_t s) {
  if (stage == ENS_STAGE_COARSE) return ENS_EUNSUPPORTED;
  const int64_t P = a.ra.R * (int64_t)a.ra.S;
  if (!scratch || scratch_bytes < tc_fwd_scratch_bytes(a.ra.R, a.ra.S, stage)) return ENS_EUNSUPPORTED;
  char *base = reinterpret_cast<char *>(scratch);
  double *pts = reinterpret_cast<double *>(base);
  double *z = a.z_out ? a.z_out : reinterpret_cast<double *>(base + P * 24);
  float *raw = a.raw_out ? a.raw_out : reinterpret_cast<float *>(base + P * 32);
  a.ra.rpc = NT_MMA / a.ra.S;
  const unsigned g = (unsigned)((a.ra.R + a.ra.rpc - 1) / a.ra.rpc);
  place_kernel<<<g, NT_MMA, 0, s>>>(a.sc, a.ra, z, pts);
  ENS_CHECK_CUDA();
  // every decoder of the stage side by side in one launch (small batch: all tiles resident at once; otherwise persistent
  // CTAs dealt to the decoders by cost), outputs combined by the compositing kernel; ENS_TC_MULTI=0: one launch per decoder
  const float4 *planes = nullptr;
  int n_planes = 0;
  const int sms = sm_count();
  const int ndec = stage == ENS_STAGE_MIDDLE ? 1 : (stage == ENS_STAGE_FINE ? 2 : 3);
  const int64_t pairs = ((P + 127) / 128 + 1) / 2;
  // (measured: one launch wins up to a few tile rounds per SM -- 0.153 -> 0.128 ms at 1000 rays -- and loses 3-10 % to the
  // per-decoder launches from 16 k rays on, where the static split of the SMs matters more than the round quantisation)
  if (ndec > 1 && pairs * ndec <= 8 * (int64_t)sms && use_tc_multi()) {
    float *pl = reinterpret_cast<float *>(base + P * 48);
    const int rc = stage == ENS_STAGE_FINE ? launch_decode_tc_multi<ENS_STAGE_FINE>(a, pts, P, pl, s, sms)
                                           : launch_decode_tc_multi<ENS_STAGE_COLOR>(a, pts, P, pl, s, sms);
    if (rc != ENS_OK) return rc;
    planes = reinterpret_cast<const float4 *>(pl);
    n_planes = ndec;
  } else {
    const int rc = tc_eval_points(a.sc, stage, pts, 1, P, 1, raw, s, a.save_masks, a.n_tiles * 160, a.save_r,
                                  ((P + 127) / 128) * TC_RSAVE_TILE_FLOATS);
    if (rc != ENS_OK) return rc;
  }
  if (a.ra.R <= 4096 && a.ra.S <= ENS_MAX_SAMPLES)
    composite_warp_kernel<<<(unsigned)((a.ra.R + 3) / 4), 128, 0, s>>>(reinterpret_cast<float4 *>(raw), z, a.ra.R, a.ra.S,
                                                                       a.depth, a.var, a.color, a.w_out, planes, n_planes);
  else
    composite_kernel<<<(unsigned)((a.ra.R + 127) / 128), 128, 0, s>>>(reinterpret_cast<float4 *>(raw), z, a.ra.R, a.ra.S,
                                                                      a.depth, a.var, a.color, a.w_out, planes, n_planes);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

}  // namespace ens
