// Renderer.render_batch_ray BACKWARD on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in TMEM):
// autograd of /root/reference/src/utils/Renderer.py:64-199 + src/conv_onet/models/decoder.py:177-203, 312-342 +
// src/common.py:256-297, stages middle / fine / color.
//
//   place_kernel          sample placement (float64, the fused kernels' code)           -> z [R][S], points [R*S][3] f64
//   composite_bwd_kernel  raw2outputs_nerf_color backward, one warp per ray             -> g_out [R*S][4] = (g_rgb, g_occ)
//   bwd_tc_kernel         one persistent CTA per SM, blockIdx.y = ROLE (a decoder), 128-point tiles, thread = point = TMEM lane:
//       data gradients    g_h_{i-1} = g_u_i Wh_i,  g_c += g_u_i M_{i-1},  g_e += g_u_{0,3} W_{0,3e}
//                         TS form: A = g_u (value + TF32 remainder) written to TMEM with tcgen05.st, B = the transposed
//                         matrices of MlpPackTCB in shared memory (K-major), 3xTF32, M = 128 points;
//       weight gradients  S_i = sum_pt g_u_i^T r_{i-1},  Q_i = sum_pt g_u_i^T c,  E_i = sum_pt g_u_i^T e   (K = the POINTS)
//                         SS form, both operands MN-major (SWIZZLE_128B_BASE32B): every thread stages its rows of
//                         [r_{i-1} | c] (or Fourier chunks) and of g_u, value and remainder.  The four 32-column blocks
//                         [x_hi | c_hi | x_lo | c_lo] are ONE M = 128 operand, so the TMEM lanes 64..127 of an accumulator
//                         hold the remainder products and two passes (B = g_hi, B = g_lo) give the 3xTF32 sum.  The
//                         accumulators stay in TMEM for ALL tiles of the CTA and are flushed once (RED.ADD) at the end;
//       tail              trilinear backward (8 lanes per 128-byte voxel line, red.global.add.v4.f32) and coordinate
//                         gradient, Fourier backward, d L / d p of the point.
//   unfold_kernel         the sums above are those of the FOLDED network the tcgen05 forward evaluates (r_i = relu(u_i),
//                         feature injection folded into the next block); by linearity the reference's parameter gradients are
//                             dW_i   = S_i + Q_i Wc_{i-1}^T + b^_i bc_{i-1}^T        dWc_i = Wh_{i+1}^T Q_{i+1}
//                             dbc_i  = Wh_{i+1}^T b^_{i+1}                           dWo   = (sum g_out^T r_4) + Qo Wc_4^T + dbo bc_4^T
//                         -- a few 32x32 products per decoder, one tiny launch.
//   rays_reduce_kernel    g_o = sum_s g_p, g_d = sum_s z_s g_p
//
// Nothing per-point crosses HBM between the data-gradient and the weight-gradient GEMMs (no g_h scratch): what the
// backward reads per point is what the forward saved -- the relu outputs r_0..r_4 (decoder gradients wanted: saved kind 3)
// or one mask word per block (pose only: saved kind 2).
// Roles: the fine decoder's feature vector is 64 wide ([fine | middle], decoder.py:182-187); its middle half needs five
// more accumulators than a CTA's TMEM holds, so a fourth role (FINE_CM) repeats the fine decoder's hidden chain (Wh only)
// and accumulates  Q_i[:, 32:64].
#include <cstdlib>
#include "ens_tc.cuh"

namespace ens {

enum { ROLE_MIDDLE = 0, ROLE_FINE = 1, ROLE_COLOR = 2, ROLE_FINE_CM = 3 };

// raw (folded-network) weight-gradient sums of one role, floats
constexpr int RAW_ACC = 8 * 2048;                 // 8 accumulators [64 rows][32 units]
constexpr int RAW_BHAT = RAW_ACC;                 // [5][32]   sum g_u_i
constexpr int RAW_DWOR = RAW_BHAT + 160;          // [4][32]   sum g_out[o] r_4[k]
constexpr int RAW_QO = RAW_DWOR + 128;            // [4][32]   sum g_out[o] c[ch]
constexpr int RAW_DBO = RAW_QO + 128;             // [4]
constexpr int RAW_DB = RAW_DBO + 4;               // [3][96]   sum p[r] g_q[k]
constexpr int RAW_FLOATS = ((RAW_DB + 288 + 63) / 64) * 64;

struct BwdTcArgs {
  DevScene sc;
  const double *pts;        // [P][3]
  const float4 *gout;       // [P]  (g_r, g_g, g_b, g_occ)
  int64_t P, n_tiles;       // n_tiles = ceil(P / 128)
  const float *save_r;      // [3 decoders][n_tiles][5][128][32]   (decoder gradients wanted)
  const uint32_t *save_m;   // [3 decoders][m_stride]: per 32-point tile [5][32] mask words (pose only)
  int64_t m_stride;
  float *ggrid[4];
  float *raw_acc;           // [4 roles][RAW_FLOATS]
  float *gp;                // [3 planes][P][3], or null
  int ctas[4];              // persistent CTAs per role
  int exp_flags;            // ENS_WGRAD_EXP (timing experiments on wgrad_tc_kernel only; results are invalid when set):
                            // 1 = producers stage the first tile only, 2 = issue the g_hi MMAs only, 4 = no MMAs at all,
                            // 8 = no sin, 16 = no row loads, 32 = no feature gather / output-layer sums
  float *gu_buf;            // split backward: g_u rows, [3 decoders][n_tiles][5 blocks][128][32], data kernel -> wgrad kernel
  long long *dbg;           // optional timestamps (tools/time_passes.py)
};

// TMEM columns
constexpr int TB_XH = 0, TB_XL = 32, TB_DH = 64, TB_DC = 96, TB_DE = 128, TB_ACC = 224;

__device__ __forceinline__ void grp_sync128(int grp) { asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory"); }

// column sums over the warp: lane k gets sum_lanes v[k]  (reduce-scatter: 16 + 8 + 4 + 2 + 1 shuffles)
__device__ __forceinline__ float warp_colsum32(const float (&v)[32]) {
  const int lane = threadIdx.x & 31;
  float a16[16], a8[8], a4[4], a2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float mine = h16 ? v[16 + j] : v[j], send = h16 ? v[j] : v[16 + j];
    a16[j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mine = h8 ? a16[8 + j] : a16[j], send = h8 ? a16[j] : a16[8 + j];
    a8[j] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float mine = h4 ? a8[4 + j] : a8[j], send = h4 ? a8[j] : a8[4 + j];
    a4[j] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float mine = h2 ? a4[2 + j] : a4[j], send = h2 ? a4[j] : a4[2 + j];
    a2[j] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float mine = h1 ? a2[1] : a2[0], send = h1 ? a2[0] : a2[1];
  return mine + __shfl_xor_sync(0xffffffffu, send, 1);
}

// stage one point's 32 values as row `pt` of an MN-major [128][32] block pair (value, TF32 remainder).
// (Lanes pt and pt + 4 of a quarter-warp store hit the same banks -- a 2-way conflict on every store.  Swapping the two
// 16-byte halves for lanes with bit 2 of pt set removes it at the price of 8 selects per store pair; measured: no gain.)
__device__ __forceinline__ void stage_row(float *__restrict__ hi, float *__restrict__ lo, int pt, const float (&v)[32]) {
  float *rh = hi + pt * 32, *rl = lo + pt * 32;
  const int sw = pt & 3;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 h = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    float4 l;
    l.x = h.x - __uint_as_float(__float_as_uint(h.x) & 0xffffe000u);
    l.y = h.y - __uint_as_float(__float_as_uint(h.y) & 0xffffe000u);
    l.z = h.z - __uint_as_float(__float_as_uint(h.z) & 0xffffe000u);
    l.w = h.w - __uint_as_float(__float_as_uint(h.w) & 0xffffe000u);
    const int off = (((q >> 1) ^ sw) << 3) | ((q & 1) << 2);        // 32-byte chunk (q / 2) ^ (pt % 4), half q % 2
    *reinterpret_cast<float4 *>(rh + off) = h;
    *reinterpret_cast<float4 *>(rl + off) = l;
  }
}

// one weight-gradient pass: acc[128 x 32] += [A blocks 0..3]^T-by-points x g  (B = value, then remainder)
__device__ __forceinline__ void issue_wgrad(uint32_t acc, uint32_t sA, uint32_t sB) {
  constexpr uint32_t IDESC = tc_idesc_mn(128, 32);
  const uint64_t da = umma_desc_mn(sA, 16384u), dbh = umma_desc_mn(sB, 16384u), dbl = umma_desc_mn(sB + 16384u, 16384u);
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) umma_ss(acc, da + (uint64_t)(64 * ks), dbl + (uint64_t)(64 * ks), IDESC, 1u);
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) umma_ss(acc, da + (uint64_t)(64 * ks), dbh + (uint64_t)(64 * ks), IDESC, 1u);
}

// SPLIT (with !WG): the data-gradient half of the split mapping backward -- the pose-only chain, which additionally writes
// every block's g_u row for wgrad (bwd_tc_wg_kernel<STAGE, true>) and keeps the sums that need no GEMM (bias sums, dB).
template <int ROLE, bool WG, bool SPLIT = false>
__device__ __forceinline__ void bwd_tc_body(const BwdTcArgs &a, float *smem_raw) {
  constexpr int LEVEL = (ROLE == ROLE_MIDDLE) ? ENS_LEVEL_MIDDLE : (ROLE == ROLE_COLOR ? ENS_LEVEL_COLOR : ENS_LEVEL_FINE);
  constexpr int CD = (LEVEL == ENS_LEVEL_FINE) ? 64 : 32;
  constexpr int NO = (ROLE == ROLE_COLOR) ? 3 : 1;              // outputs that carry gradient (decoder.py:341 drops colour's 4th)
  constexpr int DEC = (ROLE == ROLE_MIDDLE) ? 0 : (ROLE == ROLE_COLOR ? 2 : 1);
  constexpr bool TAIL = ROLE != ROLE_FINE_CM;                    // scatter / embedding / d L / d p
  constexpr int CLEVEL = (ROLE == ROLE_FINE_CM) ? ENS_LEVEL_MIDDLE : LEVEL;     // level whose features fill slot B
  using PB = MlpPackTCB;
  static_assert(!(ROLE == ROLE_FINE_CM) || WG, "the FINE_CM role only exists for decoder gradients");

  // ---- shared memory: [M-side 4 blocks | N-side 2 blocks] (WG) or one staging block, then the weight blob ----
  float *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;    // swizzle atoms need SHARED-space alignment
  // Pose-only form (!WG): TWO 128-point tiles in flight per CTA, as in the decode kernel -- tile group 0 = warps 0-3, group
  // 1 = warps 4-7, each with its own 256 TMEM columns, staging block, mbarrier and named barrier; one group's tail and
  // epilogues overlap the other's MMAs (4 warps per SM could not hide anything).
  const int tid_all = threadIdx.x;
  const int grp = WG ? 0 : (tid_all >> 7);
  float *sMB = base + (WG ? 0 : grp * 4096);           // hiA, hiB, loA, loB   (hiA doubles as the warps' 32x32 staging tiles)
  float *sNB = base + (WG ? 16384 : 8192);             // g_hi, g_lo (WG only)
  float *sw = sNB + (WG ? 8192 : 0);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 sP[256];                           // p.float() of the tile's points (dB reduction), per tile group
  const int tid = tid_all & 127, warp = tid >> 5, lane = tid & 31;      // thread / warp WITHIN the tile group
  uint64_t *const barp = &bars[grp];
  float *stile = sMB + warp * 1024;
  constexpr uint32_t TCOLS = 512u;

  {
    const float *gw = a.sc.w[LEVEL] + off_tcb<CD>();
    const uint32_t s0 = smem_u32(sw);
    for (int i = tid_all; i < PB::total() / 4; i += (int)blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + i * 16), "l"(gw + i * 4) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  if (tid_all == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((tid_all >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb0 = tmem_base_s + (WG ? 0u : (uint32_t)(256 * grp));
  const uint32_t tb = tb0 + ((uint32_t)(32 * warp) << 16);
  const uint32_t swb = smem_u32(sw), sMBa = smem_u32(sMB), sNBa = smem_u32(sNB);
  uint32_t parity = 0;

  if (WG) {      // the weight-gradient accumulators start at zero and only ever accumulate
    uint32_t zz[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) zz[k] = 0u;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                   "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                   "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                   :: "r"(tb + TB_ACC + 32 * c), "r"(zz[0]), "r"(zz[1]), "r"(zz[2]), "r"(zz[3]), "r"(zz[4]), "r"(zz[5]), "r"(zz[6]), "r"(zz[7]),
                     "r"(zz[8]), "r"(zz[9]), "r"(zz[10]), "r"(zz[11]), "r"(zz[12]), "r"(zz[13]), "r"(zz[14]), "r"(zz[15]),
                     "r"(zz[16]), "r"(zz[17]), "r"(zz[18]), "r"(zz[19]), "r"(zz[20]), "r"(zz[21]), "r"(zz[22]), "r"(zz[23]),
                     "r"(zz[24]), "r"(zz[25]), "r"(zz[26]), "r"(zz[27]), "r"(zz[28]), "r"(zz[29]), "r"(zz[30]), "r"(zz[31])
                   : "memory");
    }
    tmem_st_done();
    tc_fence_before();
    grp_sync128(grp);
    tc_fence_after();
  }

  // per-lane sums that live in registers across the CTA's tiles (lane k = column k; one partial per warp)
  float bhat[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float dwor[NO], qo[NO], dbo[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) dwor[o] = qo[o] = dbo[o] = 0.f;
  float dBacc[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};

  const bool want_rays = TAIL && a.gp != nullptr;
  const int nctas = a.ctas[ROLE];
  constexpr int NGRP = WG ? 1 : 2;
  for (int64_t tile = (int64_t)blockIdx.x * NGRP + grp; tile < a.n_tiles; tile += (int64_t)nctas * NGRP) {
    const int64_t pt = tile * 128 + tid;
    const bool valid = pt < a.P;
    double p[3] = {0.0, 0.0, 0.0};
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const double *pp = a.pts + pt * 3;
      p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2];
      g4 = a.gout[pt];
    }
    float pn[3], p32[3];
    normalize64(p, a.sc.lo, a.sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k) p32[k] = __double2float_rn(p[k]);
    float go[NO];
    if (ROLE == ROLE_COLOR) { go[0] = g4.x; if (NO > 1) { go[1 % NO] = g4.y; go[2 % NO] = g4.z; } }
    else go[0] = g4.w;
    const Vox vox = make_vox(pn, a.sc.dims[LEVEL]);

    // ---- features of the point (operand of Q_i = sum g_u_i^T c) -> slot B, value + remainder ----
    float c[32];
    if (WG) {
      const Vox vc = (CLEVEL == LEVEL) ? vox : make_vox(pn, a.sc.dims[CLEVEL]);
      __syncwarp();
      gather_warp<32>(a.sc.grid[CLEVEL], a.sc.dims[CLEVEL], vc, stile, 0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = *reinterpret_cast<const float4 *>(stile + lane * 32 + ((4 * q) ^ ((lane & 3) << 3)));
        c[4 * q] = x.x; c[4 * q + 1] = x.y; c[4 * q + 2] = x.z; c[4 * q + 3] = x.w;
      }
      __syncwarp();
      stage_row(sMB + 4096, sMB + 3 * 4096, tid, c);
    }
    if ((WG || SPLIT) && TAIL) sP[128 * grp + tid] = make_float4(p32[0], p32[1], p32[2], 0.f);

    // ---- g_h4 = Wo^T g_out ----
    float g[32];
    {
      const float *Wo = sw + PB::off_Wo();
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < NO; ++o) s = fmaf(Wo[o * 32 + k], go[o], s);
        g[k] = s;
      }
    }
    const float *rbase = WG ? a.save_r + (((int64_t)DEC * a.n_tiles + tile) * 5) * 4096 + tid * 32 : nullptr;
    float rcur[32];
    if (WG) {
      const float4 *src = reinterpret_cast<const float4 *>(rbase + 4 * 4096);
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float4 x = __ldg(src + q); rcur[4 * q] = x.x; rcur[4 * q + 1] = x.y; rcur[4 * q + 2] = x.z; rcur[4 * q + 3] = x.w; }
      // output layer: sum g_out[o] r_4[k], sum g_out[o] c[ch], sum g_out[o]
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        float t[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = go[o] * rcur[k];
        dwor[o] += warp_colsum32(t);
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = go[o] * c[k];
        qo[o] += warp_colsum32(t);
        float sg = go[o];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, off);
        dbo[o] += sg;
      }
    }

    // ---- blocks 4..0 ----
#pragma unroll 1
    for (int i = 4; i >= 0; --i) {
      float gu[32];
      if (WG) {
#pragma unroll
        for (int k = 0; k < 32; ++k) gu[k] = (rcur[k] > 0.f) ? g[k] : 0.f;
      } else {
        const uint32_t mw = valid ? a.save_m[(int64_t)DEC * a.m_stride + (pt >> 5) * 160 + i * 32 + (pt & 31)] : 0u;
#pragma unroll
        for (int k = 0; k < 32; ++k) gu[k] = ((mw >> k) & 1u) ? g[k] : 0.f;
      }
      if (SPLIT) {
        // the row the weight-gradient kernel multiplies with [r_{i-1} | c | e]: 128 B per thread, stays in L2
        float4 *dst = reinterpret_cast<float4 *>(a.gu_buf + ((((int64_t)DEC * a.n_tiles + tile) * 5 + i) * 128 + tid) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(gu[4 * q], gu[4 * q + 1], gu[4 * q + 2], gu[4 * q + 3]);
        const float bs = warp_colsum32(gu);
#pragma unroll
        for (int k = 0; k < 5; ++k) if (k == i) bhat[k] += bs;
      }
      if (WG) {
        const float bs = warp_colsum32(gu);
#pragma unroll
        for (int k = 0; k < 5; ++k) if (k == i) bhat[k] += bs;
        stage_row(sNB, sNB + 4096, tid, gu);
        if (i >= 1) {                                   // slot A: r_{i-1}, which is also the next block's relu mask
          const float4 *src = reinterpret_cast<const float4 *>(rbase + (i - 1) * 4096);
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float4 x = __ldg(src + q); rcur[4 * q] = x.x; rcur[4 * q + 1] = x.y; rcur[4 * q + 2] = x.z; rcur[4 * q + 3] = x.w; }
          stage_row(sMB, sMB + 2 * 4096, tid, rcur);
        } else if (TAIL) {                              // block 0: slot A = Fourier chunk 0
          float e[32];
          const float *B = sw + PB::off_B();
#pragma unroll
          for (int k = 0; k < 32; ++k) e[k] = fast_sin(fmaf(p32[2], B[2 * EMBP + k], fmaf(p32[1], B[EMBP + k], p32[0] * B[k])));
          stage_row(sMB, sMB + 2 * 4096, tid, e);
        }
      }
      const bool need_a = (i >= 1) || TAIL;             // block 0's g_u only feeds the embedding gradient
      if (need_a) tmem_st32_split(tb + TB_XH, tb + TB_XL, gu);
      tmem_st_done();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      grp_sync128(grp);
      if (warp == 0) {
        tc_fence_after();
        if (i >= 1) {
          issue_gemm<32, 32>(tb0 + TB_DH, tb0 + TB_XH, tb0 + TB_XL, swb, PB::off_WhT(i), PB::TOT(), 0u);
          if (TAIL) issue_gemm<32, 32>(tb0 + TB_DC, tb0 + TB_XH, tb0 + TB_XL, swb, PB::off_MT(i - 1), PB::TOT(), i == 4 ? 0u : 1u);
        }
        if (TAIL && (i == 3 || i == 0))
          issue_gemm<32, 32, 96>(tb0 + TB_DE, tb0 + TB_XH, tb0 + TB_XL, swb, i == 3 ? PB::off_W3eT() : PB::off_W0T(), PB::TOT(), i == 3 ? 0u : 1u);
        if (WG) {
          // accumulators: 0 L4 [r3|c]  1 L3 [r2|c]  2 L3 [e0|e1]  3 L3 [e2|c]  4 L2 [r1|c]  5 L1 [r0|c]  6 L0 [e0|c]  7 L0 [e1|e2]
          // (FINE_CM: 4 - i, rows 32..63 = c_middle)
          const int acc = (ROLE == ROLE_FINE_CM) ? (4 - i) : (i == 4 ? 0 : (i == 3 ? 1 : (i == 2 ? 4 : (i == 1 ? 5 : 6))));
          issue_wgrad(tb0 + TB_ACC + 32 * acc, sMBa, sNBa);
        }
        umma_commit(barp);
        __syncwarp();
      }
      mbar_wait(barp, parity); parity ^= 1;
      tc_fence_after();
      if (i >= 1) tmem_ld32(tb + TB_DH, g);
      // ---- the Fourier-feature operands of blocks 3 and 0:  E_i = sum g_u_i^T e ----
      if (WG && TAIL && (i == 3 || i == 0)) {
        const float *B = sw + PB::off_B();
        float e[32];
        // pass 1: i == 3: [e0 | e1] (slot B loses c);  i == 0: [e1 | e2] (c is dead after block 0)
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int jc = (i == 3) ? h : 1 + h;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            e[k] = fast_sin(fmaf(p32[2], B[2 * EMBP + 32 * jc + k], fmaf(p32[1], B[EMBP + 32 * jc + k], p32[0] * B[32 * jc + k])));
          stage_row(sMB + h * 4096, sMB + (2 + h) * 4096, tid, e);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        grp_sync128(grp);
        if (warp == 0) {
          tc_fence_after();
          issue_wgrad(tb0 + TB_ACC + 32 * (i == 3 ? 2 : 7), sMBa, sNBa);
          umma_commit(barp);
          __syncwarp();
        }
        mbar_wait(barp, parity); parity ^= 1;
        tc_fence_after();
        if (i == 3) {       // pass 2: [e2 | c] -- slot B gets the features back for blocks 2, 1, 0
#pragma unroll
          for (int k = 0; k < 32; ++k)
            e[k] = fast_sin(fmaf(p32[2], B[2 * EMBP + 64 + k], fmaf(p32[1], B[EMBP + 64 + k], p32[0] * B[64 + k])));
          stage_row(sMB, sMB + 2 * 4096, tid, e);
          stage_row(sMB + 4096, sMB + 3 * 4096, tid, c);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          grp_sync128(grp);
          if (warp == 0) {
            tc_fence_after();
            issue_wgrad(tb0 + TB_ACC + 32 * 3, sMBa, sNBa);
            umma_commit(barp);
            __syncwarp();
          }
          mbar_wait(barp, parity); parity ^= 1;
          tc_fence_after();
        }
      }
    }

    // ---- tail: feature gradient -> grid scatter + coordinate gradient; embedding gradient ----
    if (TAIL) {
      double gp[3] = {0.0, 0.0, 0.0};
      float *ggrid = a.ggrid[LEVEL];
      if (ggrid != nullptr || want_rays) {
        float gc[32];
        tmem_ld32(tb + TB_DC, gc);
        const float *MoF = sw + PB::off_MoF();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          float s = gc[k];
#pragma unroll
          for (int o = 0; o < NO; ++o) s = fmaf(MoF[o * 32 + k], go[o], s);
          gc[k] = s;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4 *>(stile + lane * 32 + ((4 * q) ^ ((lane & 3) << 3))) = make_float4(gc[4 * q], gc[4 * q + 1], gc[4 * q + 2], gc[4 * q + 3]);
        __syncwarp();
        float gpn[3];
        gather_bwd_warp<32>(a.sc.grid[LEVEL], ggrid, a.sc.dims[LEVEL], vox, valid, stile, want_rays, gpn);
        if (want_rays) {
#pragma unroll
          for (int k = 0; k < 3; ++k) gp[k] += __ddiv_rn((double)gpn[k] * 2.0, __dsub_rn(a.sc.hi[k], a.sc.lo[k]));
        }
        __syncwarp();
      }
      if (WG || SPLIT || want_rays) {
        const float *B = sw + PB::off_B();
        float gpe[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int jc = 0; jc < 3; ++jc) {
          float ge[32];
          tmem_ld32(tb + TB_DE + 32 * jc, ge);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float bx = B[32 * jc + k], by = B[EMBP + 32 * jc + k], bz = B[2 * EMBP + 32 * jc + k];
            float sq, cq;
            fast_sincos(fmaf(p32[2], bz, fmaf(p32[1], by, p32[0] * bx)), sq, cq);
            const float gq = ge[k] * cq;
            ge[k] = gq;
            gpe[0] = fmaf(bx, gq, gpe[0]); gpe[1] = fmaf(by, gq, gpe[1]); gpe[2] = fmaf(bz, gq, gpe[2]);
          }
          if (WG || SPLIT) {
            // dB[r][32 jc + k] = sum_pt p[pt][r] g_q[pt][k]: transpose through the warp's tile, lane k walks its column
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4 *>(stile + lane * 32 + ((4 * q) ^ ((lane & 3) << 3))) = make_float4(ge[4 * q], ge[4 * q + 1], ge[4 * q + 2], ge[4 * q + 3]);
            __syncwarp();
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const float v = stile[r * 32 + (lane ^ ((r & 3) << 3))];
              const float4 pr = sP[128 * grp + 32 * warp + r];
              s0 = fmaf(pr.x, v, s0); s1 = fmaf(pr.y, v, s1); s2 = fmaf(pr.z, v, s2);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) if (q == jc) { dBacc[q][0] += s0; dBacc[q][1] += s1; dBacc[q][2] += s2; }
            __syncwarp();
          }
        }
        if (want_rays) {
#pragma unroll
          for (int k = 0; k < 3; ++k) gp[k] += (double)gpe[k];
        }
      }
      if (want_rays && valid) {
        float *dst = a.gp + ((int64_t)DEC * a.P + pt) * 3;          // pose-only form: one plane per decoder
        dst[0] = (float)gp[0]; dst[1] = (float)gp[1]; dst[2] = (float)gp[2];
      }
    }
    // the next tile's staging / tcgen05.st must not overtake this tile's TMEM loads and shared-tile reads
    tc_fence_before();
    grp_sync128(grp);
    tc_fence_after();
  }

  // ---- flush the CTA's sums ----
  if (WG) {
    float *raw = a.raw_acc + (int64_t)ROLE * RAW_FLOATS;
    constexpr int NACC = (ROLE == ROLE_FINE_CM) ? 5 : 8;
#pragma unroll 1
    for (int acc = 0; acc < NACC; ++acc) {
      float v[32];
      tmem_ld32(tb + TB_ACC + 32 * acc, v);
      float *dst = raw + acc * 2048 + (tid & 63) * 32;              // lanes 64..127 hold the remainder products of rows 0..63
#pragma unroll
      for (int q = 0; q < 8; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) atomicAdd(raw + RAW_BHAT + i * 32 + lane, bhat[i]);
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      atomicAdd(raw + RAW_DWOR + o * 32 + lane, dwor[o]);
      atomicAdd(raw + RAW_QO + o * 32 + lane, qo[o]);
      if (lane == 0) atomicAdd(raw + RAW_DBO + o, dbo[o]);
    }
    if (TAIL) {
#pragma unroll
      for (int jc = 0; jc < 3; ++jc)
#pragma unroll
        for (int r = 0; r < 3; ++r) atomicAdd(raw + RAW_DB + r * 96 + 32 * jc + lane, dBacc[jc][r]);
    }
  }
  if (SPLIT) {
    float *raw = a.raw_acc + (int64_t)ROLE * RAW_FLOATS;
#pragma unroll
    for (int i = 0; i < 5; ++i) atomicAdd(raw + RAW_BHAT + i * 32 + lane, bhat[i]);
    if (TAIL) {
#pragma unroll
      for (int jc = 0; jc < 3; ++jc)
#pragma unroll
        for (int r = 0; r < 3; ++r) atomicAdd(raw + RAW_DB + r * 96 + 32 * jc + lane, dBacc[jc][r]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if ((tid_all >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(TCOLS) : "memory");
}


// =============================================================================================================
// Backward WITH decoder gradients: 8 warps.  Warps 0..3 ("E") own the TMEM lanes: g_u_i, bias sums, N-side staging,
// the A operand, the tail.  Warps 4..7 ("H", warp 4 + w on the same SM sub-partition as E warp w and working on the same 32
// points) own the M-side: features, relu outputs from the forward, Fourier chunks; warp 4 issues every MMA.  All threads walk
// the same sequence of weight-gradient passes; before staging for pass p a thread waits for the MMAs of pass p - 1 (barW),
// E additionally for the data-gradient MMAs of the current block (barD, committed before the weight-gradient MMAs are
// issued) -- so the weight-gradient GEMM of a pass overlaps the E warps' work on the next block and the H warps'
// global loads / sines for the next pass.
//   passes (accumulator = pass):  0 L4 [r3|c]  1 L3 [r2|c]  2 L3 [e0|e1]  3 L3 [e2|c]  4 L2 [r1|c]  5 L1 [r0|c]  6 L0 [e0|c]  7 L0 [e1|e2]
//   FINE_CM: five passes 0..4 = blocks 4..0, slot B = the middle-level features, slot A unused.
// =============================================================================================================

__device__ __forceinline__ void load_row32(const float *__restrict__ src, float (&v)[32]) {
  const float4 *s4 = reinterpret_cast<const float4 *>(src);
#pragma unroll
  // streaming load (evict-first): the saved relu outputs are read exactly once and must not push the grids out of L2
  for (int q = 0; q < 8; ++q) { const float4 x = __ldcs(s4 + q); v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w; }
}

// tcgen05.st of 32 zero columns
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n"
               :: "r"(taddr), "r"(z) : "memory");
}

// weight-gradient pass, issued by ONE thread in a rolled loop (small code: the issuing warp shares its instruction cache
// with the staging code)
__device__ __forceinline__ void issue_wgrad_loop(uint32_t acc, uint32_t sA, uint32_t sB) {
  constexpr uint32_t IDESC = tc_idesc_mn(128, 32);
  const uint64_t da = umma_desc_mn(sA, 16384u), dbh = umma_desc_mn(sB, 16384u), dbl = umma_desc_mn(sB + 16384u, 16384u);
#pragma unroll 4
  for (int ks = 0; ks < 16; ++ks) umma_ss(acc, da + (uint64_t)(64 * ks), dbl + (uint64_t)(64 * ks), IDESC, 1u);
#pragma unroll 4
  for (int ks = 0; ks < 16; ++ks) umma_ss(acc, da + (uint64_t)(64 * ks), dbh + (uint64_t)(64 * ks), IDESC, 1u);
}

// (timing experiment) only the value half of the N side
__device__ __forceinline__ void issue_wgrad_half(uint32_t acc, uint32_t sA, uint32_t sB) {
  constexpr uint32_t IDESC = tc_idesc_mn(128, 32);
  const uint64_t da = umma_desc_mn(sA, 16384u), dbh = umma_desc_mn(sB, 16384u);
#pragma unroll 4
  for (int ks = 0; ks < 16; ++ks) umma_ss(acc, da + (uint64_t)(64 * ks), dbh + (uint64_t)(64 * ks), IDESC, 1u);
}
// the same pass spread over TWO accumulators (even / odd k-steps): consecutive MMAs into one accumulator form a dependent
// chain, and the wgrad kernel of the split backward has the TMEM columns to break it (the halves are added at the flush)
__device__ __forceinline__ void issue_wgrad_loop2(uint32_t acc0, uint32_t acc1, uint32_t sA, uint32_t sB) {
  constexpr uint32_t IDESC = tc_idesc_mn(128, 32);
  const uint64_t da = umma_desc_mn(sA, 16384u), dbh = umma_desc_mn(sB, 16384u), dbl = umma_desc_mn(sB + 16384u, 16384u);
#pragma unroll 4
  for (int ks = 0; ks < 16; ++ks) umma_ss((ks & 1) ? acc1 : acc0, da + (uint64_t)(64 * ks), dbl + (uint64_t)(64 * ks), IDESC, 1u);
#pragma unroll 4
  for (int ks = 0; ks < 16; ++ks) umma_ss((ks & 1) ? acc1 : acc0, da + (uint64_t)(64 * ks), dbh + (uint64_t)(64 * ks), IDESC, 1u);
}

// the k-step `ks` (8 of the K columns) of  D += A[128 x K] * W[N x K]^T  in 3xTF32 -- see issue_gemm
template <int KMAT, int N>
__device__ __forceinline__ void issue_gemm_k(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t sw_base, int w_off, int tot, int ks) {
  const uint64_t dh = umma_desc(sw_base + (uint32_t)w_off * 4u, (KMAT / 4) * 128u) + (uint64_t)(16 * ks);
  const uint64_t dl = umma_desc(sw_base + (uint32_t)(w_off + tot) * 4u, (KMAT / 4) * 128u) + (uint64_t)(16 * ks);
  constexpr uint32_t IDESC = tc_idesc(N);
  umma_ts(d, a_lo + 8 * ks, dh, IDESC, 1u);
  umma_ts(d, a_hi + 8 * ks, dl, IDESC, 1u);
  umma_ts(d, a_hi + 8 * ks, dh, IDESC, 1u);
}
// a quarter of a weight-gradient pass: points 8 (part + 4 j), j = 0..3, remainder then value of g
__device__ __forceinline__ void issue_wgrad_part(uint32_t acc, uint32_t sA, uint32_t sB, int part) {
  constexpr uint32_t IDESC = tc_idesc_mn(128, 32);
  const uint64_t da = umma_desc_mn(sA, 16384u) + (uint64_t)(64 * part);
  const uint64_t dbh = umma_desc_mn(sB, 16384u) + (uint64_t)(64 * part), dbl = umma_desc_mn(sB + 16384u, 16384u) + (uint64_t)(64 * part);
#pragma unroll
  for (int j = 0; j < 4; ++j) umma_ss(acc, da + (uint64_t)(256 * j), dbl + (uint64_t)(256 * j), IDESC, 1u);
#pragma unroll
  for (int j = 0; j < 4; ++j) umma_ss(acc, da + (uint64_t)(256 * j), dbh + (uint64_t)(256 * j), IDESC, 1u);
}

__device__ __forceinline__ void cta_sync288() { asm volatile("bar.sync 2, 288;" ::: "memory"); }
// barrier 3: the four E warps arrive (g_u_i is in TMEM), the issuer warp waits
__device__ __forceinline__ void cta_arrive_a() { asm volatile("bar.arrive 3, 160;" ::: "memory"); }
__device__ __forceinline__ void cta_sync_a() { asm volatile("bar.sync 3, 160;" ::: "memory"); }

// SPLIT = true is the first form of the split backward's weight-gradient kernel (this kernel without data GEMMs, masks and
// tails, single-buffered staging; 0.287 ms).  It is no longer launched -- wgrad_tc_kernel below, with two staging stages,
// replaced it -- and is kept only as the record of that measurement.
template <int ROLE, bool SPLIT = false>
__device__ __forceinline__ void bwd_tc_wg_body(const BwdTcArgs &a, float *smem_raw) {
  constexpr int LEVEL = (ROLE == ROLE_MIDDLE) ? ENS_LEVEL_MIDDLE : (ROLE == ROLE_COLOR ? ENS_LEVEL_COLOR : ENS_LEVEL_FINE);
  constexpr int CD = (LEVEL == ENS_LEVEL_FINE) ? 64 : 32;
  constexpr int NO = (ROLE == ROLE_COLOR) ? 3 : 1;
  constexpr int DEC = (ROLE == ROLE_MIDDLE) ? 0 : (ROLE == ROLE_COLOR ? 2 : 1);
  constexpr bool TAIL = ROLE != ROLE_FINE_CM;
  constexpr int CLEVEL = (ROLE == ROLE_FINE_CM) ? ENS_LEVEL_MIDDLE : LEVEL;
  constexpr int NPASS = TAIL ? 8 : 5;
  using PB = MlpPackTCB;

  float *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;
  float *sMB = base;                 // hiA, hiB, loA, loB
  float *sNB = base + 16384;         // g_hi, g_lo
  float *sw = sNB + 8192;
  __shared__ __align__(8) uint64_t barD, barW;
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 sP[128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool isE = warp < 4, isI = warp == 8;       // E: TMEM lanes; H (4..7): M-side; I (8): issues every MMA
  const int pl = tid & 127;          // point of the tile this thread works on
  const int w4 = warp & 3;

  {
    const float *gw = a.sc.w[LEVEL] + off_tcb<CD>();
    const uint32_t s0 = smem_u32(sw);
    for (int i = tid; i < PB::total() / 4; i += 288)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + i * 16), "l"(gw + i * 4) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&barD)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&barW)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb0 = tmem_base_s;
  const uint32_t tb = tb0 + ((uint32_t)(32 * w4) << 16);
  const uint32_t swb = smem_u32(sw), sMBa = smem_u32(sMB), sNBa = smem_u32(sNB);
  uint32_t pw = 0, pd = 0;           // phase parities of barW / barD as this thread has consumed them
  bool first = true;                 // this thread has no un-consumed weight-gradient pass behind it

  if (isE) {      // every accumulator starts at zero and only ever accumulates: g_h | g_c | g_e and the weight-gradient sums
#pragma unroll 1
    for (int c = TB_DH; c < TB_ACC + 32 * NPASS; c += 32) tmem_zero32(tb + c);
    tmem_st_done();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // sums that live in registers across the CTA's tiles (lane k = column k; one partial per warp)
  float bhat[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float dBacc[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float dwor[NO], qo[NO], dbo[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) dwor[o] = qo[o] = dbo[o] = 0.f;

  const bool want_rays = TAIL && a.gp != nullptr;
  const int nctas = a.ctas[ROLE];
  // the block of a pass (-1: Fourier-chunk pass without data gradients)
  auto pass_layer = [](int p) { return TAIL ? (p == 0 ? 4 : (p == 1 ? 3 : (p == 4 ? 2 : (p == 5 ? 1 : (p == 6 ? 0 : -1))))) : 4 - p; };

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += nctas) {
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tile == (int64_t)nctas && (tid == 0 || tid == 128);
#define ENS_DBG(ps, slot) do { if (dbg_on) a.dbg[((ROLE * 10 + (ps)) * 2 + (tid >> 7)) * 8 + (slot)] = clock64(); } while (0)
    ENS_DBG(8, 0);
    const int64_t pt = tile * 128 + pl;
    const bool valid = pt < a.P;
    double p[3] = {0.0, 0.0, 0.0};
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const double *pp = a.pts + pt * 3;
      p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2];
      g4 = a.gout[pt];
    }
    float pn[3], p32[3];
    normalize64(p, a.sc.lo, a.sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k) p32[k] = __double2float_rn(p[k]);
    float go[NO];
    if (ROLE == ROLE_COLOR) { go[0] = g4.x; if (NO > 1) { go[1 % NO] = g4.y; go[2 % NO] = g4.z; } }
    else go[0] = g4.w;
    const float *rbase = a.save_r + (((int64_t)DEC * a.n_tiles + tile) * 5) * 4096 + pl * 32;

    if (isI) {
      // ================================ issuer warp ================================
      // issuing an MMA blocks while the tensor pipe's queue is full, i.e. for about as long as the pass runs: nothing else
      // may sit behind it, so this warp does nothing but issue.  All GEMMs accumulate (their readers zero the accumulators).
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        const int i = pass_layer(ps);
        // the data-gradient GEMM goes out as soon as the E warps have put g_u_i into TMEM (barrier 3); they stage the same
        // rows for the weight-gradient pass while it runs
        if (!SPLIT && i >= 1) {
          cta_sync_a();
          tc_fence_after();
          // [g_h | g_c | g_e] += g_u [WhT_i ; MT_{i-1} ; W3eT]
          if (!TAIL) issue_gemm<32, 32, 32>(tb0 + TB_DH, tb0 + TB_XH, tb0 + TB_XL, swb, PB::off_G(i), PB::TOT(), 1u);
          else if (i == 3) issue_gemm<32, 32, 160>(tb0 + TB_DH, tb0 + TB_XH, tb0 + TB_XL, swb, PB::off_G(3), PB::TOT(), 1u);
          else issue_gemm<32, 32, 64>(tb0 + TB_DH, tb0 + TB_XH, tb0 + TB_XL, swb, PB::off_G(i), PB::TOT(), 1u);
          umma_commit(&barD);
        } else if (!SPLIT && i == 0 && TAIL) {
          cta_sync_a();
          tc_fence_after();
          issue_gemm<32, 32, 96>(tb0 + TB_DE, tb0 + TB_XH, tb0 + TB_XL, swb, PB::off_G(0), PB::TOT(), 1u);
          umma_commit(&barD);
        }
        tc_fence_before();
        cta_sync288();
        tc_fence_after();
        issue_wgrad_loop(tb0 + TB_ACC + 32 * ps, sMBa, sNBa);
        umma_commit(&barW);
        __syncwarp();
      }
      first = false;
    } else if (isE) {
      // ================================ E warps ================================
      const Vox vox = make_vox(pn, a.sc.dims[LEVEL]);
      uint32_t mw[5];
#pragma unroll
      for (int i = 0; i < 5; ++i)
        mw[i] = valid ? __ldg(a.save_m + (int64_t)DEC * a.m_stride + (pt >> 5) * 160 + i * 32 + (pt & 31)) : 0u;
      float g[32];
      const float *gubase = SPLIT ? a.gu_buf + (((int64_t)DEC * a.n_tiles + tile) * 5) * 4096 + pl * 32 : nullptr;
      if (SPLIT) {
        load_row32(gubase + pass_layer(0) * 4096, g);             // pass 0's g_u row (the next one is fetched a pass ahead)
      } else {
        const float *Wo = sw + PB::off_Wo();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          float s = 0.f;
#pragma unroll
          for (int o = 0; o < NO; ++o) s = fmaf(Wo[o * 32 + k], go[o], s);
          g[k] = s;
        }
      }
#pragma unroll 1
      ENS_DBG(8, 1);
      for (int ps = 0; ps < NPASS; ++ps) {
        const int i = pass_layer(ps);
        ENS_DBG(ps, 0);
        if (SPLIT && i >= 0) {
          if (!first) { mbar_wait(&barW, pw); pw ^= 1; }
          stage_row(sNB, sNB + 4096, pl, g);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          // the next block pass's row, in flight during this pass's MMAs and the Fourier passes in between
          int nl = -1;
#pragma unroll
          for (int q = NPASS - 1; q >= 0; --q) if (q > ps && pass_layer(q) >= 0) nl = pass_layer(q);
          if (nl >= 0) load_row32(gubase + nl * 4096, g);
        } else if (i >= 0) {
          uint32_t m = 0u;
#pragma unroll
          for (int k = 0; k < 5; ++k) if (k == i) m = mw[k];
#pragma unroll
          for (int k = 0; k < 32; ++k) g[k] = ((m >> k) & 1u) ? g[k] : 0.f;          // g_u_i
          if (i >= 1 || TAIL) {                                                     // A operand of the data-gradient GEMM
            tmem_st32_split(tb + TB_XH, tb + TB_XL, g);
            tmem_st_done();
            tc_fence_before();
            cta_arrive_a();
          }
          ENS_DBG(ps, 1);
          const float bs = warp_colsum32(g);
#pragma unroll
          for (int k = 0; k < 5; ++k) if (k == i) bhat[k] += bs;
          if (!first) { mbar_wait(&barW, pw); pw ^= 1; }
          ENS_DBG(ps, 2);
          stage_row(sNB, sNB + 4096, pl, g);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          ENS_DBG(ps, 3);
        } else {
          // Fourier pass: the H warps fill one slot, these warps the other (pass 2: e1 -> slot B, 3: e2 -> slot A, 7: e2 -> slot B)
          const int je = (ps == 2) ? 1 : 2;
          const float *B = sw + PB::off_B();
          float ve[32];
#pragma unroll
          for (int k = 0; k < 32; ++k)
            ve[k] = fast_sin(fmaf(p32[2], B[2 * EMBP + 32 * je + k], fmaf(p32[1], B[EMBP + 32 * je + k], p32[0] * B[32 * je + k])));
          if (!first) { mbar_wait(&barW, pw); pw ^= 1; }
          if (ps == 3) stage_row(sMB, sMB + 2 * 4096, pl, ve);
          else stage_row(sMB + 4096, sMB + 3 * 4096, pl, ve);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        first = false;
        tc_fence_before();
        cta_sync288();
        ENS_DBG(ps, 4);
        if (!SPLIT && i >= 1) {
          mbar_wait(&barD, pd); pd ^= 1;
          tc_fence_after();
          ENS_DBG(ps, 5);
          tmem_ld32(tb + TB_DH, g);
          tmem_zero32(tb + TB_DH);                    // the next block's GEMM accumulates into zero
        } else if (!SPLIT && i == 0 && TAIL) {
          mbar_wait(&barD, pd); pd ^= 1;              // g_e complete
          tc_fence_after();
        }
        ENS_DBG(ps, 6);
      }
      ENS_DBG(8, 2);
      // ---- tail: feature gradient -> grid scatter + coordinate gradient; embedding gradient ----
      if (TAIL && !SPLIT) {
        mbar_wait(&barW, pw); pw ^= 1;                // last pass done: the N-side block doubles as this warp's 32x32 tile
        first = true;                                 // ... and that phase is consumed
        tc_fence_after();
        ENS_DBG(9, 0);
        float *stile = sNB + w4 * 1024;
        double gp[3] = {0.0, 0.0, 0.0};
        float *ggrid = a.ggrid[LEVEL];
        {
          float gc[32];
          tmem_ld32(tb + TB_DC, gc);
          tmem_zero32(tb + TB_DC);
          if (ggrid != nullptr || want_rays) {
            const float *MoF = sw + PB::off_MoF();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              float s = gc[k];
#pragma unroll
              for (int o = 0; o < NO; ++o) s = fmaf(MoF[o * 32 + k], go[o], s);
              gc[k] = s;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4 *>(stile + lane * 32 + ((4 * q) ^ ((lane & 3) << 3))) = make_float4(gc[4 * q], gc[4 * q + 1], gc[4 * q + 2], gc[4 * q + 3]);
            __syncwarp();
            ENS_DBG(9, 1);
            float gpn[3];
            gather_bwd_warp<32>(a.sc.grid[LEVEL], ggrid, a.sc.dims[LEVEL], vox, valid, stile, want_rays, gpn);
            ENS_DBG(9, 2);
            if (want_rays) {
#pragma unroll
              for (int k = 0; k < 3; ++k) gp[k] += __ddiv_rn((double)gpn[k] * 2.0, __dsub_rn(a.sc.hi[k], a.sc.lo[k]));
            }
            __syncwarp();
          }
        }
        if (want_rays && valid) {                     // plane 2 DEC: the trilinear share of d L / d p (H adds the Fourier share)
          float *dst = a.gp + ((int64_t)(2 * DEC) * a.P + pt) * 3;
          dst[0] = (float)gp[0]; dst[1] = (float)gp[1]; dst[2] = (float)gp[2];
        }
        tmem_st_done();
        tc_fence_before();
      }
      ENS_DBG(8, 3);
    } else {
      // ================================ H warps ================================
      float c[32], rn[32];
      if (TAIL) load_row32(rbase + 4 * 4096, rn);                 // r_4: in flight during the gather
      // (an L2 prefetch of the next tile's saved rows from here, a whole tile ahead, was measured: no gain at 1000 rays, -2 % at 16 k)
      const Vox vc = make_vox(pn, a.sc.dims[CLEVEL]);
      if (!first) { mbar_wait(&barW, pw); pw ^= 1; }              // the previous tile's last pass: the M-side blocks are free
      // features straight into slot B: gather_warp's tile layout IS the MN-major swizzle (32-byte chunk ^ (row & 3))
      float *tileB = sMB + 4096 + w4 * 1024;
      __syncwarp();
      gather_warp<32>(a.sc.grid[CLEVEL], a.sc.dims[CLEVEL], vc, tileB, 0);
      if (TAIL) {
        // output layer sums: sum g_out[o] r_4[k], sum g_out[o]
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = go[o] * rn[k];
          dwor[o] += warp_colsum32(t);
          float sg = go[o];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, off);
          dbo[o] += sg;
        }
        load_row32(rbase + 3 * 4096, rn);                         // r_3: slot A of pass 0 (in flight during the sums below)
      }
      {
        float *lrow = sMB + 3 * 4096 + w4 * 1024 + lane * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int off = (4 * q) ^ ((lane & 3) << 3);
          const float4 x = *reinterpret_cast<const float4 *>(tileB + lane * 32 + off);
          c[4 * q] = x.x; c[4 * q + 1] = x.y; c[4 * q + 2] = x.z; c[4 * q + 3] = x.w;
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          *reinterpret_cast<float4 *>(lrow + off) = l;
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = go[o] * c[k];
          qo[o] += warp_colsum32(t);
        }
      }
      const float *B = sw + PB::off_B();
      bool waited = true;       // the M-side blocks are free (the wait above)
      ENS_DBG(8, 1);
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        const int i = pass_layer(ps);
        ENS_DBG(ps, 0);
        if (TAIL) {
          // slot A <- r_{i-1} (blocks 4..1) or a Fourier chunk; slot B <- a Fourier chunk or the features again
          //   pass:    0   1   2      3      4   5   6    7
          //   slot A:  r3  r2  e0     e2*    r1  r0  e0   e1        (* = computed and staged by the E warps, idle in those passes)
          //   slot B:  .   .   e1*    c      .   .   .    e2*
          const int ja = (ps == 2 || ps == 6) ? 0 : (ps == 7 ? 1 : -1);
          float va[32];
          if (ja >= 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              va[k] = fast_sin(fmaf(p32[2], B[2 * EMBP + 32 * ja + k], fmaf(p32[1], B[EMBP + 32 * ja + k], p32[0] * B[32 * ja + k])));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) va[k] = rn[k];
          }
          // the next main pass's slot A (r_{i-2}) is fetched while this pass's MMAs run
          const int nxt = (ps + 1 < NPASS) ? pass_layer(ps + 1) : -1;
          if (nxt >= 1) load_row32(rbase + (nxt - 1) * 4096, rn);
          ENS_DBG(ps, 1);
          if (!waited) { mbar_wait(&barW, pw); pw ^= 1; }
          ENS_DBG(ps, 2);
          if (ps == 3) stage_row(sMB + 4096, sMB + 3 * 4096, pl, c);
          else stage_row(sMB, sMB + 2 * 4096, pl, va);
          ENS_DBG(ps, 3);
        } else {
          if (!waited) { mbar_wait(&barW, pw); pw ^= 1; }
        }
        waited = false;
        first = false;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        cta_sync288();
        ENS_DBG(ps, 4);
        ENS_DBG(ps, 5);
      }
      if (TAIL && !SPLIT) {
        // ---- Fourier backward (the E warps do the trilinear backward meanwhile): g_q = g_e cos(p B), d L / d p += B g_q, dB ----
        mbar_wait(&barW, pw); pw ^= 1;                // last pass done (so is every data-gradient GEMM issued before it):
        first = true;                                 // slot B doubles as this warp's 32x32 tile; that phase is consumed
        tc_fence_after();
        ENS_DBG(9, 0);
        float *stile = sMB + 4096 + w4 * 1024;
        sP[pl] = make_float4(p32[0], p32[1], p32[2], 0.f);
        {
          float gpe[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
          for (int jc = 0; jc < 3; ++jc) {
            float ge[32];
            tmem_ld32(tb + TB_DE + 32 * jc, ge);
            tmem_zero32(tb + TB_DE + 32 * jc);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float bx = B[32 * jc + k], by = B[EMBP + 32 * jc + k], bz = B[2 * EMBP + 32 * jc + k];
              float sq, cq;
              fast_sincos(fmaf(p32[2], bz, fmaf(p32[1], by, p32[0] * bx)), sq, cq);
              const float gq = ge[k] * cq;
              ge[k] = gq;
              gpe[0] = fmaf(bx, gq, gpe[0]); gpe[1] = fmaf(by, gq, gpe[1]); gpe[2] = fmaf(bz, gq, gpe[2]);
            }
            // dB[r][32 jc + k] = sum_pt p[pt][r] g_q[pt][k]: transpose through the warp's tile, lane k walks its column
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4 *>(stile + lane * 32 + ((4 * q) ^ ((lane & 3) << 3))) = make_float4(ge[4 * q], ge[4 * q + 1], ge[4 * q + 2], ge[4 * q + 3]);
            __syncwarp();
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const float v = stile[r * 32 + (lane ^ ((r & 3) << 3))];
              const float4 pr = sP[32 * w4 + r];
              s0 = fmaf(pr.x, v, s0); s1 = fmaf(pr.y, v, s1); s2 = fmaf(pr.z, v, s2);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) if (q == jc) { dBacc[q][0] += s0; dBacc[q][1] += s1; dBacc[q][2] += s2; }
            __syncwarp();
          }
          if (want_rays && valid) {                   // plane 2 DEC + 1
            float *dst = a.gp + ((int64_t)(2 * DEC + 1) * a.P + pt) * 3;
            dst[0] = gpe[0]; dst[1] = gpe[1]; dst[2] = gpe[2];
          }
        }
        tmem_st_done();
        ENS_DBG(9, 1);
      }
    }
#undef ENS_DBG
  }

  // ---- flush the CTA's sums ----
  if (!first && !isI) { mbar_wait(&barW, pw); pw ^= 1; }
  tc_fence_after();
  float *raw = a.raw_acc + (int64_t)ROLE * RAW_FLOATS;
  if (isI) {
  } else if (isE) {
#pragma unroll 1
    for (int acc = 0; acc < NPASS; ++acc) {
      float v[32];
      tmem_ld32(tb + TB_ACC + 32 * acc, v);
      float *dst = raw + acc * 2048 + (pl & 63) * 32;               // lanes 64..127 hold the remainder products of rows 0..63
#pragma unroll
      for (int q = 0; q < 8; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    if (!SPLIT) {
#pragma unroll
      for (int i = 0; i < 5; ++i) atomicAdd(raw + RAW_BHAT + i * 32 + lane, bhat[i]);
    }
  } else {
    if (TAIL && !SPLIT) {
#pragma unroll
      for (int jc = 0; jc < 3; ++jc)
#pragma unroll
        for (int r = 0; r < 3; ++r) atomicAdd(raw + RAW_DB + r * 96 + 32 * jc + lane, dBacc[jc][r]);
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      atomicAdd(raw + RAW_DWOR + o * 32 + lane, dwor[o]);
      atomicAdd(raw + RAW_QO + o * 32 + lane, qo[o]);
      if (lane == 0) atomicAdd(raw + RAW_DBO + o, dbo[o]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(512u) : "memory");
}

// =============================================================================================================
// Weight-gradient kernel of the SPLIT backward (ENS_BWD_TC_SPLIT=1): the eight passes of bwd_tc_wg_kernel with the g_u rows read
// from the data-gradient kernel's buffer -- no data GEMMs, so no weights in shared memory, and the room goes into a SECOND
// staging stage: while the 32 MMAs of pass p read stage p % 2, the producers stage pass p + 1 into the other one.
//   E warps (0..3): the N side (g_u row of the pass's block, value + remainder; the next row is prefetched) and the Fourier chunk
//                   of the e-passes;  H warps (4..7): the M side (r rows, features, Fourier chunks) and the output-layer sums;
//   warp 8 issues.  barW[s] (tcgen05.commit) says stage s has been read; a producer waits on it before it writes the stage again.
// Slot B of BOTH stages holds the features (staged in passes 0 and 1); the e-passes that overwrite it (2 and 7) are followed by
// a restore (pass 4) or by the next tile's passes 0 / 1.
template <int ROLE>
__device__ __forceinline__ void wgrad_tc_body(const BwdTcArgs &a, float *smem_raw) {
  constexpr int LEVEL = (ROLE == ROLE_MIDDLE) ? ENS_LEVEL_MIDDLE : (ROLE == ROLE_COLOR ? ENS_LEVEL_COLOR : ENS_LEVEL_FINE);
  constexpr int CD = (LEVEL == ENS_LEVEL_FINE) ? 64 : 32;
  constexpr int NO = (ROLE == ROLE_COLOR) ? 3 : 1;
  constexpr int DEC = (ROLE == ROLE_MIDDLE) ? 0 : (ROLE == ROLE_COLOR ? 2 : 1);
  constexpr bool TAIL = ROLE != ROLE_FINE_CM;
  constexpr int CLEVEL = (ROLE == ROLE_FINE_CM) ? ENS_LEVEL_MIDDLE : LEVEL;
  constexpr int NPASS = TAIL ? 8 : 5;
  using PB = MlpPackTCB;

  float *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;
  // stage s: M side [hiA | hiB | loA | loB] (4 x 4096 floats), N side [g_hi | g_lo] (2 x 4096)
  float *sB = base + 2 * 24576;      // Fourier matrix [3][EMBP]
  __shared__ __align__(8) uint64_t barW[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool isE = warp < 4, isI = warp == 8;
  const int pl = tid & 127, w4 = warp & 3;

  {
    const float *gB = a.sc.w[LEVEL] + off_tcb<CD>() + PB::off_B();
    for (int i = tid; i < 3 * EMBP; i += 288) sB[i] = gB[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&barW[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&barW[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb0 = tmem_base_s;
  const uint32_t tb = tb0 + ((uint32_t)(32 * w4) << 16);
  if (isE) {
#pragma unroll 1
    for (int c = 0; c < 32 * NPASS; c += 32) { tmem_zero32(tb + c); tmem_zero32(tb + 256 + c); }
    tmem_st_done();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  uint32_t pwv[2] = {0u, 0u};        // phase parities of barW[s] as this thread has consumed them
  bool pend[2] = {false, false};     // a pass that read stage s has been issued and this thread has not waited for it yet
  int pcount = 0;                    // passes so far (stage = pcount & 1)
  float dwor[NO], qo[NO], dbo[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) dwor[o] = qo[o] = dbo[o] = 0.f;
  const int nctas = a.ctas[ROLE];
  // block whose g_u row is the N side of a pass; e-passes (2, 3, 7) reuse the row of the block pass before them
  auto pass_block = [](int p) { return TAIL ? (p == 0 ? 4 : (p <= 3 ? 3 : (p == 4 ? 2 : (p == 5 ? 1 : 0)))) : 4 - p; };
  auto is_epass = [](int p) { return TAIL && (p == 2 || p == 3 || p == 7); };
  auto wait_stage = [&](int st) {
    if (pend[st]) { mbar_wait(&barW[st], pwv[st]); pwv[st] ^= 1u; pend[st] = false; }
  };

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += nctas) {
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tile == (int64_t)nctas && (tid == 0 || tid == 128 || tid == 256);
#define ENS_DBG(ps, slot) do { if (dbg_on) a.dbg[((ROLE * 10 + (ps)) * 2 + ((tid >> 7) & 1)) * 8 + (slot)] = clock64(); } while (0)
#define ENS_DBGI(who, slot) do { if (dbg_on) a.dbg[((ROLE * 10 + 9) * 2 + (who)) * 8 + (slot)] = clock64(); } while (0)
    ENS_DBG(8, 0);
    const int64_t pt = tile * 128 + pl;
    const bool valid = pt < a.P;
    if (isI) {
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        const int st = pcount & 1;
        tc_fence_before();
        cta_sync288();
        tc_fence_after();
        ENS_DBGI(0, ps);
        if (a.exp_flags & 4) { }
        else if (a.exp_flags & 2) issue_wgrad_half(tb0 + 32 * ps, smem_u32(base + st * 24576), smem_u32(base + st * 24576 + 16384));
        else issue_wgrad_loop2(tb0 + 32 * ps, tb0 + 256 + 32 * ps, smem_u32(base + st * 24576), smem_u32(base + st * 24576 + 16384));
        umma_commit(&barW[st]);
        __syncwarp();
        ENS_DBGI(1, ps);
        ++pcount;
      }
      continue;
    }
    double p[3] = {0.0, 0.0, 0.0};
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const double *pp = a.pts + pt * 3;
      p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2];
      g4 = a.gout[pt];
    }
    float pn[3], p32[3];
    normalize64(p, a.sc.lo, a.sc.hi, pn);
#pragma unroll
    for (int k = 0; k < 3; ++k) p32[k] = __double2float_rn(p[k]);
    float go[NO];
    if (ROLE == ROLE_COLOR) { go[0] = g4.x; if (NO > 1) { go[1 % NO] = g4.y; go[2 % NO] = g4.z; } }
    else go[0] = g4.w;

    if (isE) {
      // ================================ E warps: N side + one Fourier chunk per e-pass ================================
      const float *gubase = a.gu_buf + (((int64_t)DEC * a.n_tiles + tile) * 5) * 4096 + pl * 32;
      float g[32], gn[32];
      if (!(a.exp_flags & 16)) load_row32(gubase + pass_block(0) * 4096, g);
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        const int st = pcount & 1;
        float *sM = base + st * 24576, *sN = sM + 16384;
        if (ps > 0 && pass_block(ps) != pass_block(ps - 1)) {
#pragma unroll
          for (int k = 0; k < 32; ++k) g[k] = gn[k];
        }
        ENS_DBG(ps, 0);
        float ve[32];
        if (is_epass(ps) && !(a.exp_flags & 8)) {
          const int je = (ps == 2) ? 1 : 2;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            ve[k] = fast_sin(fmaf(p32[2], sB[2 * EMBP + 32 * je + k], fmaf(p32[1], sB[EMBP + 32 * je + k], p32[0] * sB[32 * je + k])));
        }
        ENS_DBG(ps, 1);
        wait_stage(st);
        ENS_DBG(ps, 2);
        const bool do_stage = !(a.exp_flags & 1) || tile == (int64_t)blockIdx.x;
        if (do_stage) stage_row(sN, sN + 4096, pl, g);
        if (is_epass(ps) && do_stage) {
          if (ps == 3) stage_row(sM, sM + 2 * 4096, pl, ve);                 // e2 -> slot A
          else stage_row(sM + 4096, sM + 3 * 4096, pl, ve);                  // e1 / e2 -> slot B
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the next block's row, in flight during the barrier and the passes that reuse this one
        if (ps + 1 < NPASS && pass_block(ps + 1) != pass_block(ps) && !(a.exp_flags & 16)) load_row32(gubase + pass_block(ps + 1) * 4096, gn);
        pend[st] = true;
        ENS_DBG(ps, 3);
        tc_fence_before();
        cta_sync288();
        ENS_DBG(ps, 4);
        ++pcount;
      }
    } else {
      // ================================ H warps: M side + output-layer sums ================================
      const float *rbase = a.save_r + (((int64_t)DEC * a.n_tiles + tile) * 5) * 4096 + pl * 32;
      float c[32], rn[32];
      const bool skip_pro = (a.exp_flags & 32) != 0;
      if (TAIL && !(a.exp_flags & 16)) load_row32(rbase + 4 * 4096, rn);                 // r_4: in flight during the gather
      const Vox vc = make_vox(pn, a.sc.dims[CLEVEL]);
      const int st0 = pcount & 1;
      float *sM0 = base + st0 * 24576;
      wait_stage(st0);                                            // pass 0's stage: its slot B takes the gather directly
      float *tileB = sM0 + 4096 + w4 * 1024;
      __syncwarp();
      if (!skip_pro) gather_warp<32>(a.sc.grid[CLEVEL], a.sc.dims[CLEVEL], vc, tileB, 0);
      if (TAIL && !skip_pro) {
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = go[o] * rn[k];
          dwor[o] += warp_colsum32(t);
          float sg = go[o];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, off);
          dbo[o] += sg;
        }
        if (!(a.exp_flags & 16)) load_row32(rbase + 3 * 4096, rn);                         // r_3: slot A of pass 0
      }
      if (!skip_pro) {
        float *lrow = sM0 + 3 * 4096 + w4 * 1024 + lane * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int off = (4 * q) ^ ((lane & 3) << 3);
          const float4 x = *reinterpret_cast<const float4 *>(tileB + lane * 32 + off);
          c[4 * q] = x.x; c[4 * q + 1] = x.y; c[4 * q + 2] = x.z; c[4 * q + 3] = x.w;
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          *reinterpret_cast<float4 *>(lrow + off) = l;
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          float t[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = go[o] * c[k];
          qo[o] += warp_colsum32(t);
        }
      }
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        const int st = pcount & 1;
        float *sM = base + st * 24576;
        const int i = TAIL ? (ps == 0 ? 4 : (ps == 1 ? 3 : (ps == 4 ? 2 : (ps == 5 ? 1 : (ps == 6 ? 0 : -1))))) : 4 - ps;
        ENS_DBG(ps, 0);
        float va[32];
        if (TAIL) {
          //   pass:    0   1   2      3      4   5   6    7
          //   slot A:  r3  r2  e0     e2*    r1  r0  e0   e1        (* = staged by the E warps)
          //   slot B:  c   c   e1*    .      c   .   .    e2*       (c: both stages hold the features)
          const int ja = (ps == 2 || ps == 6) ? 0 : (ps == 7 ? 1 : -1);
          if (ja >= 0 && !(a.exp_flags & 8)) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              va[k] = fast_sin(fmaf(p32[2], sB[2 * EMBP + 32 * ja + k], fmaf(p32[1], sB[EMBP + 32 * ja + k], p32[0] * sB[32 * ja + k])));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) va[k] = rn[k];
          }
          // the next block pass's slot A (r_{i-2}) is fetched while this pass is staged
          int nxt = -1;
          if (ps + 1 < NPASS) nxt = (ps + 1 == 1) ? 3 : (ps + 1 == 4 ? 2 : (ps + 1 == 5 ? 1 : -1));
          if (nxt >= 1 && !(a.exp_flags & 16)) load_row32(rbase + (nxt - 1) * 4096, rn);
        }
        ENS_DBG(ps, 1);
        wait_stage(st);
        ENS_DBG(ps, 2);
        const bool do_stage = !(a.exp_flags & 1) || tile == (int64_t)blockIdx.x;
        if (do_stage && TAIL && ps != 3) stage_row(sM, sM + 2 * 4096, pl, va);
        if (do_stage && (ps == 1 || (TAIL && ps == 4))) stage_row(sM + 4096, sM + 3 * 4096, pl, c);
        (void)i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        pend[st] = true;
        ENS_DBG(ps, 3);
        tc_fence_before();
        cta_sync288();
        ENS_DBG(ps, 4);
        ++pcount;
      }
    }
#undef ENS_DBG
#undef ENS_DBGI
  }

  // ---- flush the CTA's sums ----
  if (!isI) { wait_stage(0); wait_stage(1); }
  tc_fence_after();
  float *raw = a.raw_acc + (int64_t)ROLE * RAW_FLOATS;
  if (isI) {
  } else if (isE) {
#pragma unroll 1
    for (int acc = 0; acc < NPASS; ++acc) {
      float v[32], v2[32];
      tmem_ld32(tb + 32 * acc, v);
      tmem_ld32(tb + 256 + 32 * acc, v2);
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] += v2[k];
      float *dst = raw + acc * 2048 + (pl & 63) * 32;               // lanes 64..127 hold the remainder products of rows 0..63
#pragma unroll
      for (int q = 0; q < 8; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  } else {
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      atomicAdd(raw + RAW_DWOR + o * 32 + lane, dwor[o]);
      atomicAdd(raw + RAW_QO + o * 32 + lane, qo[o]);
      if (lane == 0) atomicAdd(raw + RAW_DBO + o, dbo[o]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(512u) : "memory");
}

template <int STAGE>
__global__ void __launch_bounds__(288, 1) wgrad_tc_kernel(BwdTcArgs a) {
  extern __shared__ __align__(128) float smem[];
  const int role = blockIdx.y;
  if ((int)blockIdx.x >= a.ctas[role]) return;
  if (role == ROLE_MIDDLE) wgrad_tc_body<ROLE_MIDDLE>(a, smem);
  else if (role == ROLE_FINE) { if constexpr (STAGE >= ENS_STAGE_FINE) wgrad_tc_body<ROLE_FINE>(a, smem); }
  else if (role == ROLE_COLOR) { if constexpr (STAGE == ENS_STAGE_COLOR) wgrad_tc_body<ROLE_COLOR>(a, smem); }
  else { if constexpr (STAGE >= ENS_STAGE_FINE) wgrad_tc_body<ROLE_FINE_CM>(a, smem); }
}

template <int STAGE, bool SPLIT = false>
__global__ void __launch_bounds__(288, 1) bwd_tc_wg_kernel(BwdTcArgs a) {
  extern __shared__ __align__(128) float smem[];
  const int role = blockIdx.y;
  if ((int)blockIdx.x >= a.ctas[role]) return;
  if (role == ROLE_MIDDLE) bwd_tc_wg_body<ROLE_MIDDLE, SPLIT>(a, smem);
  else if (role == ROLE_FINE) { if constexpr (STAGE >= ENS_STAGE_FINE) bwd_tc_wg_body<ROLE_FINE, SPLIT>(a, smem); }
  else if (role == ROLE_COLOR) { if constexpr (STAGE == ENS_STAGE_COLOR) bwd_tc_wg_body<ROLE_COLOR, SPLIT>(a, smem); }
  else { if constexpr (STAGE >= ENS_STAGE_FINE) bwd_tc_wg_body<ROLE_FINE_CM, SPLIT>(a, smem); }
}

template <int STAGE, bool WG, bool SPLIT = false>
__global__ void __launch_bounds__(256, 1) bwd_tc_kernel(BwdTcArgs a) {
  extern __shared__ __align__(128) float smem[];
  const int role = blockIdx.y;
  if ((int)blockIdx.x >= a.ctas[role]) return;
  if (role == ROLE_MIDDLE) bwd_tc_body<ROLE_MIDDLE, WG, SPLIT>(a, smem);
  else if (role == ROLE_FINE) { if constexpr (STAGE >= ENS_STAGE_FINE) bwd_tc_body<ROLE_FINE, WG, SPLIT>(a, smem); }
  else if (role == ROLE_COLOR) { if constexpr (STAGE == ENS_STAGE_COLOR) bwd_tc_body<ROLE_COLOR, WG, SPLIT>(a, smem); }
}

// ---------------------------------------------------------------------------------------------
// raw2outputs_nerf_color backward (common.py:256-297), one warp per ray: the lanes fetch the samples and
// evaluate the sigmoids, lane 0 runs the sequential products / sums in the order of the fused kernels, the lanes write
// d L / d (decoder outputs) of their samples.  Out-of-bound samples (Renderer.py:58 overwrote their occupancy) get none.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) composite_bwd_kernel(DevScene sc, const float4 *__restrict__ raw, const double *__restrict__ z,
                                                            const double *__restrict__ pts, int64_t R, int S,
                                                            const double *__restrict__ g_depth, const double *__restrict__ g_var,
                                                            const float *__restrict__ g_color, float4 *__restrict__ gout) {
  __shared__ float4 s_raw[4][ENS_MAX_SAMPLES];
  __shared__ double s_z[4][ENS_MAX_SAMPLES], s_gw[4][ENS_MAX_SAMPLES];
  __shared__ float s_a[4][ENS_MAX_SAMPLES], s_T[4][ENS_MAX_SAMPLES], s_w[4][ENS_MAX_SAMPLES], s_suf[4][ENS_MAX_SAMPLES];
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + wi;
  if (ray >= R) return;
  for (int k = lane; k < S; k += 32) {
    const int64_t pi = ray * S + k;
    const float4 r4 = raw[pi];
    s_raw[wi][k] = r4;
    s_z[wi][k] = z[pi];
    s_a[wi][k] = 1.f / (1.f + expf(-(10.f * r4.w)));
  }
  __syncwarp();
  float gcl[3] = {0.f, 0.f, 0.f};
  if (g_color) { gcl[0] = g_color[ray * 3]; gcl[1] = g_color[ray * 3 + 1]; gcl[2] = g_color[ray * 3 + 2]; }
  if (S <= 64) {
    // Lane-parallel form: lane l owns samples l and l + 32.  The transmittance is an exclusive prefix product, the term behind
    // every alpha an exclusive suffix sum: warp scans instead of five serial 48-step loops of one lane (the kernel was 13 us for
    // 200 rays and for 1000: pure latency).  Scan order differs from the forward's sequential order by float rounding only.
    const int k0 = lane, k1 = lane + 32;
    const bool h0 = k0 < S, h1 = k1 < S;
    const float a0 = h0 ? s_a[wi][k0] : 0.f, a1 = h1 ? s_a[wi][k1] : 0.f;
    const float om0 = h0 ? __fadd_rn(__fsub_rn(1.f, a0), 1e-10f) : 1.f, om1 = h1 ? __fadd_rn(__fsub_rn(1.f, a1), 1e-10f) : 1.f;
    float p0 = om0, p1 = om1;                               // inclusive prefix products within each block of 32
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float q0 = __shfl_up_sync(0xffffffffu, p0, o), q1 = __shfl_up_sync(0xffffffffu, p1, o);
      if (lane >= o) { p0 *= q0; p1 *= q1; }
    }
    const float tot0 = __shfl_sync(0xffffffffu, p0, 31);
    float T0 = __shfl_up_sync(0xffffffffu, p0, 1), T1 = __shfl_up_sync(0xffffffffu, p1, 1);
    if (lane == 0) { T0 = 1.f; T1 = 1.f; }
    T1 *= tot0;
    const float w0 = a0 * T0, w1 = a1 * T1;
    const double z0 = h0 ? s_z[wi][k0] : 0.0, z1 = h1 ? s_z[wi][k1] : 0.0;
    double dep = (double)w0 * z0 + (double)w1 * z1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dep += __shfl_xor_sync(0xffffffffu, dep, o);
    double wdz = (double)w0 * (z0 - dep) + (double)w1 * (z1 - dep);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wdz += __shfl_xor_sync(0xffffffffu, wdz, o);
    const double gd = g_depth ? g_depth[ray] : 0.0;
    const double gv = g_var ? g_var[ray] : 0.0;
    const double gdt = gd + gv * (-2.0 * wdz);
    double gw0 = 0.0, gw1 = 0.0;
    if (h0) { const float4 rk = s_raw[wi][k0]; const double dz = z0 - dep; gw0 = (double)rk.x * gcl[0] + (double)rk.y * gcl[1] + (double)rk.z * gcl[2] + gdt * z0 + gv * dz * dz; }
    if (h1) { const float4 rk = s_raw[wi][k1]; const double dz = z1 - dep; gw1 = (double)rk.x * gcl[0] + (double)rk.y * gcl[1] + (double)rk.z * gcl[2] + gdt * z1 + gv * dz * dz; }
    // exclusive suffix sums of w_k gw_k
    const float t0 = w0 * (float)gw0, t1 = w1 * (float)gw1;
    float q0 = t0, q1 = t1;                                 // inclusive suffix sums within each block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float r0 = __shfl_down_sync(0xffffffffu, q0, o), r1 = __shfl_down_sync(0xffffffffu, q1, o);
      if (lane + o < 32) { q0 += r0; q1 += r1; }
    }
    const float tot1 = __shfl_sync(0xffffffffu, q1, 0);     // everything in the second block
    const float suf0 = (q0 - t0) + tot1, suf1 = q1 - t1;
    if (h0) { s_T[wi][k0] = T0; s_w[wi][k0] = w0; s_gw[wi][k0] = gw0; s_suf[wi][k0] = suf0; }
    if (h1) { s_T[wi][k1] = T1; s_w[wi][k1] = w1; s_gw[wi][k1] = gw1; s_suf[wi][k1] = suf1; }
  } else if (lane == 0) {
    float T = 1.f;
    for (int k = 0; k < S; ++k) {
      s_T[wi][k] = T;
      s_w[wi][k] = __fmul_rn(s_a[wi][k], T);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, s_a[wi][k]), 1e-10f));
    }
    double dep = 0.0;
    for (int k = 0; k < S; ++k) dep += (double)s_w[wi][k] * s_z[wi][k];
    double wdz = 0.0;
    for (int k = 0; k < S; ++k) wdz += (double)s_w[wi][k] * (s_z[wi][k] - dep);
    const double gd = g_depth ? g_depth[ray] : 0.0;
    const double gv = g_var ? g_var[ray] : 0.0;
    const double gdt = gd + gv * (-2.0 * wdz);
    for (int k = 0; k < S; ++k) {
      const float4 rk = s_raw[wi][k];
      const double zk = s_z[wi][k], dz = zk - dep;
      s_gw[wi][k] = (double)rk.x * gcl[0] + (double)rk.y * gcl[1] + (double)rk.z * gcl[2] + gdt * zk + gv * dz * dz;
    }
    float accs = 0.f;
    for (int k = S - 1; k >= 0; --k) {
      s_suf[wi][k] = accs;
      accs += s_w[wi][k] * (float)s_gw[wi][k];
    }
  }
  __syncwarp();
  for (int k = lane; k < S; k += 32) {
    const int64_t pi = ray * S + k;
    bool inside = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) { const double pk = pts[pi * 3 + q]; inside &= (pk < sc.hi[q]) && (pk > sc.lo[q]); }
    const float alpha = s_a[wi][k];
    const float gw = (float)s_gw[wi][k];
    const float om = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    const float g_alpha = s_T[wi][k] * gw - s_suf[wi][k] / om;
    const float wv = s_w[wi][k];
    gout[pi] = make_float4(wv * gcl[0], wv * gcl[1], wv * gcl[2], inside ? 10.f * alpha * (1.f - alpha) * g_alpha : 0.f);
  }
}

// g_rays_o = sum_s g_p, g_rays_d = sum_s z_s g_p over the role planes (float64 sums, as the fused kernels); a warp per ray
__global__ void __launch_bounds__(128) rays_reduce_kernel(const float *__restrict__ gp, int n_planes, const double *__restrict__ z,
                                                          int64_t R, int S, float *__restrict__ g_rays_o, float *__restrict__ g_rays_d) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (ray >= R) return;
  const int64_t P = R * (int64_t)S;
  double so[3] = {0.0, 0.0, 0.0}, sd[3] = {0.0, 0.0, 0.0};
  for (int k = lane; k < S; k += 32) {
    const int64_t pi = ray * S + k;
    const double zk = z[pi];
    double gk[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < n_planes; ++q) {
      const float *src = gp + ((int64_t)q * P + pi) * 3;
      gk[0] += (double)src[0]; gk[1] += (double)src[1]; gk[2] += (double)src[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { so[c] += gk[c]; sd[c] += gk[c] * zk; }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      so[c] += __shfl_xor_sync(0xffffffffu, so[c], off);
      sd[c] += __shfl_xor_sync(0xffffffffu, sd[c], off);
    }
  if (lane < 3) {
    const double vo = lane == 0 ? so[0] : (lane == 1 ? so[1] : so[2]);
    const double vd = lane == 0 ? sd[0] : (lane == 1 ? sd[1] : sd[2]);
    if (g_rays_o) g_rays_o[ray * 3 + lane] = (float)vo;
    if (g_rays_d) g_rays_d[ray * 3 + lane] = (float)vd;
  }
}

// ---------------------------------------------------------------------------------------------
// folded sums -> the reference's parameter gradients (see the header of this file).  One CTA per decoder.
// blob: the decoder's packed weights (fma section MlpPack<CD>: W_iT [K_i][32], b_i, Wc_iT [CD][32], bc_i, WoT [32][4], bo).
// ---------------------------------------------------------------------------------------------
// NO = output rows that carry gradient, NOF = rows of output_linear (the colour decoder's 4th row gets none, decoder.py:341)
// Every CTA stages what the decoder's unfolding reads -- the raw sums (69 KB; FINE_CM's too for the fine decoder) and the fma
// section of the packed weights -- in shared memory (coalesced), then its threads take output elements round-robin.
template <int CD, int NO, int NOF>
__device__ __forceinline__ void unfold_body(const float *__restrict__ graw, const float *__restrict__ graw_cm,
                                            const float *__restrict__ gblob, float *__restrict__ gdec, float *__restrict__ smem) {
  using PK = MlpPack<CD>;
  using GO = MlpGrad<CD, NOF>;
  float *raw = smem, *raw_cm = smem + RAW_FLOATS, *blob = raw_cm + (CD == 64 ? RAW_FLOATS : 0);
  {
    // cp.async: every 16-byte copy of the CTA in flight at once
    auto copy16 = [&](float *dst, const float *src, int n4) {
      const uint32_t s0 = smem_u32(dst);
      for (int i = threadIdx.x; i < n4; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + i * 16), "l"(src + i * 4) : "memory");
    };
    copy16(raw, graw, RAW_FLOATS / 4);
    if (CD == 64) copy16(raw_cm, graw_cm, RAW_FLOATS / 4);
    static_assert(PK::total() % 4 == 0, "blob section must be a whole number of 16-byte copies");
    copy16(blob, gblob, PK::total() / 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const int tid = blockIdx.y * blockDim.x + threadIdx.x, nt = gridDim.y * blockDim.x;
  auto accS = [](int i) { return i == 4 ? 0 : (i == 3 ? 1 : (i == 2 ? 4 : 5)); };                  // i = 1..4
  auto S = [&](int i, int n, int k) { return raw[accS(i) * 2048 + k * 32 + n]; };
  auto Q = [&](int i, int n, int ch) {                                                             // i = 0..4
    if (ch >= 32) return raw_cm[(4 - i) * 2048 + ch * 32 + n];                                     // FINE_CM rows 32..63
    const int acc = (i == 0) ? 6 : accS(i);
    return raw[acc * 2048 + (32 + ch) * 32 + n];
  };
  auto bhat = [&](int i, int n) { return raw[RAW_BHAT + i * 32 + n]; };
  auto Qo = [&](int o, int ch) { return ch >= 32 ? raw_cm[RAW_QO + o * 32 + ch - 32] : raw[RAW_QO + o * 32 + ch]; };
  auto Wh = [&](int i, int n, int k) { return blob[PK::off_W(i) + ((i == 3 ? EMB : 0) + k) * 32 + n]; };   // hidden part of W_i, i = 1..4
  auto Wc = [&](int i, int k, int ch) { return blob[PK::off_Wc(i) + ch * 32 + k]; };                        // fc_c[i].weight[k][ch]
  auto bc = [&](int i, int k) { return blob[PK::off_bc(i) + k]; };
  auto Wo = [&](int o, int k) { return blob[PK::off_Wo() + k * 4 + o]; };

  // fc_c.i.weight [32][CD], fc_c.i.bias [32]
  for (int idx = tid; idx < 5 * 32 * CD; idx += nt) {
    const int i = idx / (32 * CD), k = (idx / CD) % 32, ch = idx % CD;
    float s = 0.f;
    // the lanes of a warp differ in ch = the ROW of the staged sums: start each lane's walk over n at its own column
    // (bank-conflict free; the order of a 32-term fp32 sum is free)
    if (i < 4) { for (int q = 0; q < 32; ++q) { const int n = (q + threadIdx.x) & 31; s = fmaf(Wh(i + 1, n, k), Q(i + 1, n, ch), s); } }
    else { for (int o = 0; o < NO; ++o) s = fmaf(Wo(o, k), Qo(o, ch), s); }
    gdec[GO::off_Wc(i) + k * CD + ch] += s;
  }
  for (int idx = tid; idx < 5 * 32; idx += nt) {
    const int i = idx / 32, k = idx % 32;
    float s = 0.f;
    if (i < 4) { for (int n = 0; n < 32; ++n) s = fmaf(Wh(i + 1, n, k), bhat(i + 1, n), s); }
    else { for (int o = 0; o < NO; ++o) s = fmaf(Wo(o, k), raw[RAW_DBO + o], s); }
    gdec[GO::off_bc(i) + k] += s;
    gdec[GO::off_b(i) + k] += bhat(i, k);                                                   // pts_linears.i.bias
  }
  // embedder._B [3][93]
  for (int idx = tid; idx < 3 * EMB; idx += nt) gdec[GO::off_B() + idx] += raw[RAW_DB + (idx / EMB) * 96 + idx % EMB];
  // pts_linears.0.weight [32][93] and the embedding half of pts_linears.3.weight [32][125]
  for (int idx = tid; idx < 32 * EMB; idx += nt) {
    const int n = idx / EMB, k = idx % EMB;
    const float e0 = (k < 32) ? raw[6 * 2048 + k * 32 + n] : raw[7 * 2048 + (k - 32) * 32 + n];
    const float e3 = (k < 64) ? raw[2 * 2048 + k * 32 + n] : raw[3 * 2048 + (k - 64) * 32 + n];
    gdec[GO::off_W(0) + n * EMB + k] += e0;
    gdec[GO::off_W(3) + n * 125 + k] += e3;
  }
  // hidden parts of pts_linears.1..4.weight:  S_i + Q_i Wc_{i-1}^T + b^_i bc_{i-1}^T
  for (int idx = tid; idx < 4 * 1024; idx += nt) {
    const int i = 1 + idx / 1024, n = (idx / 32) % 32, k = idx % 32;
    float s = fmaf(bhat(i, n), bc(i - 1, k), S(i, n, k));
    for (int ch = 0; ch < CD; ++ch) s = fmaf(Q(i, n, ch), Wc(i - 1, k, ch), s);
    const int K = (i == 3) ? 125 : 32;
    gdec[GO::off_W(i) + n * K + (i == 3 ? EMB : 0) + k] += s;
  }
  // output_linear.weight [NOF][32], bias  (NO rows carry gradient)
  for (int idx = tid; idx < NO * 32; idx += nt) {
    const int o = idx / 32, k = idx % 32;
    float s = fmaf(raw[RAW_DBO + o], bc(4, k), raw[RAW_DWOR + o * 32 + k]);
    for (int ch = 0; ch < CD; ++ch) s = fmaf(Qo(o, ch), Wc(4, k, ch), s);
    gdec[GO::off_Wo() + o * 32 + k] += s;
  }
  for (int o = tid; o < NO; o += nt) gdec[GO::off_bo() + o] += raw[RAW_DBO + o];
}

struct UnfoldArgs {
  const float *raw;        // [4 roles][RAW_FLOATS]
  const float *w[4];       // packed blobs
  float *gdec[4];
};

constexpr int UNFOLD_SMEM_FLOATS = 2 * RAW_FLOATS + MlpPack<64>::total();

__global__ void __launch_bounds__(512) unfold_kernel(UnfoldArgs a, int stage) {
  extern __shared__ __align__(16) float usm[];
  const int d = blockIdx.x;
  if (d == 0) unfold_body<32, 1, 1>(a.raw + ROLE_MIDDLE * RAW_FLOATS, nullptr, a.w[ENS_LEVEL_MIDDLE], a.gdec[ENS_LEVEL_MIDDLE], usm);
  else if (d == 1) {
    if (stage >= ENS_STAGE_FINE)
      unfold_body<64, 1, 1>(a.raw + ROLE_FINE * RAW_FLOATS, a.raw + ROLE_FINE_CM * RAW_FLOATS, a.w[ENS_LEVEL_FINE], a.gdec[ENS_LEVEL_FINE], usm);
  } else if (stage == ENS_STAGE_COLOR) {
    unfold_body<32, 3, 4>(a.raw + ROLE_COLOR * RAW_FLOATS, nullptr, a.w[ENS_LEVEL_COLOR], a.gdec[ENS_LEVEL_COLOR], usm);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT_MMA) place_bwd_kernel(DevScene sc, RayArgs ra, double *__restrict__ z_out,
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

// workspace: points [P][3] f64 | z [P] f64 | g_out [P][4] | g_p planes [3][P][3] | raw sums [4][RAW_FLOATS]
static bool use_split_tc() {
  const char *v = std::getenv("ENS_BWD_TC_SPLIT");          // 1: data-gradient kernel + weight-gradient kernel (A/B; measured slower)
  return v && v[0] == '1';
}

// ... | with decoder gradients (split backward): the g_u rows, [3 decoders][n_tiles][5][128][32] f32
int64_t tc_bwd_workspace_bytes(int64_t n_rays, int S, int stage, bool wg) {
  if (stage == ENS_STAGE_COARSE || n_rays <= 0 || S < 1) return 0;
  const int64_t P = n_rays * (int64_t)S;
  const int64_t n_tiles = (P + 127) / 128;
  return P * (24 + 8 + 16 + 72) + 4 * (int64_t)RAW_FLOATS * 4 + 256 + 256 + ((wg && use_split_tc()) ? 3 * n_tiles * 5 * 4096 * 4 : 0);
}


// MODE 0: pose-only chain; 1: fused data + weight gradients; 2: split -- data-gradient kernel; 3: split -- weight-gradient kernel
template <int STAGE, int MODE>
static int launch_bwd_tc(const BwdTcArgs &a, int nroles, int max_ctas, cudaStream_t s) {
  constexpr bool WGK = MODE == 1 || MODE == 3;
  const size_t smem = (size_t)((WGK ? 16384 + 8192 : 8192) + MlpPackTCB::total()) * 4 + 1024;
  if (MODE == 1) {
    ENS_CUDA_CALL(cudaFuncSetAttribute(bwd_tc_wg_kernel<STAGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_tc_wg_kernel<STAGE, false><<<dim3((unsigned)max_ctas, (unsigned)nroles), 288, smem, s>>>(a);
  } else if (MODE == 3) {
    const size_t smem3 = (size_t)(2 * 24576 + 3 * EMBP) * 4 + 1024;          // two staging stages + the Fourier matrix
    ENS_CUDA_CALL(cudaFuncSetAttribute(wgrad_tc_kernel<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    wgrad_tc_kernel<STAGE><<<dim3((unsigned)max_ctas, (unsigned)nroles), 288, smem3, s>>>(a);
  } else if (MODE == 2) {
    ENS_CUDA_CALL(cudaFuncSetAttribute(bwd_tc_kernel<STAGE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_tc_kernel<STAGE, false, true><<<dim3((unsigned)max_ctas, (unsigned)nroles), 256, smem, s>>>(a);
  } else {
    ENS_CUDA_CALL(cudaFuncSetAttribute(bwd_tc_kernel<STAGE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_tc_kernel<STAGE, false, false><<<dim3((unsigned)max_ctas, (unsigned)nroles), 256, smem, s>>>(a);
  }
  ENS_CHECK_CUDA();
  return ENS_OK;
}

int tc_render_bwd(BwdArgs &b, int stage, bool wg, void *workspace, int64_t workspace_bytes, cudaStream_t s) {
  if (stage == ENS_STAGE_COARSE) return ENS_EUNSUPPORTED;
  const int64_t R = b.ra.R;
  const int S = b.ra.S;
  const int64_t P = R * (int64_t)S;
  if (!workspace || workspace_bytes < tc_bwd_workspace_bytes(R, S, stage, wg)) return ENS_ESHAPE;
  if (b.save_masks == nullptr || (wg && b.save_r == nullptr)) return ENS_EINVAL;
  char *base = reinterpret_cast<char *>(workspace);
  double *pts = reinterpret_cast<double *>(base);
  double *z = reinterpret_cast<double *>(base + P * 24);
  float4 *gout = reinterpret_cast<float4 *>(base + P * 32);
  float *gp = reinterpret_cast<float *>(base + P * 48);
  float *raw = reinterpret_cast<float *>(base + ((P * 120 + 255) / 256) * 256);
  float *gu_buf = raw + 4 * RAW_FLOATS + 64;
  const bool want_rays = b.g_rays_o != nullptr || b.g_rays_d != nullptr;

  b.ra.rpc = NT_MMA / S;
  place_bwd_kernel<<<(unsigned)((R + b.ra.rpc - 1) / b.ra.rpc), NT_MMA, 0, s>>>(b.sc, b.ra, z, pts);
  ENS_CHECK_CUDA();
  composite_bwd_kernel<<<(unsigned)((R + 3) / 4), 128, 0, s>>>(b.sc, reinterpret_cast<const float4 *>(b.raw), z, pts, R, S,
                                                              b.g_depth, b.g_var, b.g_color, gout);
  ENS_CHECK_CUDA();
  if (wg) ENS_CUDA_CALL(cudaMemsetAsync(raw, 0, 4 * (size_t)RAW_FLOATS * 4, s));

  BwdTcArgs a;
  a.sc = b.sc; a.pts = pts; a.gout = gout; a.P = P; a.n_tiles = (P + 127) / 128;
  a.save_r = b.save_r; a.save_m = b.save_masks; a.m_stride = b.n_tiles * 160;
  for (int l = 0; l < 4; ++l) a.ggrid[l] = b.ggrid[l];
  a.raw_acc = raw; a.gp = want_rays ? gp : nullptr;
  a.gu_buf = wg ? gu_buf : nullptr;
  { const char *v = std::getenv("ENS_WGRAD_EXP"); a.exp_flags = v ? std::atoi(v) : 0; }
  {
    const char *v = std::getenv("ENS_BWD_TC_DBG");          // tools/time_passes.py: device address of a long long[640]
    a.dbg = v ? reinterpret_cast<long long *>(std::strtoull(v, nullptr, 0)) : nullptr;
  }
  const int sms = sm_count();
  const int ndec = stage == ENS_STAGE_MIDDLE ? 1 : (stage == ENS_STAGE_FINE ? 2 : 3);
  const bool split = wg && use_split_tc();
  // persistent CTAs per role, in proportion to the work of a tile (the FINE_CM role runs the hidden chain only)
  // (a greedy split that also evens out the round quantisation -- 42 / 47 / 44 / 15 CTAs for 375 tiles -- measured 4-5 %
  // SLOWER than this proportional one in the same session, at 1000 and at 16 384 rays)
  auto deal = [&](int nroles, bool two_tiles, int &max_ctas) {
    const double cost[4] = {1.0, 1.0, 1.0, 0.36};
    double tot = 0.0;
    for (int r = 0; r < 4; ++r) { a.ctas[r] = 0; if (r < ndec || (r == 3 && nroles == 4)) tot += cost[r]; }
    max_ctas = 0;
    for (int r = 0; r < 4; ++r) {
      if (!(r < ndec || (r == 3 && nroles == 4))) continue;
      int64_t n = (int64_t)(sms * cost[r] / tot);
      if (n < 1) n = 1;
      const int64_t cap = two_tiles ? (a.n_tiles + 1) / 2 : a.n_tiles;      // two tiles in flight per CTA
      if (n > cap) n = cap;
      a.ctas[r] = (int)n;
      if ((int)n > max_ctas) max_ctas = (int)n;
    }
  };
  int rc = ENS_OK, max_ctas = 0;
#define ENS_TCB(MODE, NR)                                                                 \
  (stage == ENS_STAGE_MIDDLE ? launch_bwd_tc<ENS_STAGE_MIDDLE, MODE>(a, NR, max_ctas, s)  \
   : stage == ENS_STAGE_FINE ? launch_bwd_tc<ENS_STAGE_FINE, MODE>(a, NR, max_ctas, s)    \
                             : launch_bwd_tc<ENS_STAGE_COLOR, MODE>(a, NR, max_ctas, s))
  if (!wg) {
    deal(ndec, true, max_ctas);
    rc = ENS_TCB(0, ndec);
  } else if (split) {
    // data gradients (two tile groups per CTA; writes the g_u rows, bias sums, dB, grid scatter, d L / d p) ...
    deal(ndec, true, max_ctas);
    rc = ENS_TCB(2, ndec);
    if (rc != ENS_OK) return rc;
    // ... then the weight-gradient GEMMs over the point dimension
    const int nroles = ndec > 1 ? 4 : ndec;
    deal(nroles, false, max_ctas);
    rc = ENS_TCB(3, nroles);
  } else {
    const int nroles = ndec > 1 ? 4 : ndec;
    deal(nroles, false, max_ctas);
    rc = ENS_TCB(1, nroles);
  }
#undef ENS_TCB
  if (rc != ENS_OK) return rc;
  if (wg) {
    UnfoldArgs u;
    u.raw = raw;
    for (int l = 0; l < 4; ++l) { u.w[l] = b.sc.w[l]; u.gdec[l] = b.gdec[l]; }
    ENS_CUDA_CALL(cudaFuncSetAttribute(unfold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UNFOLD_SMEM_FLOATS * 4));
    unfold_kernel<<<dim3(ndec, 32), 512, UNFOLD_SMEM_FLOATS * 4, s>>>(u, stage);   // 96 CTAs: the work is index arithmetic, not bytes
    ENS_CHECK_CUDA();
  }
  if (want_rays) {
    rays_reduce_kernel<<<(unsigned)((R + 3) / 4), 128, 0, s>>>(gp, (wg && !split) ? 2 * ndec : ndec, z, R, S, b.g_rays_o, b.g_rays_d);
    ENS_CHECK_CUDA();
  }
  return ENS_OK;
}

}  // namespace ens
