// Fused render BACKWARD on the tensor pipe (variant "mma"): recompute + data gradients + weight gradients
// of Renderer.render_batch_ray, stages middle / fine / color.
//
// Mapping (same as the forward, ens_render_mma.cu): a CTA owns whole rays; thread tid = sample point
// (ray tid / S, sample tid % S); a warp's 32 points are two m16 row tiles; lane (g = L/4, t = L%4) holds
// rows {g, g+8, g+16, g+24} x columns {8n+2t, 8n+2t+1} of every 32x32 activation / gradient tile.
//
// Per decoder:
//   1. forward recompute (3xTF32 mma.sync, activations chained in registers); keeps the five relu masks
//      (one bit per accumulator element) and, when decoder gradients are wanted, spills h_0..h_3 as
//      swizzled 32x32 tiles to an L2-resident scratch (they are the x operands of the weight-gradient GEMMs);
//   2. data gradients  g_x = g_u W  with the TRANSPOSED weight blob (MlpPackV2B): the gradient tile in
//      accumulator layout is again directly the A fragment of the next GEMM;  g_c += g_h Wc  accumulates the
//      feature gradient over the five blocks;
//   3. weight gradients  dW = sum_pt g_u^T x  are GEMMs over the POINT dimension.  Both operands are
//      accumulator-layout tiles with points on the fragment rows, so they are staged through shared memory
//      (swizzled, conflict-free both ways) and the CTA's warps split (row tile, third of the points) items
//      of the output; results leave with RED.ADD to the flat per-decoder gradient buffer;
//   4. trilinear backward: the feature gradient is transposed through the warp's own feature tile so that 8
//      lanes own one point's 128-byte voxel line: red.global.add.v4.f32 per corner, and the coordinate
//      gradient from the re-read corner values;
//   5. Fourier embedding backward in three 32-column chunks (sincos recomputed in accumulator layout).
//
// Two instantiations per stage: WG = decoder gradients on (mapping: 192 threads = 4 rays, three staging
// tiles) and WG = off (tracking / event render: 256 threads = 5 rays, no staging, no CTA barriers in the
// layer loop).
#include "ens_mma.cuh"

namespace ens {

// asynchronous 4 KB tile copy global -> shared by one warp (no registers)
__device__ __forceinline__ void cp_async_tile(float *__restrict__ sdst, const float *__restrict__ gsrc, int lane) {
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sdst);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int idx = (q * 32 + lane) * 4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + idx * 4), "l"(gsrc + idx) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// weight-gradient GEMM item:  d[nt](n, k) += sum_{pt in [p0,p1)} A[pt][16m + n] * B[pt][bcol0 + 8nt + k]
// A: swizzled [pts][LDA] tile (columns a0 ..), B: swizzled [pts][LDB] tile.  3xTF32.
// bsum[0..1] accumulate the column sums of A (rows 16m+g, 16m+g+8 of the output) = bias gradients.
// ---------------------------------------------------------------------------------------------
template <int LDA, int LDB, int NTILES>
__device__ __forceinline__ void wgrad_mma(float (&d)[NTILES][4], float (&bsum)[2], const float *__restrict__ A,
                                          int acol0, int m, const float *__restrict__ B, int bcol0, int p0, int p1,
                                          int g, int t) {
  const int sw = t << 3;                      // ((pt0 + t) & 3) << 3 for pt0 % 8 == 0; rows t and t+4 agree
  const int ac0 = (acol0 + 16 * m + g) ^ sw, ac1 = (acol0 + 16 * m + g + 8) ^ sw;
#pragma unroll 2
  for (int pt0 = p0; pt0 < p1; pt0 += 8) {
    const float *ar0 = A + (pt0 + t) * LDA, *ar1 = ar0 + 4 * LDA;
    const float a0 = ar0[ac0], a1 = ar0[ac1], a2 = ar1[ac0], a3 = ar1[ac1];
    bsum[0] += a0 + a2;
    bsum[1] += a1 + a3;
    uint32_t ah[4], al[4];
    split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]);
    split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
    const float *br0 = B + (pt0 + t) * LDB, *br1 = br0 + 4 * LDB;
    uint32_t bh[NTILES][2], bl[NTILES][2];
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) {
      const int bc = (bcol0 + 8 * nt + g) ^ sw;
      split_tf32(br0[bc], bh[nt][0], bl[nt][0]);
      split_tf32(br1[bc], bh[nt][1], bl[nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) mma_tf32(d[nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) mma_tf32(d[nt], ah, bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) mma_tf32(d[nt], ah, bh[nt][0], bh[nt][1]);
  }
}

// emit a [16 x 8*NTILES] output strip: out[(16m + n) * ldw + kofs + k] += d, k < klim
template <int NTILES>
__device__ __forceinline__ void emit_strip(float *__restrict__ out, int ldw, int kofs, int klim, int m,
                                           const float (&d)[NTILES][4], int g, int t) {
#pragma unroll
  for (int nt = 0; nt < NTILES; ++nt) {
    const int k = 8 * nt + 2 * t;
    float *r0 = out + (16 * m + g) * ldw + kofs, *r1 = r0 + 8 * ldw;
    if (k < klim) { atomicAdd(r0 + k, d[nt][0]); atomicAdd(r1 + k, d[nt][2]); }
    if (k + 1 < klim) { atomicAdd(r0 + k + 1, d[nt][1]); atomicAdd(r1 + k + 1, d[nt][3]); }
  }
}
// bias: column sums held per lane for rows 16m+g and 16m+g+8, partial over t
__device__ __forceinline__ void emit_bias(float *__restrict__ out, int m, float (&bsum)[2], int g, int t) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v = bsum[h];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) atomicAdd(out + 16 * m + g + 8 * h, v);
  }
}

// ---------------------------------------------------------------------------------------------
// kernel configuration
// ---------------------------------------------------------------------------------------------
// RECOMP = true: the forward saved nothing usable, recompute it here (steps 1-2 above).  RECOMP = false: relu masks
// (and, for WG, the activation tiles h_0..h_4) come from the buffer the forward kernel wrote.
// SPLIT = true (with WG = false, RECOMP = false): the "split" mapping backward.  This kernel does everything but the
// weight-gradient GEMMs -- it has the 2-CTA/SM footprint of the pose-only variant -- and writes the five g_h tiles,
// per-point relu mask words and the points to a scratch; wgrad_split_kernel (below) turns them into decoder gradients.
template <int STAGE, bool WG, bool RECOMP, bool SPLIT = false>
struct BwdCfg {
  static constexpr int NT = (WG || !RECOMP) ? 192 : 256;
  static constexpr int NW = NT / 32;
  // feature tile row length: 64 when a fine decoder needs [fine | middle] features; the pose-only saved-forward
  // variant never gathers features (the tile is only the staging area of the trilinear backward): 32
  static constexpr int RS = (WG || RECOMP) ? MmaStage<STAGE>::RS : 32;
  // that variant (tracking, event render: thousands of rays) is tuned for throughput: 102 KB of shared memory and
  // <= 168 registers, so two CTAs (12 warps) share an SM
  static constexpr int MIN_CTAS = (WG || RECOMP) ? 1 : 2;
  static constexpr int WREG = RECOMP ? MmaStage<STAGE>::WMAX : MlpPackV2B::total();   // forward blob >= backward blob
  static constexpr int TILE = NT * 32;                                  // floats of one staging tile
  static constexpr int NTILES = WG ? 3 : 0;                             // sG, sG2, sX
  static constexpr int MISC_BYTES = NT * 56;                            // placement / compositing scratch, later sgp
  static constexpr size_t smem_bytes() {
    return (size_t)(WREG + NT * RS + NTILES * TILE + (WG ? NT * 4 : 0)) * 4 + MISC_BYTES;
  }
};
static_assert(MlpPackV2B::total() <= MlpPackV2<32>::total(), "backward blob must fit the forward weight region");

template <int CD, int NO>
__device__ __forceinline__ int grad_off_W(int i) {      // runtime-i MlpGrad<CD,NO>::off_W
  return MlpGrad<CD, NO>::off_W(0) + (i >= 1 ? 32 * 93 + 32 : 0) + (i >= 2 ? 1056 : 0) + (i >= 3 ? 1056 : 0) +
         (i >= 4 ? 32 * 125 + 32 : 0);
}

struct WarpCtx {
  int lane, g, t, warp;
  float *crow;       // warp's feature-tile rows
  float *sG, *sG2, *sX;   // CTA tiles (WG)
  float *sP;         // [NT][4] points (WG)
  float *hs;         // RECOMP: global scratch of this warp, 4 tiles of 1024 floats (WG)
  int64_t gtile;     // global index of this warp's 32-point tile (saved-for-backward layout)
};

// ---------------------------------------------------------------------------------------------
// one decoder: recompute + backward for the warp's 32 points.  CTA-collective.
//   gout[NO]   : d L / d (decoder outputs) of the OWNER lane's point
//   p32        : owner lane's p.float();  pn: its normalised coordinates
//   gp         : owner lane's accumulated d L / d p (float64)
// ---------------------------------------------------------------------------------------------
template <int STAGE, bool WG, bool RECOMP, bool SPLIT, int LEVEL, int CD, int NO>
__device__ __forceinline__ void decoder_bwd_mma(const BwdArgs &a, float *__restrict__ sw, const WarpCtx &w,
                                                const float pn[3], const float p32[3], const float (&gout)[NO],
                                                bool valid, bool want_rays, double gp[3]) {
  using CFG = BwdCfg<STAGE, WG, RECOMP, SPLIT>;
  constexpr int RS = CFG::RS;
  constexpr int NT = CFG::NT;
  constexpr int DEC = (LEVEL == ENS_LEVEL_MIDDLE) ? 0 : (LEVEL == ENS_LEVEL_FINE ? 1 : 2);
  using PF = MlpPackV2<CD>;
  using PB = MlpPackV2B;
  using GO = MlpGrad<CD, NO>;
  constexpr int C0 = (LEVEL == ENS_LEVEL_MIDDLE && RS == 64) ? 32 : 0;   // where this decoder's own features live
  const int lane = w.lane, g = w.g, t = w.t;
  float *gdec = a.gdec[LEVEL];
  float *ggrid = a.ggrid[LEVEL];
  const bool need_emb = WG || SPLIT || want_rays;

  // ---- 1. stage the weights (RECOMP: forward blob; otherwise directly the transposed blob), gather features ----
  // activation tiles h_0..h_4 of this warp: own scratch when recomputed, else what the forward saved
  const float *hsrc = RECOMP ? w.hs : ((WG || SPLIT) ? a.save_h + ((int64_t)DEC * a.n_tiles + w.gtile) * 5120 : nullptr);
  __syncthreads();
  if (RECOMP) stage_blob(sw, a.sc.w[LEVEL] + off_v2<CD>(), PF::total());
  else stage_blob(sw, a.sc.w[LEVEL] + off_v2b<CD>(), PB::total());
  const Vox v = make_vox(pn, a.sc.dims[LEVEL]);
  if (RECOMP || WG) gather_warp<RS>(a.sc.grid[LEVEL], a.sc.dims[LEVEL], v, w.crow, C0);   // features: recompute / dWc operand
  stage_blob_wait();
  __syncthreads();

  float rx[4], ry[4], rz[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    rx[j] = __shfl_sync(0xffffffffu, p32[0], g + 8 * j);
    ry[j] = __shfl_sync(0xffffffffu, p32[1], g + 8 * j);
    rz[j] = __shfl_sync(0xffffffffu, p32[2], g + 8 * j);
  }

  // ---- 2. forward recompute ----
  uint32_t mask[5];
  float gh[2][4][4];
  {
    float acc[2][4][4];
    if (RECOMP) {
    float acc3[2][4][4];
    set_bias(acc, sw + PF::off_L(0) + PF::in_b(), t);
    set_bias(acc3, sw + PF::off_L(3) + PF::in_b(), t);
#pragma unroll 1
    for (int kt = 0; kt < EMBP / 8; ++kt) {
      const float *B = sw + PF::off_B() + 8 * kt + 2 * t;
      const float2 b0 = *reinterpret_cast<const float2 *>(B);
      const float2 b1 = *reinterpret_cast<const float2 *>(B + EMBP);
      const float2 b2 = *reinterpret_cast<const float2 *>(B + 2 * EMBP);
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * m + h;
          const float q0 = fmaf(rz[j], b2.x, fmaf(ry[j], b1.x, rx[j] * b0.x));
          const float q1 = fmaf(rz[j], b2.y, fmaf(ry[j], b1.y, rx[j] * b0.y));
          split_tf32(fast_sin(q0), ah[m][h], al[m][h]);
          split_tf32(fast_sin(q1), ah[m][2 + h], al[m][2 + h]);
        }
      }
      mma_ktile<EMBP>(acc, ah, al, sw + PF::off_W0(), kt, g, t);
      mma_ktile<EMBP>(acc3, ah, al, sw + PF::off_W3e(), kt, g, t);
    }
    float x[2][4][4];
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
      const float *L = sw + PF::off_L(0) + i * PF::block_floats();
      if (i > 0) {
        ENS_FOR_TILE(m, nt, e) {
          x[m][nt][e] = acc[m][nt][e];
          acc[m][nt][e] = acc3[m][nt][e];
        }
        if (i != 3) set_bias(acc, L + PF::in_b(), t);
        gemm_hidden(acc, x, L + PF::in_Wh(), g, t);
      }
      const uint32_t mk = relu_add_bias_mask(acc, L + PF::in_bc(), t);
#pragma unroll
      for (int k = 0; k < 5; ++k) if (k == i) mask[k] = mk;
      gemm_features<CD, RS>(acc, w.crow, C0 == 32 ? 32 : 0, L + PF::in_Wc(), g, t);
      if (WG && i < 4) store_tile<32>(w.hs + i * 1024, 0, acc, g, t);      // h_i = x_{i+1}
    }
    } else {
      const uint32_t *mt = a.save_masks + (((int64_t)DEC * a.n_tiles + w.gtile) * 5) * 32;
      if (a.mask_fmt == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) mask[k] = mt[k * 32 + lane];
      } else {
        // one word per point (bit n = unit n), written by the tcgen05 forward: gather the words of my four rows and
        // re-order their bits into the accumulator-fragment order  bit ((m*4 + nt)*4 + e) <- row 16m + g + 8(e>>1),
        // unit 8nt + 2t + (e&1)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          uint32_t mk = 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {                          // row g + 8j: m = j >> 1, e>>1 = j & 1
            const uint32_t pw = mt[k * 32 + g + 8 * j];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              mk |= ((pw >> (8 * nt + 2 * t)) & 3u) << ((((j >> 1) * 4 + nt) * 4) + 2 * (j & 1));
          }
          mask[k] = mk;
        }
      }
      if (WG || SPLIT) load_tile<32>(hsrc + 4 * 1024, 0, acc, g, t);        // h_4 for dWo
    }
    // decoder-output gradients of my four rows
    float gr[4][NO];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 0; o < NO; ++o) gr[j][o] = __shfl_sync(0xffffffffu, gout[o], g + 8 * j);
    if (WG || SPLIT) {
      // dWo[o][n] = sum_pt gout[pt][o] h4[pt][n]; dbo[o] = sum_pt gout[pt][o].  Per lane: partial over its four rows for
      // its eight columns (slot cs = 2 nt + c); reduce-scatter over the eight row groups (4 + 2 + 1 shuffles per output).
      const bool b0 = g & 1, b1 = (g >> 1) & 1, b2 = (g >> 2) & 1;
      const int cs_mine = 4 * (int)b0 + 2 * (int)b1 + (int)b2;
      const int col_mine = 8 * (cs_mine >> 1) + 2 * t + (cs_mine & 1);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        float sv[8];
#pragma unroll
        for (int cs = 0; cs < 8; ++cs) {
          const int nt = cs >> 1, c = cs & 1;
          float s = gr[0][o] * acc[0][nt][c];
          s = fmaf(gr[1][o], acc[0][nt][2 + c], s);
          s = fmaf(gr[2][o], acc[1][nt][c], s);
          s = fmaf(gr[3][o], acc[1][nt][2 + c], s);
          sv[cs] = s;
        }
        float v4[4], v2[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float mine = b0 ? sv[4 + k] : sv[k], send = b0 ? sv[k] : sv[4 + k];
          v4[k] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float mine = b1 ? v4[2 + k] : v4[k], send = b1 ? v4[k] : v4[2 + k];
          v2[k] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        const float mine = b2 ? v2[1] : v2[0], send = b2 ? v2[0] : v2[1];
        atomicAdd(gdec + GO::off_Wo() + o * 32 + col_mine, mine + __shfl_xor_sync(0xffffffffu, send, 16));
        float sb = gout[o];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, off);
        if (lane == 0) atomicAdd(gdec + GO::off_bo() + o, sb);
      }
    }
    // g_h4 = Wo^T gout
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float2 wo[NO];
#pragma unroll
      for (int o = 0; o < NO; ++o)
        wo[o] = *reinterpret_cast<const float2 *>(sw + (RECOMP ? PF::off_Wo() : PB::off_Wo()) + o * 32 + 8 * nt + 2 * t);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 2 * m + (e >> 1);
          float s = 0.f;
#pragma unroll
          for (int o = 0; o < NO; ++o) s = fmaf(gr[j][o], (e & 1) ? wo[o].y : wo[o].x, s);
          gh[m][nt][e] = s;
        }
      }
    }
  }

  // ---- 3. stage the backward (transposed) blob ----
  if (RECOMP) {
    __syncthreads();
    stage_blob(sw, a.sc.w[LEVEL] + off_v2b<CD>(), PB::total());
    stage_blob_wait();
    __syncthreads();
  }
  if (WG) cp_async_tile(w.sX, hsrc + 3 * 1024, lane);                     // x_4 = h_3

  // ---- 4. blocks 4..0 ----
  // PARK_GC: in the two-CTA/SM variants (168 registers) the running feature gradient g_c lives in the warp's -- then
  // idle -- feature tile between blocks instead of in 32 registers; after block 0 the tile is what step 5 wants anyway.
  constexpr bool PARK_GC = !WG && !RECOMP;
  float gc[2][4][4], gu[2][4][4], gu3[2][4][4];
  if (!PARK_GC) zero_tile(gc);
#pragma unroll 1
  for (int i = 4; i >= 0; --i) {
    const float *L = sw + PB::off_L(0) + i * 2048;
    uint32_t mk = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) if (k == i) mk = mask[k];
    if (PARK_GC) {
      if (i == 4) zero_tile(gc); else load_tile<RS>(w.crow, 0, gc, g, t);
      gemm_hidden(gc, gh, L + PB::in_WcT(), g, t);                        // g_c += g_h Wc[:, :32]
      store_tile<RS>(w.crow, 0, gc, g, t);
    }
    ENS_FOR_TILE(m, nt, e) gu[m][nt][e] = ((mk >> ((m * 4 + nt) * 4 + e)) & 1u) ? gh[m][nt][e] : 0.f;
    if (WG) {
      store_tile<32>(w.sG, 0, gh, g, t);
      store_tile<32>(w.sG2, 0, gu, g, t);
    }
    if (SPLIT) {
      float *gdst = a.split_gh + (((int64_t)DEC * a.n_tiles + w.gtile) * 5 + i) * 1024;
      store_tile<32>(gdst, 0, gh, g, t);
      // relu mask of block i as one word per POINT (bit n = unit n): quad OR of the lane-layout bits
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t pw = 0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const int e = 2 * (j & 1) + c2;
            pw |= ((mk >> (((j >> 1) * 4 + nt) * 4 + e)) & 1u) << (8 * nt + 2 * t + c2);
          }
        pw |= __shfl_xor_sync(0xffffffffu, pw, 1);
        pw |= __shfl_xor_sync(0xffffffffu, pw, 2);
        if (t == j) a.split_mw[(((int64_t)DEC * a.n_tiles + w.gtile) * 5 + i) * 32 + g + 8 * j] = pw;
      }
    }
    if (i == 3) { ENS_FOR_TILE(m, nt, e) gu3[m][nt][e] = gu[m][nt][e]; }
    if (!PARK_GC) gemm_hidden(gc, gh, L + PB::in_WcT(), g, t);            // g_c += g_h Wc[:, :32]
    if (i > 0) {
      zero_tile(gh);
      gemm_hidden(gh, gu, L + PB::in_WhT(), g, t);                        // g_h_{i-1} = g_u W_i (hidden part)
    }
    if (WG) {
      cp_async_wait_all();
      __syncthreads();
      // items: round r = B source (0: x tile -> dW_i ; 1: own features -> dWc_i[:, :32] ; 2: concat half);
      // warp -> (row tile m = warp / 3, third of the points warp % 3)
      const int m = w.warp / 3, p0 = (w.warp % 3) * (NT / 3), p1 = p0 + NT / 3;
      const float *feat = w.crow - w.warp * 32 * RS;                      // CTA feature tile
      const float *G = w.sG - w.warp * 1024, *G2 = w.sG2 - w.warp * 1024, *Xt = w.sX - w.warp * 1024;
      if (i > 0) {
        float d[4][4] = {}, bs[2] = {0.f, 0.f};
        wgrad_mma<32, 32, 4>(d, bs, G2, 0, m, Xt, 0, p0, p1, g, t);
        const int K = (i == 3) ? 125 : 32;
        float *out = gdec + grad_off_W<CD, NO>(i);
        emit_strip<4>(out, K, (i == 3) ? EMB : 0, 32, m, d, g, t);
        emit_bias(out + 32 * K, m, bs, g, t);
      }
      {
        float d[4][4] = {}, bs[2] = {0.f, 0.f};
        wgrad_mma<32, RS, 4>(d, bs, G, 0, m, feat, C0, p0, p1, g, t);
        float *out = gdec + GO::off_Wc(0) + i * (32 * CD + 32);
        emit_strip<4>(out, CD, 0, 32, m, d, g, t);
        emit_bias(out + 32 * CD, m, bs, g, t);
      }
      if (CD == 64) {
        float d[4][4] = {}, bs[2] = {0.f, 0.f};
        wgrad_mma<32, RS, 4>(d, bs, G, 0, m, feat, 32, p0, p1, g, t);
        float *out = gdec + GO::off_Wc(0) + i * (32 * CD + 32);
        emit_strip<4>(out, CD, 32, 32, m, d, g, t);
      }
      __syncthreads();
      if (i >= 2) cp_async_tile(w.sX, hsrc + (i - 2) * 1024, lane);       // x_{i-1} = h_{i-2}
    }
  }
  // gu = g_u0 (registers; WG: also in sG2), gu3 = g_u3.

  // ---- 5. trilinear backward ----
  if (ggrid != nullptr || want_rays) {
    __syncwarp();
    if (!PARK_GC) store_tile<RS>(w.crow, 0, gc, g, t);
    __syncwarp();
    float gpn[3];
    gather_bwd_warp<RS>(a.sc.grid[LEVEL], ggrid, a.sc.dims[LEVEL], v, valid, w.crow, want_rays, gpn);
    if (want_rays) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        gp[k] += __ddiv_rn((double)gpn[k] * 2.0, __dsub_rn(a.sc.hi[k], a.sc.lo[k]));
    }
    __syncwarp();
  }

  // ---- 6. Fourier embedding backward, three 32-column chunks ----
  if (need_emb) {
    if (WG) store_tile<32>(w.sG, 0, gu3, g, t);
    float gpp[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) gpp[j][0] = gpp[j][1] = gpp[j][2] = 0.f;
#pragma unroll 1
    for (int jc = 0; jc < 3; ++jc) {
      float acc[2][4][4];
      zero_tile(acc);
      gemm_hidden(acc, gu, sw + PB::off_W0T() + jc * 1024, g, t);
      gemm_hidden(acc, gu3, sw + PB::off_W3eT() + jc * 1024, g, t);
      float ev[2][4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float *B = sw + PB::off_B() + 32 * jc + 8 * nt + 2 * t;
        const float2 b0 = *reinterpret_cast<const float2 *>(B);
        const float2 b1 = *reinterpret_cast<const float2 *>(B + EMBP);
        const float2 b2 = *reinterpret_cast<const float2 *>(B + 2 * EMBP);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 2 * m + (e >> 1);
            const float bx = (e & 1) ? b0.y : b0.x, by = (e & 1) ? b1.y : b1.x, bz = (e & 1) ? b2.y : b2.x;
            const float q = fmaf(rz[j], bz, fmaf(ry[j], by, rx[j] * bx));
            float sq, cq;
            fast_sincos(q, sq, cq);
            const float gq = acc[m][nt][e] * cq;
            acc[m][nt][e] = gq;
            ev[m][nt][e] = sq;
            gpp[j][0] = fmaf(bx, gq, gpp[j][0]);
            gpp[j][1] = fmaf(by, gq, gpp[j][1]);
            gpp[j][2] = fmaf(bz, gq, gpp[j][2]);
          }
        }
      }
      if (SPLIT) {
        // dB[r][32jc + col] = sum_pt p[pt][r] g_q[pt][col].  Per lane: partial over its four rows for its eight
        // columns (slot cs = 2 nt + c2) x three components; then a reduce-SCATTER over the eight row groups (lane bits
        // 2..4): each stage halves the slots a lane keeps, 12 + 6 + 3 shuffles instead of a 72-shuffle butterfly, and
        // every lane ends with ONE column's three sums.
        float sv[8][3];
#pragma unroll
        for (int cs = 0; cs < 8; ++cs) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float q = acc[j >> 1][cs >> 1][2 * (j & 1) + (cs & 1)];
            s0 = fmaf(rx[j], q, s0); s1 = fmaf(ry[j], q, s1); s2 = fmaf(rz[j], q, s2);
          }
          sv[cs][0] = s0; sv[cs][1] = s1; sv[cs][2] = s2;
        }
        const bool b0 = g & 1, b1 = (g >> 1) & 1, b2 = (g >> 2) & 1;
        float v4[4][3], v2[2][3], v1[3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float mine = b0 ? sv[4 + k][r] : sv[k][r], send = b0 ? sv[k][r] : sv[4 + k][r];
            v4[k][r] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float mine = b1 ? v4[2 + k][r] : v4[k][r], send = b1 ? v4[k][r] : v4[2 + k][r];
            v2[k][r] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float mine = b2 ? v2[1][r] : v2[0][r], send = b2 ? v2[0][r] : v2[1][r];
          v1[r] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        const int cs = 4 * (int)b0 + 2 * (int)b1 + (int)b2;
        const int col = 32 * jc + 8 * (cs >> 1) + 2 * t + (cs & 1);
        if (col < EMB) {
          atomicAdd(gdec + GO::off_B() + col, v1[0]);
          atomicAdd(gdec + GO::off_B() + EMB + col, v1[1]);
          atomicAdd(gdec + GO::off_B() + 2 * EMB + col, v1[2]);
        }
      }
      if (WG) {
        store_tile<32>(w.sX, 0, ev, g, t);
        store_tile<RS>(w.crow, 0, acc, g, t);                              // g_q chunk (feature columns 0..31 are dead)
        __syncthreads();
        const int m = w.warp / 3, p0 = (w.warp % 3) * (NT / 3), p1 = p0 + NT / 3;
        const float *feat = w.crow - w.warp * 32 * RS;
        const float *G = w.sG - w.warp * 1024, *G2 = w.sG2 - w.warp * 1024, *Xt = w.sX - w.warp * 1024;
        const int klim = EMB - 32 * jc;                                    // 93 real columns
        {   // dW0[:, 32jc ..] = g_u0^T e ; db_0
          float d[4][4] = {}, bs[2] = {0.f, 0.f};
          wgrad_mma<32, 32, 4>(d, bs, G2, 0, m, Xt, 0, p0, p1, g, t);
          float *out = gdec + GO::off_W(0);
          emit_strip<4>(out, EMB, 32 * jc, klim, m, d, g, t);
          if (jc == 0) emit_bias(out + 32 * EMB, m, bs, g, t);
        }
        {   // dW3[:, 32jc ..] (embedding half of the skip layer) = g_u3^T e
          float d[4][4] = {}, bs[2] = {0.f, 0.f};
          wgrad_mma<32, 32, 4>(d, bs, G, 0, m, Xt, 0, p0, p1, g, t);
          emit_strip<4>(gdec + GO::off_W(3), 125, 32 * jc, klim, m, d, g, t);
        }
        {   // dB[r][32jc + k] = sum_pt p[pt][r] g_q[pt][k]:  rows = k, one 8-wide column tile = (x, y, z, 0, ...)
          float d[1][4] = {}, bs[2] = {0.f, 0.f};
          const int sw8 = t << 3;
          const int ac0 = (16 * m + g) ^ sw8, ac1 = (16 * m + g + 8) ^ sw8;
#pragma unroll 2
          for (int pt0 = p0; pt0 < p1; pt0 += 8) {
            const float *ar0 = feat + (pt0 + t) * RS, *ar1 = ar0 + 4 * RS;
            uint32_t ah[4], al[4];
            split_tf32(ar0[ac0], ah[0], al[0]); split_tf32(ar0[ac1], ah[1], al[1]);
            split_tf32(ar1[ac0], ah[2], al[2]); split_tf32(ar1[ac1], ah[3], al[3]);
            const float pb0 = (g < 4) ? w.sP[(pt0 + t) * 4 + g] : 0.f;
            const float pb1 = (g < 4) ? w.sP[(pt0 + t + 4) * 4 + g] : 0.f;
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(pb0, bh0, bl0);
            split_tf32(pb1, bh1, bl1);
            mma_tf32(d[0], al, bh0, bh1);
            mma_tf32(d[0], ah, bl0, bl1);
            mma_tf32(d[0], ah, bh0, bh1);
          }
          (void)bs;
          // d: (k = 16m+g (+8), r = 2t, 2t+1)
          float *out = gdec + GO::off_B() + 32 * jc;
          const int k0 = 16 * m + g, k1 = k0 + 8;
          const int r = 2 * t;
          if (r < 3) {
            if (k0 < klim) atomicAdd(out + r * EMB + k0, d[0][0]);
            if (k1 < klim) atomicAdd(out + r * EMB + k1, d[0][2]);
          }
          if (r + 1 < 3) {
            if (k0 < klim) atomicAdd(out + (r + 1) * EMB + k0, d[0][1]);
            if (k1 < klim) atomicAdd(out + (r + 1) * EMB + k1, d[0][3]);
          }
        }
        __syncthreads();
      }
    }
    if (want_rays) {
      // quad-reduce the per-row partials and hand row j of quad g to its owner lane g + 8j
      float mine[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float s = gpp[j][k];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          const float got = __shfl_sync(0xffffffffu, s, (lane & 7) * 4);
          if ((lane >> 3) == j) mine[k] = got;
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) gp[k] += (double)mine[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int STAGE, bool WG, bool RECOMP, bool SPLIT>
__global__ void __launch_bounds__(BwdCfg<STAGE, WG, RECOMP, SPLIT>::NT, BwdCfg<STAGE, WG, RECOMP, SPLIT>::MIN_CTAS)
render_bwd_mma_kernel(BwdArgs a) {
  using CFG = BwdCfg<STAGE, WG, RECOMP, SPLIT>;
  constexpr int NT = CFG::NT;
  constexpr int RS = CFG::RS;
  extern __shared__ __align__(16) float smem[];
  float *sw = smem;
  float *sfeat = smem + CFG::WREG;
  float *tiles = sfeat + NT * RS;
  float *sP = tiles + CFG::NTILES * CFG::TILE;
  char *misc = reinterpret_cast<char *>(sP + (WG ? NT * 4 : 0));
  // misc region: placement / compositing scratch; later the per-point ray gradients
  double *zc = reinterpret_cast<double *>(misc);
  double *zs = zc + NT;
  double *sgw = zs + NT;
  float4 *sraw = reinterpret_cast<float4 *>(sgw + NT);
  float *salpha = reinterpret_cast<float *>(sraw + NT);
  float *sT = salpha + NT;
  float *sw_ = sT + NT;
  float *ssuf = sw_ + NT;
  double *sgp = reinterpret_cast<double *>(misc);            // [NT][3] doubles = 24 B/pt, reuses the region at the end

  const RayArgs &ra = a.ra;
  const int S = ra.S;
  const int rl = threadIdx.x / S, s = threadIdx.x % S;
  const int64_t ray = (int64_t)blockIdx.x * ra.rpc + rl;
  const bool valid = (rl < ra.rpc) && (ray < ra.R);
  const int64_t pidx = valid ? ray * S + s : 0;
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

  // ---- compositing forward (from the saved raw) and backward ----
  float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) raw = reinterpret_cast<const float4 *>(a.raw)[pidx];
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
      sw_[threadIdx.x + k] = __fmul_rn(salpha[threadIdx.x + k], T);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, salpha[threadIdx.x + k]), 1e-10f));
    }
    double dep = 0.0;
    for (int k = 0; k < S; ++k) dep += (double)sw_[threadIdx.x + k] * zs[threadIdx.x + k];
    double wdz = 0.0;
    for (int k = 0; k < S; ++k) wdz += (double)sw_[threadIdx.x + k] * (zs[threadIdx.x + k] - dep);
    const double gd = a.g_depth ? a.g_depth[ray] : 0.0;
    const double gv = a.g_var ? a.g_var[ray] : 0.0;
    const double gdt = gd + gv * (-2.0 * wdz);
    float gcl[3] = {0.f, 0.f, 0.f};
    if (a.g_color) { gcl[0] = a.g_color[ray * 3]; gcl[1] = a.g_color[ray * 3 + 1]; gcl[2] = a.g_color[ray * 3 + 2]; }
    for (int k = 0; k < S; ++k) {
      const float4 rk = sraw[threadIdx.x + k];
      const double zk = zs[threadIdx.x + k], dz = zk - dep;
      sgw[threadIdx.x + k] = (double)rk.x * gcl[0] + (double)rk.y * gcl[1] + (double)rk.z * gcl[2] + gdt * zk + gv * dz * dz;
    }
    float accs = 0.f;
    for (int k = S - 1; k >= 0; --k) {
      ssuf[threadIdx.x + k] = accs;
      accs += sw_[threadIdx.x + k] * (float)sgw[threadIdx.x + k];
    }
  }
  __syncthreads();
  float g_occ = 0.f, g_rgb[3] = {0.f, 0.f, 0.f};
  if (valid) {
    const float gw = (float)sgw[threadIdx.x];
    const float om = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    const float g_alpha = sT[threadIdx.x] * gw - ssuf[threadIdx.x] / om;
    g_occ = inside ? 10.f * alpha * (1.f - alpha) * g_alpha : 0.f;       // raw[~mask,3]=100 cuts the graph
    const float wv = sw_[threadIdx.x];
    if (a.g_color) {
      g_rgb[0] = wv * a.g_color[ray * 3]; g_rgb[1] = wv * a.g_color[ray * 3 + 1]; g_rgb[2] = wv * a.g_color[ray * 3 + 2];
    }
  }
  const bool want_rays = (a.g_rays_o != nullptr) || (a.g_rays_d != nullptr);
  double gp[3] = {0.0, 0.0, 0.0};

  WarpCtx w;
  w.lane = threadIdx.x & 31; w.g = w.lane >> 2; w.t = w.lane & 3; w.warp = threadIdx.x >> 5;
  w.crow = sfeat + w.warp * 32 * RS;
  w.sG = tiles + w.warp * 1024;
  w.sG2 = tiles + (WG ? CFG::TILE : 0) + w.warp * 1024;
  w.sX = tiles + (WG ? 2 * CFG::TILE : 0) + w.warp * 1024;
  w.sP = sP;
  w.hs = (WG && RECOMP) ? a.hscratch + ((size_t)blockIdx.x * CFG::NW + w.warp) * 4096 : nullptr;
  w.gtile = (int64_t)blockIdx.x * CFG::NW + w.warp;
  if (WG) *reinterpret_cast<float4 *>(sP + threadIdx.x * 4) = make_float4(p32[0], p32[1], p32[2], 0.f);
  if (SPLIT) {   // the points of this warp's tile, for the weight-gradient kernel (p.float() and normalised coordinates)
    float4 *dst = reinterpret_cast<float4 *>(a.split_pts + (w.gtile * 32 + w.lane) * 8);
    dst[0] = make_float4(p32[0], p32[1], p32[2], valid ? 1.f : 0.f);
    dst[1] = make_float4(pn[0], pn[1], pn[2], 0.f);
  }

  // decoder-parallel mode (pose-only saved path, small batches): this CTA runs only decoder blockIdx.y of the stage
  constexpr bool DP_OK = !WG && !RECOMP && !SPLIT;
  const int dsel = (DP_OK && a.dec_par) ? (int)blockIdx.y : -1;
  const float go1[1] = {g_occ};
  if (dsel < 0 || dsel == 0)
    decoder_bwd_mma<STAGE, WG, RECOMP, SPLIT, ENS_LEVEL_MIDDLE, 32, 1>(a, sw, w, pn, p32, go1, valid, want_rays, gp);
  if ((STAGE == ENS_STAGE_FINE || STAGE == ENS_STAGE_COLOR) && (dsel < 0 || dsel == 1))
    decoder_bwd_mma<STAGE, WG, RECOMP, SPLIT, ENS_LEVEL_FINE, 64, 1>(a, sw, w, pn, p32, go1, valid, want_rays, gp);
  if (STAGE == ENS_STAGE_COLOR && (dsel < 0 || dsel == 2)) {
    const float go4[4] = {g_rgb[0], g_rgb[1], g_rgb[2], 0.f};            // output 3 is overwritten (decoder.py:341)
    decoder_bwd_mma<STAGE, WG, RECOMP, SPLIT, ENS_LEVEL_COLOR, 32, 4>(a, sw, w, pn, p32, go4, valid, want_rays, gp);
  }

  // ---- points -> rays: g_o = sum_s g_p, g_d = sum_s z_s g_p ----
  if (want_rays) {
    __syncthreads();
    sgp[threadIdx.x * 3 + 0] = gp[0]; sgp[threadIdx.x * 3 + 1] = gp[1]; sgp[threadIdx.x * 3 + 2] = gp[2];
    double *zz = reinterpret_cast<double *>(misc + NT * 24);   // after sgp
    zz[threadIdx.x] = z;
    __syncthreads();
    if (valid && s < 3) {
      double so = 0.0, sd = 0.0;
      const int base = threadIdx.x - s;
      for (int k = 0; k < S; ++k) {
        const double gk = sgp[(base + k) * 3 + s];
        so += gk;
        sd += gk * zz[base + k];
      }
      if (dsel >= 0) {            // one of up to three CTAs of this ray: accumulate (the launcher zeroed the buffers)
        if (a.g_rays_o) atomicAdd(a.g_rays_o + ray * 3 + s, (float)so);
        if (a.g_rays_d) atomicAdd(a.g_rays_d + ray * 3 + s, (float)sd);
      } else {
        if (a.g_rays_o) a.g_rays_o[ray * 3 + s] = (float)so;
        if (a.g_rays_d) a.g_rays_d[ray * 3 + s] = (float)sd;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel of the split mapping backward
// ---------------------------------------------------------------------------------------------
// grid (ctas, 3 decoders); a CTA of 8 warps walks its share of the 32-point tiles of ONE decoder and keeps the whole
// decoder gradient in registers: the 30 (fine: 40) output strips of 16 x 32 -- dW_1..4, dWc_0..4, dW_0 and the
// embedding half of dW_3 in three 32-column chunks each -- are dealt round-robin to the warps, each strip four
// accumulator fragments.  Per tile: cp.async stages the five g_h tiles (from the data-gradient kernel), the
// activation tiles h_0..h_3 (from the forward), the mask words and the points; the Fourier features are recomputed
// (12 sines per thread) and the grid features re-gathered into shared tiles; then every warp runs its strips as
// 3xTF32 m16n8k8 GEMMs over the 32 points (g_u = g_h masked on the fly by the per-point words).  RED.ADD once at the end.
struct WgradArgs {
  DevScene sc;
  const float *gh;        // [3][stride][5][1024]
  const uint32_t *mw;     // [3][stride][5][32]
  const float *pts;       // [stride][32][8]
  const float *save_h;    // [3][stride][5][1024]   (forward-saved)
  int64_t n_tiles;        // tiles the data-gradient kernel wrote (its grid x 6 warps)
  int64_t stride;         // tiles per decoder in gh / mw / save_h (the forward's tile count)
  float *gdec[4];
};

template <int LDB, bool MASKED>
__device__ __forceinline__ void wgrad_strip(float (&d)[4][4], float (&bsum)[2], const float *__restrict__ A,
                                            const uint32_t *__restrict__ mwords, int m, const float *__restrict__ B,
                                            int bcol0, int g, int t) {
  const int sw = t << 3;
  const int n0 = 16 * m + g, n1 = n0 + 8;
  const int ac0 = n0 ^ sw, ac1 = n1 ^ sw;
#pragma unroll
  for (int pt0 = 0; pt0 < 32; pt0 += 8) {
    const float *ar0 = A + (pt0 + t) * 32, *ar1 = ar0 + 4 * 32;
    float a0 = ar0[ac0], a1 = ar0[ac1], a2 = ar1[ac0], a3 = ar1[ac1];
    if (MASKED) {
      const uint32_t w0 = mwords[pt0 + t], w1 = mwords[pt0 + t + 4];
      a0 = ((w0 >> n0) & 1u) ? a0 : 0.f; a1 = ((w0 >> n1) & 1u) ? a1 : 0.f;
      a2 = ((w1 >> n0) & 1u) ? a2 : 0.f; a3 = ((w1 >> n1) & 1u) ? a3 : 0.f;
    }
    bsum[0] += a0 + a2;
    bsum[1] += a1 + a3;
    uint32_t ah[4], al[4];
    split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]);
    split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
    const float *br0 = B + (pt0 + t) * LDB, *br1 = br0 + 4 * LDB;
    uint32_t bh[4][2], bl[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int bc = (bcol0 + 8 * nt + g) ^ sw;
      split_tf32(br0[bc], bh[nt][0], bl[nt][0]);
      split_tf32(br1[bc], bh[nt][1], bl[nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(d[nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(d[nt], ah, bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(d[nt], ah, bh[nt][0], bh[nt][1]);
  }
}

// trilinear features of FOUR points by one warp (8 lanes per point, 4 channels per lane), written to rows
// 4*grp .. 4*grp+3 of a swizzled [32][RS] tile at columns c0..c0+31.  Same corner / fma order as gather_warp.
template <int RS>
__device__ __forceinline__ void gather_four(const float *__restrict__ grid, const int dims[3], const float *__restrict__ sPT,
                                            int grp, float *__restrict__ crow, int c0, int lane) {
  const int cq = lane & 7, pt = 4 * grp + (lane >> 3);
  const float pn[3] = {sPT[pt * 8 + 4], sPT[pt * 8 + 5], sPT[pt * 8 + 6]};
  const Vox v = make_vox(pn, dims);
  const int X = dims[2], Y = dims[1], Z = dims[0];
  const bool okx = v.x0 + 1 < X, oky = v.y0 + 1 < Y, okz = v.z0 + 1 < Z;
  const float fx1 = okx ? v.fx : 0.f, fy1 = oky ? v.fy : 0.f, fz1 = okz ? v.fz : 0.f;
  const int ox = okx ? C : 0, oy = oky ? X * C : 0, oz = okz ? X * Y * C : 0;
  const float *p = grid + ((v.z0 * Y + v.y0) * X + v.x0) * C + 4 * cq;
  float4 a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    a[c] = __ldg(reinterpret_cast<const float4 *>(p + ((c & 1) ? ox : 0) + ((c & 2) ? oy : 0) + ((c & 4) ? oz : 0)));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float w = __fmul_rn(__fmul_rn((c & 1) ? fx1 : v.gx, (c & 2) ? fy1 : v.gy), (c & 4) ? fz1 : v.gz);
    r.x = fmaf(a[c].x, w, r.x); r.y = fmaf(a[c].y, w, r.y); r.z = fmaf(a[c].z, w, r.z); r.w = fmaf(a[c].w, w, r.w);
  }
  *reinterpret_cast<float4 *>(crow + pt * RS + ((c0 + 4 * cq) ^ ((pt & 3) << 3))) = r;
}

constexpr int WG_STAGE_FLOATS = 5 * 1024 + 4 * 1024 + 160 + 256;          // one staged tile: g_h, h_0..3, mask words, points

template <int LEVEL, int CD, int NO>
__device__ __forceinline__ void wgrad_decoder(const WgradArgs &a, float *__restrict__ smem) {
  using GO = MlpGrad<CD, NO>;
  constexpr int DEC = (LEVEL == ENS_LEVEL_MIDDLE) ? 0 : (LEVEL == ENS_LEVEL_FINE ? 1 : 2);
  constexpr int NSTRIPS = 8 + 10 * (CD / 32) + 12;          // dW_1..4 | dWc (per 32 feature columns) | dW_0, dW_3e chunks
  constexpr int PER_WARP = (NSTRIPS + 7) / 8;
  float *sE = smem + 2 * WG_STAGE_FLOATS;          // [3][1024]   sin(p.B), 32-column chunks
  float *sC = sE + 3 * 1024;                       // [32][CD]    features (swizzled rows)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const float *Bm = a.sc.w[LEVEL] + off_v2b<CD>() + MlpPackV2B::off_B();      // Fourier matrix [3][96]
  float *gdec = a.gdec[LEVEL];

  float acc[PER_WARP][4][4];
  float bsum[PER_WARP][2];
#pragma unroll
  for (int q = 0; q < PER_WARP; ++q) {
    bsum[q][0] = bsum[q][1] = 0.f;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[q][x][y] = 0.f;
  }

  const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * per, t1 = (t0 + per < a.n_tiles) ? t0 + per : a.n_tiles;

  auto prefetch = [&](int64_t T, int buf) {        // cp.async the operands of tile T into stage buffer buf
    float *st = smem + buf * WG_STAGE_FLOATS;
    const float *gsrc = a.gh + ((int64_t)DEC * a.stride + T) * 5120;
    const float *hsrc = a.save_h + ((int64_t)DEC * a.stride + T) * 5120;
    const uint32_t s_gh = (uint32_t)__cvta_generic_to_shared(st), s_h = s_gh + 5 * 4096;
    for (int i = tid; i < 1280; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_gh + i * 16), "l"(gsrc + i * 4) : "memory");
    for (int i = tid; i < 1024; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_h + i * 16), "l"(hsrc + i * 4) : "memory");
    if (tid < 40) {
      const uint32_t *msrc = a.mw + ((int64_t)DEC * a.stride + T) * 160;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_h + 4 * 4096 + tid * 16), "l"(msrc + tid * 4) : "memory");
    } else if (tid < 104) {
      const float *psrc = a.pts + T * 256;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_h + 4 * 4096 + 640 + (tid - 40) * 16), "l"(psrc + (tid - 40) * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (t0 < t1) prefetch(t0, 0);
  int buf = 0;
  for (int64_t T = t0; T < t1; ++T, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                    // tile T staged; every warp is done with tile T-1 (stage buffer buf^1, sE, sC)
    if (T + 1 < t1) prefetch(T + 1, buf ^ 1);                  // in flight during this tile's compute
    const float *sGH = smem + buf * WG_STAGE_FLOATS;
    const float *sH = sGH + 5 * 1024;
    const uint32_t *sMW = reinterpret_cast<const uint32_t *>(sH + 4 * 1024);
    const float *sPT = reinterpret_cast<const float *>(sMW + 160);
    // ---- recompute the B operands that are cheap to recompute: Fourier features and grid features ----
    {
      const int pt = tid >> 3, c0 = (tid & 7) * 12;                 // 12 of the 96 columns of one point
      const float px = sPT[pt * 8 + 0], py = sPT[pt * 8 + 1], pz = sPT[pt * 8 + 2];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int col = c0 + k;
        const float q = fmaf(pz, __ldg(Bm + 2 * EMBP + col), fmaf(py, __ldg(Bm + EMBP + col), px * __ldg(Bm + col)));
        sE[(col >> 5) * 1024 + pt * 32 + ((col & 31) ^ ((pt & 3) << 3))] = fast_sin(q);
      }
      gather_four<CD>(a.sc.grid[LEVEL], a.sc.dims[LEVEL], sPT, warp, sC, 0, lane);     // warp w: points 4w..4w+3
      if (CD == 64)                                                                   // fine decoder: [fine | middle]
        gather_four<CD>(a.sc.grid[ENS_LEVEL_MIDDLE], a.sc.dims[ENS_LEVEL_MIDDLE], sPT, warp, sC, 32, lane);
    }
    __syncthreads();
    // ---- the strips of this warp ----
#pragma unroll
    for (int q = 0; q < PER_WARP; ++q) {
      const int id = warp + 8 * q;
      if (id >= NSTRIPS) break;
      const int m = id & 1;
      if (id < 8) {                                   // dW_i, i = 1..4:  g_u_i^T h_{i-1}
        const int i = 1 + (id >> 1);
        wgrad_strip<32, true>(acc[q], bsum[q], sGH + i * 1024, sMW + i * 32, m, sH + (i - 1) * 1024, 0, g, t);
      } else if (id < 8 + 10 * (CD / 32)) {           // dWc_i, i = 0..4, 32 feature columns per strip pair:  g_h_i^T c
        const int r = id - 8, i = (r >> 1) % 5, half = r / 10;
        wgrad_strip<CD, false>(acc[q], bsum[q], sGH + i * 1024, nullptr, m, sC, 32 * half, g, t);
      } else {                                        // dW_0 / dW_3 (embedding half), chunk jc:  g_u^T e
        const int r = id - (8 + 10 * (CD / 32)), which = r / 6, jc = (r % 6) >> 1;
        const int i = which ? 3 : 0;
        wgrad_strip<32, true>(acc[q], bsum[q], sGH + i * 1024, sMW + i * 32, m, sE + jc * 1024, 0, g, t);
      }
    }
  }
  // ---- emit ----
#pragma unroll
  for (int q = 0; q < PER_WARP; ++q) {
    const int id = warp + 8 * q;
    if (id >= NSTRIPS) break;
    const int m = id & 1;
    if (id < 8) {
      const int i = 1 + (id >> 1), K = (i == 3) ? 125 : 32;
      float *out = gdec + grad_off_W<CD, NO>(i);
      emit_strip<4>(out, K, (i == 3) ? EMB : 0, 32, m, acc[q], g, t);
      emit_bias(out + 32 * K, m, bsum[q], g, t);
    } else if (id < 8 + 10 * (CD / 32)) {
      const int r = id - 8, i = (r >> 1) % 5, half = r / 10;
      float *out = gdec + GO::off_Wc(0) + i * (32 * CD + 32);
      emit_strip<4>(out, CD, 32 * half, 32, m, acc[q], g, t);
      if (half == 0) emit_bias(out + 32 * CD, m, bsum[q], g, t);
    } else {
      const int r = id - (8 + 10 * (CD / 32)), which = r / 6, jc = (r % 6) >> 1;
      if (which == 0) {
        float *out = gdec + GO::off_W(0);
        emit_strip<4>(out, EMB, 32 * jc, EMB - 32 * jc, m, acc[q], g, t);
        if (jc == 0) emit_bias(out + 32 * EMB, m, bsum[q], g, t);
      } else {
        emit_strip<4>(gdec + GO::off_W(3), 125, 32 * jc, EMB - 32 * jc, m, acc[q], g, t);
      }
    }
  }
}

constexpr int WGRAD_SMEM_FLOATS = 2 * WG_STAGE_FLOATS + 3 * 1024 + 32 * 64;

template <int STAGE>
__global__ void __launch_bounds__(256, 2) wgrad_split_kernel(WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  if (blockIdx.y == 0) wgrad_decoder<ENS_LEVEL_MIDDLE, 32, 1>(a, smem);
  else if (blockIdx.y == 1) { if (STAGE >= ENS_STAGE_FINE) wgrad_decoder<ENS_LEVEL_FINE, 64, 1>(a, smem); }
  else { if (STAGE == ENS_STAGE_COLOR) wgrad_decoder<ENS_LEVEL_COLOR, 32, 4>(a, smem); }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static bool use_decoder_parallel() {
  const char *v = std::getenv("ENS_BWD_DECODER_PARALLEL");
  return !(v && v[0] == '0');
}

template <int STAGE, bool WG, bool RECOMP, bool SPLIT>
static int launch_bwd_mma(BwdArgs &a, cudaStream_t s) {
  using CFG = BwdCfg<STAGE, WG, RECOMP, SPLIT>;
  a.ra.rpc = CFG::NT / a.ra.S;
  const size_t smem = CFG::smem_bytes();
  if (cudaFuncSetAttribute(render_bwd_mma_kernel<STAGE, WG, RECOMP, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return ENS_ECUDA;
  const unsigned g = (unsigned)((a.ra.R + a.ra.rpc - 1) / a.ra.rpc);
  unsigned gy = 1;
  a.dec_par = 0;
  if (!WG && !RECOMP && !SPLIT && use_decoder_parallel()) {
    // small pose-only batches (tracking: 200 rays = 50 CTAs) leave most SMs idle and are bound by the latency of one CTA
    // walking its decoders in turn: give every decoder its own CTA while all of them still fit in one wave
    constexpr int ndec = (STAGE == ENS_STAGE_MIDDLE) ? 1 : (STAGE == ENS_STAGE_FINE ? 2 : 3);
    if (ndec > 1 && (int64_t)g * ndec <= 2 * sm_count()) {
      a.dec_par = 1;
      gy = ndec;
      if (a.g_rays_o && cudaMemsetAsync(a.g_rays_o, 0, sizeof(float) * 3 * a.ra.R, s) != cudaSuccess) return ENS_ECUDA;
      if (a.g_rays_d && cudaMemsetAsync(a.g_rays_d, 0, sizeof(float) * 3 * a.ra.R, s) != cudaSuccess) return ENS_ECUDA;
    }
  }
  render_bwd_mma_kernel<STAGE, WG, RECOMP, SPLIT><<<dim3(g, gy), CFG::NT, smem, s>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

static bool use_split_backward() {
  const char *v = std::getenv("ENS_BWD_SPLIT");
  return !(v && v[0] == '0');
}

// layout of the split scratch inside the caller's workspace; every per-decoder array uses the FORWARD's tile count as
// its stride (the data-gradient kernel indexes the forward-saved buffers and the scratch with the same a.n_tiles)
static void split_layout(int64_t n_rays, int S, int64_t *stride, int64_t *off_mw, int64_t *off_pts, int64_t *total) {
  const int rpc_f = NT_MMA / S;
  const int64_t nt = ((n_rays + rpc_f - 1) / rpc_f) * (NT_MMA / 32);
  const int64_t gh = 3 * nt * 5120 * 4, mw = 3 * nt * 160 * 4, pts = nt * 256 * 4;
  if (stride) *stride = nt;
  if (off_mw) *off_mw = gh;
  if (off_pts) *off_pts = gh + mw;
  if (total) *total = gh + mw + pts;
}

template <int STAGE>
static int launch_bwd_mma_any(BwdArgs &a, bool wg, cudaStream_t s) {
  // the saved-forward fast path needs the masks, the activation tiles when decoder gradients are wanted, and a
  // samples-per-ray count that tiles the 192-thread CTAs (so forward and backward agree on the 32-point tiles)
  const bool saved = a.save_masks != nullptr && (192 % a.ra.S) == 0 && (!wg || a.save_h != nullptr);
  if (!saved) return wg ? launch_bwd_mma<STAGE, true, true, false>(a, s) : launch_bwd_mma<STAGE, false, true, false>(a, s);
  if (!wg) return launch_bwd_mma<STAGE, false, false, false>(a, s);
  if (!use_split_backward() || a.hscratch == nullptr) return launch_bwd_mma<STAGE, true, false, false>(a, s);
  // split mapping backward: data gradients at two CTAs per SM, then the weight-gradient GEMMs
  int64_t stride, off_mw, off_pts;
  split_layout(a.ra.R, a.ra.S, &stride, &off_mw, &off_pts, nullptr);
  if (stride != a.n_tiles) return ENS_ESHAPE;            // must be the forward's tile count
  char *base = reinterpret_cast<char *>(a.hscratch);
  a.split_gh = reinterpret_cast<float *>(base);
  a.split_mw = reinterpret_cast<uint32_t *>(base + off_mw);
  a.split_pts = reinterpret_cast<float *>(base + off_pts);
  const int rpc_b = 192 / a.ra.S;
  const int64_t tiles = ((a.ra.R + rpc_b - 1) / rpc_b) * 6;   // what the data-gradient grid covers (<= stride)
  WgradArgs wa;
  wa.sc = a.sc; wa.gh = a.split_gh; wa.mw = a.split_mw; wa.pts = a.split_pts; wa.save_h = a.save_h;
  wa.n_tiles = tiles; wa.stride = stride;
  for (int l = 0; l < 4; ++l) wa.gdec[l] = a.gdec[l];
  int rc = launch_bwd_mma<STAGE, false, false, true>(a, s);
  if (rc != ENS_OK) return rc;
  const size_t smem = (size_t)WGRAD_SMEM_FLOATS * 4;
  ENS_CUDA_CALL(cudaFuncSetAttribute(wgrad_split_kernel<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ndec = (STAGE == ENS_STAGE_MIDDLE) ? 1 : (STAGE == ENS_STAGE_FINE ? 2 : 3);
  int64_t ctas = 2 * sm_count() / ndec;
  if (ctas > tiles) ctas = tiles;
  wgrad_split_kernel<STAGE><<<dim3((unsigned)ctas, ndec), 256, smem, s>>>(wa);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

int mma_render_bwd(BwdArgs &a, int stage, bool wg, cudaStream_t s) {
  a.split_gh = nullptr; a.split_mw = nullptr; a.split_pts = nullptr;
  switch (stage) {
    case ENS_STAGE_MIDDLE: return launch_bwd_mma_any<ENS_STAGE_MIDDLE>(a, wg, s);
    case ENS_STAGE_FINE: return launch_bwd_mma_any<ENS_STAGE_FINE>(a, wg, s);
    case ENS_STAGE_COLOR: return launch_bwd_mma_any<ENS_STAGE_COLOR>(a, wg, s);
    default: return ENS_EUNSUPPORTED;
  }
}

// scratch: the recompute variant needs 4 tiles of 4 KB per warp (6 warps per CTA of 192 / S rays); the split variant
// the g_h tiles, mask words and points of every tile
int64_t mma_bwd_workspace_bytes(int64_t n_rays, int S) {
  const int rpc = 192 / S;
  if (rpc < 1) return 0;
  const int64_t ctas = (n_rays + rpc - 1) / rpc;
  const int64_t recompute = ctas * 6 * 4096 * (int64_t)sizeof(float);
  int64_t split = 0;
  if ((192 % S) == 0) split_layout(n_rays, S, nullptr, nullptr, nullptr, &split);
  return recompute > split ? recompute : split;
}

}  // namespace ens
