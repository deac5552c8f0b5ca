// Fused render kernels, variant "fma": one thread per sample point, decoder weights staged in shared
// memory and read as warp-broadcast 128-bit rows, activations in registers.
//
//   eval_points_kernel  -- Renderer.eval_points / NICE.forward      (Renderer.py:24-62, decoder.py:312-342)
//   render_fwd_kernel   -- Renderer.render_batch_ray + raw2outputs  (Renderer.py:64-199, common.py:256-297)
//   render_bwd_kernel   -- its backward
//
// A CTA owns RPC whole rays (RPC*S threads): sample placement (float64, bit-exact op order), the
// trilinear gathers, the Fourier-feature MLPs and the compositing all happen without any per-point
// tensor touching HBM.  The only per-point state that crosses kernels is `raw` (16 B/point).
#include <cstdlib>
#include <cstring>

#include "ens_device.cuh"

namespace ens {


// =============================================================================================
// decoders, thread-per-point FMA form
// =============================================================================================
__device__ __forceinline__ void axpy32(float (&acc)[32], const float *__restrict__ wrow, float x) {
  const float4 *w4 = reinterpret_cast<const float4 *>(wrow);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = w4[q];
    acc[4 * q + 0] = fmaf(a.x, x, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(a.y, x, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(a.z, x, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(a.w, x, acc[4 * q + 3]);
  }
}
__device__ __forceinline__ float dot32(const float *__restrict__ wrow, const float (&g)[32]) {
  const float4 *w4 = reinterpret_cast<const float4 *>(wrow);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = w4[q];
    s0 = fmaf(a.x, g[4 * q + 0], s0); s1 = fmaf(a.y, g[4 * q + 1], s1);
    s2 = fmaf(a.z, g[4 * q + 2], s2); s3 = fmaf(a.w, g[4 * q + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ void load32(float (&d)[32], const float *__restrict__ src) {
  const float4 *s4 = reinterpret_cast<const float4 *>(src);
#pragma unroll
  for (int q = 0; q < 8; ++q) { const float4 a = s4[q]; d[4 * q] = a.x; d[4 * q + 1] = a.y; d[4 * q + 2] = a.z; d[4 * q + 3] = a.w; }
}
__device__ __forceinline__ void store32(float *__restrict__ dst, const float (&s)[32]) {
  float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
  for (int q = 0; q < 8; ++q) d4[q] = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
}

// h = relu(acc) + bc + Wc^T c ; records the relu mask                 (decoder.py:195-197)
template <int CD>
__device__ __forceinline__ void block_epilogue(float (&h)[32], const float (&acc)[32], uint32_t &mask,
                                               const float *__restrict__ WcT, const float *__restrict__ bc,
                                               const float *__restrict__ crow) {
  float bcv[32];
  load32(bcv, bc);
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (acc[j] > 0.f) m |= (1u << j);
    h[j] = fmaxf(acc[j], 0.f) + bcv[j];
  }
  mask = m;
  const float4 *c4 = reinterpret_cast<const float4 *>(crow);
#pragma unroll 2
  for (int k4 = 0; k4 < CD / 4; ++k4) {
    const float4 cv = c4[k4];
    axpy32(h, WcT + (4 * k4 + 0) * 32, cv.x);
    axpy32(h, WcT + (4 * k4 + 1) * 32, cv.y);
    axpy32(h, WcT + (4 * k4 + 2) * 32, cv.z);
    axpy32(h, WcT + (4 * k4 + 3) * 32, cv.w);
  }
}

template <int CD>
__device__ __forceinline__ int mlp_off_layer(int i) {   // runtime-i version of MlpPack<CD>::off_layer
  return MlpPack<CD>::off_layer(1) + (i - 1) * MlpPack<CD>::layer_floats(1) + (i > 3 ? (125 - 32) * 32 : 0);
}

// MLP.forward after the gather (decoder.py:189-203).  sw: packed weights in shared memory; crow: this
// point's CD features.  SAVE: write h0..h4 (5x32 floats) to hsave for the weight-gradient pass.
template <int CD, int NO, bool SAVE>
__device__ __forceinline__ void mlp_forward(const float *__restrict__ sw, const float *__restrict__ crow, float px,
                                            float py, float pz, float (&out)[NO], uint32_t (&mask)[5],
                                            float *__restrict__ hsave) {
  using P = MlpPack<CD>;
  float acc[32], acc3[32], h[32];
  load32(acc, sw + P::off_b(0));
  load32(acc3, sw + P::off_b(3));
  {
    const float *B = sw + P::off_B();
    const float *W0 = sw + P::off_W(0);
    const float *W3 = sw + P::off_W(3);
#pragma unroll 1
    for (int k = 0; k < EMB; ++k) {
      const float q = fmaf(pz, B[2 * EMBP + k], fmaf(py, B[EMBP + k], px * B[k]));
      const float e = sinf(q);
      axpy32(acc, W0 + k * 32, e);      // layer 0
      axpy32(acc3, W3 + k * 32, e);     // embedding half of the skip layer, accumulated early
    }
  }
  block_epilogue<CD>(h, acc, mask[0], sw + P::off_Wc(0), sw + P::off_bc(0), crow);
  if (SAVE) store32(hsave, h);
#pragma unroll 1
  for (int i = 1; i < 5; ++i) {
    const int ol = mlp_off_layer<CD>(i);
    const int K = (i == 3) ? 125 : 32;
    const float *W = sw + ol + ((i == 3) ? EMB * 32 : 0);
    if (i == 3) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = acc3[j];
    } else {
      load32(acc, sw + ol + K * 32);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) axpy32(acc, W + k * 32, h[k]);
    block_epilogue<CD>(h, acc, mask[i], sw + ol + K * 32 + 32, sw + ol + K * 32 + 32 + CD * 32, crow);
    if (SAVE) store32(hsave + i * 32, h);
  }
  const float *Wo = sw + P::off_Wo();
  const float *bo = sw + P::off_bo();
#pragma unroll
  for (int o = 0; o < NO; ++o) out[o] = bo[o];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float4 w = *reinterpret_cast<const float4 *>(Wo + 4 * j);
    out[0] = fmaf(w.x, h[j], out[0]);
    if (NO > 1) { out[1] = fmaf(w.y, h[j], out[1]); out[2] = fmaf(w.z, h[j], out[2]); out[3] = fmaf(w.w, h[j], out[3]); }
  }
}

// MLP_no_xyz.forward after the gather (decoder.py:262-274).
template <bool SAVE>
__device__ __forceinline__ float coarse_forward(const float *__restrict__ sw, const float *__restrict__ crow,
                                                uint32_t (&mask)[5], float *__restrict__ hsave) {
  using P = CoarsePack;
  float acc[32], h[32];
#pragma unroll 1
  for (int i = 0; i < 5; ++i) {
    const int ow = P::off_W(1) * i + (i > 3 ? 32 * 32 : 0);   // layers are 1056 floats, layer 3 is 2080
    const int K = (i == 3) ? 64 : 32;
    load32(acc, sw + ow + K * 32);
    if (i == 0 || i == 3) {       // input (or first half of the skip concat) is the feature vector c
      const float4 *c4 = reinterpret_cast<const float4 *>(crow);
#pragma unroll 2
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 cv = c4[k4];
        axpy32(acc, sw + ow + (4 * k4 + 0) * 32, cv.x);
        axpy32(acc, sw + ow + (4 * k4 + 1) * 32, cv.y);
        axpy32(acc, sw + ow + (4 * k4 + 2) * 32, cv.z);
        axpy32(acc, sw + ow + (4 * k4 + 3) * 32, cv.w);
      }
    }
    if (i != 0) {
      const float *W = sw + ow + ((i == 3) ? 32 * 32 : 0);
#pragma unroll
      for (int k = 0; k < 32; ++k) axpy32(acc, W + k * 32, h[k]);
    }
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) { if (acc[j] > 0.f) m |= (1u << j); h[j] = fmaxf(acc[j], 0.f); }
    mask[i] = m;
    if (SAVE) store32(hsave + i * 32, h);
  }
  float out = sw[P::off_bo()];
#pragma unroll
  for (int j = 0; j < 32; ++j) out = fmaf(sw[P::off_Wo() + 4 * j], h[j], out);
  return out;
}

// =============================================================================================
// CTA-wide  out(j,k) += sum_pt G[pt][j] * X[pt][k]   (weight gradients; j < 32, k < KX)
// Threads own 4x4 output tiles; the point range is split over thread groups; results are reduced
// straight into the global gradient buffer with RED (no return value).
// =============================================================================================
template <int NT, int KX, bool BIAS_G, bool BIAS_X, class EmitW, class EmitBG, class EmitBX>
__device__ __forceinline__ void cta_outer(const float *__restrict__ G, int gs, const float *__restrict__ X, int xs,
                                          EmitW emit, EmitBG emit_bg, EmitBX emit_bx) {
  constexpr int KT = KX / 4;
  constexpr int TILES = 8 * KT;
  constexpr int GROUPS = NT / TILES;
  constexpr int PPG = (NT + GROUPS - 1) / GROUPS;
  const int t = threadIdx.x;
  const int grp = t / TILES;
  if (grp >= GROUPS) return;
  const int tile = t % TILES;
  const int j0 = (tile / KT) * 4, k0 = (tile % KT) * 4;
  float acc[4][4];
  float bg[4] = {0.f, 0.f, 0.f, 0.f}, bx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int p0 = grp * PPG, p1 = min(NT, p0 + PPG);
#pragma unroll 4
  for (int pt = p0; pt < p1; ++pt) {
    const float4 g = *reinterpret_cast<const float4 *>(G + pt * gs + j0);
    const float4 x = *reinterpret_cast<const float4 *>(X + pt * xs + k0);
    const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(gv[a], xv[b], acc[a][b]);
      if (BIAS_G) bg[a] += gv[a];
      if (BIAS_X) bx[a] += xv[a];
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 4; ++b) emit(j0 + a, k0 + b, acc[a][b]);
    if (BIAS_G && k0 == 0) emit_bg(j0 + a, bg[a]);
    if (BIAS_X && j0 == 0) emit_bx(k0 + a, bx[a]);
  }
}
struct NoEmit { __device__ __forceinline__ void operator()(int, float) const {} };

// =============================================================================================
// MLP backward, thread-per-point.  WG: also produce decoder weight gradients (CTA-collective; every
// thread of the CTA must call).  Outputs g_c[0..31] (gradient wrt the first 32 features -- for the
// fine decoder the middle half is no_grad, decoder.py:184-186) and g_pe (gradient wrt p.float()).
// =============================================================================================
template <int NT, int CD>
struct WgCtx {
  float *sG;        // [NT][36] staging
  float *sX;        // [NT][36] staging
  float *sQ;        // [NT][36] staging (aliases the feature rows once they are dead)
  float *gdec;      // flat gradient buffer of this decoder (global)
  const float *hrow;  // this thread's saved h0..h4 (global scratch)
};

template <int NT, int CD, int NO, bool WG>
__device__ __forceinline__ void mlp_backward(const float *__restrict__ sw, float *__restrict__ crow, int crow_stride,
                                             float *__restrict__ crow_base, float px, float py, float pz,
                                             const float (&gout)[NO], const uint32_t (&mask)[5], float (&g_c)[32],
                                             float (&g_pe)[3], const WgCtx<NT, CD> &wg) {
  using P = MlpPack<CD>;
  using GO = MlpGrad<CD, NO>;
  const int tid = threadIdx.x;
  float gh[32], gu3[32], gu[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) g_c[j] = 0.f;
  // ---- output layer ----
  {
    const float *Wo = sw + P::off_Wo();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float4 w = *reinterpret_cast<const float4 *>(Wo + 4 * j);
      float s = w.x * gout[0];
      if (NO > 1) { s = fmaf(w.y, gout[1], s); s = fmaf(w.z, gout[2], s); s = fmaf(w.w, gout[3], s); }
      gh[j] = s;
    }
    if (WG) {
      float h4[32];
      load32(h4, wg.hrow + 4 * 32);
      store32(wg.sG + tid * 36, h4);
      *reinterpret_cast<float4 *>(wg.sX + tid * 36) =
          make_float4(gout[0], NO > 1 ? gout[1] : 0.f, NO > 1 ? gout[2] : 0.f, NO > 1 ? gout[3] : 0.f);
      __syncthreads();
      float *gd = wg.gdec;
      cta_outer<NT, 4, false, true>(
          wg.sG, 36, wg.sX, 36,
          [gd](int j, int o, float v) { if (o < NO) atomicAdd(gd + GO::off_Wo() + o * 32 + j, v); }, NoEmit(),
          [gd](int o, float v) { if (o < NO) atomicAdd(gd + GO::off_bo() + o, v); });
      __syncthreads();
    }
  }
  // ---- blocks 4..0 ----
#pragma unroll 1
  for (int i = 4; i >= 0; --i) {
    const int ol = (i == 0) ? P::off_layer(0) : mlp_off_layer<CD>(i);
    const int K = P::K(i);
    const float *WcT = sw + ol + K * 32 + 32;
    if (WG) {   // dWc_i = gh^T c, dbc_i = sum gh
      store32(wg.sG + tid * 36, gh);
      __syncthreads();
      float *gd = wg.gdec;
      const int oWc = GO::off_Wc(0) + i * (32 * CD + 32);
#pragma unroll 1
      for (int ch = 0; ch < CD / 32; ++ch) {
        if (ch == 0)
          cta_outer<NT, 32, true, false>(
              wg.sG, 36, crow_base, crow_stride,
              [gd, oWc](int j, int k, float v) { atomicAdd(gd + oWc + j * CD + k, v); },
              [gd, oWc](int j, float v) { atomicAdd(gd + oWc + 32 * CD + j, v); }, NoEmit());
        else
          cta_outer<NT, 32, false, false>(
              wg.sG, 36, crow_base + 32, crow_stride,
              [gd, oWc](int j, int k, float v) { atomicAdd(gd + oWc + j * CD + 32 + k, v); }, NoEmit(), NoEmit());
      }
      __syncthreads();
    }
    // g_c[k] += Wc_i[:,k] . gh   (only the first 32 features carry gradient)
#pragma unroll
    for (int k = 0; k < 32; ++k) g_c[k] += dot32(WcT + k * 32, gh);
    // through the relu
#pragma unroll
    for (int j = 0; j < 32; ++j) gu[j] = ((mask[i] >> j) & 1u) ? gh[j] : 0.f;
    if (i == 3) {
#pragma unroll
      for (int j = 0; j < 32; ++j) gu3[j] = gu[j];
    }
    if (i > 0) {
      if (WG) {   // dW_i (hidden-input part) = gu^T h_{i-1}, db_i = sum gu
        float hp[32];
        load32(hp, wg.hrow + (i - 1) * 32);
        store32(wg.sG + tid * 36, gu);
        store32(wg.sX + tid * 36, hp);
        __syncthreads();
        float *gd = wg.gdec;
        int oW = GO::off_W(1) + (i - 1) * (32 * 32 + 32) + (i > 3 ? 32 * (125 - 32) : 0);
        const int Kin = K, koff = (i == 3) ? EMB : 0;
        cta_outer<NT, 32, true, false>(
            wg.sG, 36, wg.sX, 36,
            [gd, oW, Kin, koff](int j, int k, float v) { atomicAdd(gd + oW + j * Kin + koff + k, v); },
            [gd, oW, Kin](int j, float v) { atomicAdd(gd + oW + 32 * Kin + j, v); }, NoEmit());
        __syncthreads();
      }
      const float *W = sw + ol + ((i == 3) ? EMB * 32 : 0);
#pragma unroll
      for (int k = 0; k < 32; ++k) gh[k] = dot32(W + k * 32, gu);
    }
  }
  // gu now holds block 0's pre-activation gradient.
  if (WG) {   // db_0
    store32(wg.sG + tid * 36, gu);
  }
  // ---- Fourier embedding: g_q = (W0^T gu0 + W3e^T gu3) * cos(q) ----
  {
    const float *B = sw + P::off_B();
    const float *W0 = sw + P::off_W(0);
    const float *W3 = sw + P::off_W(3);
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f;
#pragma unroll 1
    for (int k0 = 0; k0 < EMBP; k0 += 32) {
#pragma unroll 1
      for (int kk = 0; kk < 32; ++kk) {
        const int k = k0 + kk;
        float e = 0.f, gq = 0.f;
        if (k < EMB) {
          const float b0 = B[k], b1 = B[EMBP + k], b2 = B[2 * EMBP + k];
          const float q = fmaf(pz, b2, fmaf(py, b1, px * b0));
          float cq;
          sincosf(q, &e, &cq);
          const float ge = dot32(W0 + k * 32, gu) + dot32(W3 + k * 32, gu3);
          gq = ge * cq;
          gp0 = fmaf(b0, gq, gp0); gp1 = fmaf(b1, gq, gp1); gp2 = fmaf(b2, gq, gp2);
        }
        if (WG) { wg.sX[tid * 36 + kk] = e; wg.sQ[tid * 36 + kk] = gq; }
      }
      if (WG) {
        // chunk [k0, k0+32): dW_0[:,k] = gu0^T e, dW_3[:,k] = gu3^T e, dB[:,k] = p^T g_q
        __syncthreads();
        float *gd = wg.gdec;
        const int oW0 = GO::off_W(0), oW3 = GO::off_W(3), oB = GO::off_B();
        const int kb = k0;
        if (k0 == 0)
          cta_outer<NT, 32, true, false>(
              wg.sG, 36, wg.sX, 36,
              [gd, oW0, kb](int j, int k, float v) { if (kb + k < EMB) atomicAdd(gd + oW0 + j * EMB + kb + k, v); },
              [gd, oW0](int j, float v) { atomicAdd(gd + oW0 + 32 * EMB + j, v); }, NoEmit());
        else
          cta_outer<NT, 32, false, false>(
              wg.sG, 36, wg.sX, 36,
              [gd, oW0, kb](int j, int k, float v) { if (kb + k < EMB) atomicAdd(gd + oW0 + j * EMB + kb + k, v); },
              NoEmit(), NoEmit());
        __syncthreads();
        // restage G <- gu3 for the skip layer's embedding columns, and p for dB
        store32(wg.sG + tid * 36, gu3);
        __syncthreads();
        cta_outer<NT, 32, false, false>(
            wg.sG, 36, wg.sX, 36,
            [gd, oW3, kb](int j, int k, float v) { if (kb + k < EMB) atomicAdd(gd + oW3 + j * 125 + kb + k, v); },
            NoEmit(), NoEmit());
        __syncthreads();
        *reinterpret_cast<float4 *>(wg.sG + tid * 36) = make_float4(px, py, pz, 0.f);
        __syncthreads();
        // out(j=k index in chunk, r) = sum_pt gq[pt][j] * p[pt][r]
        cta_outer<NT, 4, false, false>(
            wg.sQ, 36, wg.sG, 36,
            [gd, oB, kb](int j, int r, float v) { if (r < 3 && kb + j < EMB) atomicAdd(gd + oB + r * EMB + kb + j, v); },
            NoEmit(), NoEmit());
        __syncthreads();
        store32(wg.sG + tid * 36, gu);   // back to gu0 for the next chunk
      }
    }
    g_pe[0] = gp0; g_pe[1] = gp1; g_pe[2] = gp2;
  }
}

// MLP_no_xyz backward.  g_c gets the gradient wrt the 32 coarse features.
template <int NT, bool WG>
__device__ __forceinline__ void coarse_backward(const float *__restrict__ sw, float *__restrict__ crow_base,
                                                int crow_stride, float gout, const uint32_t (&mask)[5],
                                                float (&g_c)[32], const WgCtx<NT, 32> &wg) {
  using P = CoarsePack;
  using GO = CoarseGrad;
  const int tid = threadIdx.x;
  float gh[32], gu[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { g_c[j] = 0.f; gh[j] = sw[P::off_Wo() + 4 * j] * gout; }
  if (WG) {
    float h4[32];
    load32(h4, wg.hrow + 4 * 32);
    store32(wg.sG + tid * 36, h4);
    *reinterpret_cast<float4 *>(wg.sX + tid * 36) = make_float4(gout, 0.f, 0.f, 0.f);
    __syncthreads();
    float *gd = wg.gdec;
    cta_outer<NT, 4, false, true>(
        wg.sG, 36, wg.sX, 36, [gd](int j, int o, float v) { if (o == 0) atomicAdd(gd + GO::off_Wo() + j, v); },
        NoEmit(), [gd](int o, float v) { if (o == 0) atomicAdd(gd + GO::off_bo(), v); });
    __syncthreads();
  }
#pragma unroll 1
  for (int i = 4; i >= 0; --i) {
    const int ow = P::off_W(1) * i + (i > 3 ? 32 * 32 : 0);
    const int K = (i == 3) ? 64 : 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) gu[j] = ((mask[i] >> j) & 1u) ? gh[j] : 0.f;
    if (WG) {
      float *gd = wg.gdec;
      const int oW = GO::off_W(1) * i + (i > 3 ? 32 * 32 : 0);
      store32(wg.sG + tid * 36, gu);
      if (i != 0) { float hp[32]; load32(hp, wg.hrow + (i - 1) * 32); store32(wg.sX + tid * 36, hp); }
      __syncthreads();
      if (i == 0 || i == 3)   // feature-input columns
        cta_outer<NT, 32, true, false>(
            wg.sG, 36, crow_base, crow_stride,
            [gd, oW, K](int j, int k, float v) { atomicAdd(gd + oW + j * K + k, v); },
            [gd, oW, K](int j, float v) { atomicAdd(gd + oW + 32 * K + j, v); }, NoEmit());
      if (i != 0) {
        const int koff = (i == 3) ? 32 : 0;
        if (i == 3)
          cta_outer<NT, 32, false, false>(
              wg.sG, 36, wg.sX, 36,
              [gd, oW, K, koff](int j, int k, float v) { atomicAdd(gd + oW + j * K + koff + k, v); }, NoEmit(), NoEmit());
        else
          cta_outer<NT, 32, true, false>(
              wg.sG, 36, wg.sX, 36,
              [gd, oW, K, koff](int j, int k, float v) { atomicAdd(gd + oW + j * K + koff + k, v); },
              [gd, oW, K](int j, float v) { atomicAdd(gd + oW + 32 * K + j, v); }, NoEmit());
      }
      __syncthreads();
    }
    if (i == 0 || i == 3) {
#pragma unroll
      for (int k = 0; k < 32; ++k) g_c[k] += dot32(sw + ow + k * 32, gu);
    }
    if (i != 0) {
      const float *W = sw + ow + ((i == 3) ? 32 * 32 : 0);
#pragma unroll
      for (int k = 0; k < 32; ++k) gh[k] = dot32(W + k * 32, gu);
    }
  }
}

// =============================================================================================
// CTA helpers
// =============================================================================================
__device__ __forceinline__ void stage_weights(float *__restrict__ sw, const float *__restrict__ gw, int nfloats) {
  const float4 *src = reinterpret_cast<const float4 *>(gw);
  float4 *dst = reinterpret_cast<float4 *>(sw);
  for (int i = threadIdx.x; i < nfloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

template <int STAGE> struct StageInfo;
template <> struct StageInfo<ENS_STAGE_COARSE> { static constexpr int CP = 36; static constexpr int WMAX = CoarsePack::total(); };
template <> struct StageInfo<ENS_STAGE_MIDDLE> { static constexpr int CP = 36; static constexpr int WMAX = MlpPack<32>::total(); };
template <> struct StageInfo<ENS_STAGE_FINE> { static constexpr int CP = 68; static constexpr int WMAX = MlpPack<64>::total(); };
template <> struct StageInfo<ENS_STAGE_COLOR> { static constexpr int CP = 68; static constexpr int WMAX = MlpPack<64>::total(); };

// Decode one point for a stage.  CTA-collective (weight staging + __syncthreads).  raw = (r,g,b,occ).
// pn: normalised coords for slam.bound; pnc: for the coarse bound.  p32 = p.float().
template <int STAGE>
__device__ __forceinline__ float4 decode_stage(const DevScene &sc, float *__restrict__ sw, float *__restrict__ crow,
                                               const float pn[3], const float pnc[3], const float p32[3]) {
  float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t mask[5];
  if (STAGE == ENS_STAGE_COARSE) {
    stage_weights(sw, sc.w[ENS_LEVEL_COARSE], CoarsePack::total());
    const Vox v = make_vox(pnc, sc.dims[ENS_LEVEL_COARSE]);
    gather32(sc.grid[ENS_LEVEL_COARSE], sc.dims[ENS_LEVEL_COARSE], v, crow);
    __syncthreads();
    raw.w = coarse_forward<false>(sw, crow, mask, nullptr);
    return raw;
  }
  // middle decoder (its features double as the fine decoder's concat half: crow[32..63])
  float *cmid = crow + (StageInfo<STAGE>::CP == 68 ? 32 : 0);
  {
    stage_weights(sw, sc.w[ENS_LEVEL_MIDDLE], MlpPack<32>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_MIDDLE]);
    gather32(sc.grid[ENS_LEVEL_MIDDLE], sc.dims[ENS_LEVEL_MIDDLE], v, cmid);
    __syncthreads();
    float o[1];
    mlp_forward<32, 1, false>(sw, cmid, p32[0], p32[1], p32[2], o, mask, nullptr);
    raw.w = o[0];
  }
  if (STAGE == ENS_STAGE_FINE || STAGE == ENS_STAGE_COLOR) {
    __syncthreads();
    stage_weights(sw, sc.w[ENS_LEVEL_FINE], MlpPack<64>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_FINE]);
    gather32(sc.grid[ENS_LEVEL_FINE], sc.dims[ENS_LEVEL_FINE], v, crow);
    __syncthreads();
    float o[1];
    mlp_forward<64, 1, false>(sw, crow, p32[0], p32[1], p32[2], o, mask, nullptr);
    raw.w = __fadd_rn(o[0], raw.w);     // fine_occ + middle_occ (decoder.py:334,341)
  }
  if (STAGE == ENS_STAGE_COLOR) {
    __syncthreads();
    stage_weights(sw, sc.w[ENS_LEVEL_COLOR], MlpPack<32>::total());
    const Vox v = make_vox(pn, sc.dims[ENS_LEVEL_COLOR]);
    gather32(sc.grid[ENS_LEVEL_COLOR], sc.dims[ENS_LEVEL_COLOR], v, crow);
    __syncthreads();
    float o[4];
    mlp_forward<32, 4, false>(sw, crow, p32[0], p32[1], p32[2], o, mask, nullptr);
    raw.x = o[0]; raw.y = o[1]; raw.z = o[2];       // output 3 is overwritten by the occupancy
  }
  return raw;
}

// =============================================================================================
// eval_points
// =============================================================================================
template <int STAGE, bool F64>
__global__ void __launch_bounds__(128) eval_points_kernel(DevScene sc, const void *__restrict__ pts, int64_t n,
                                                          int apply_mask, float *__restrict__ out4) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NT = 128;
  float *sw = smem;
  float *crow = smem + ((StageInfo<STAGE>::WMAX + 3) & ~3) + threadIdx.x * StageInfo<STAGE>::CP;
  const int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x;
  const bool valid = t < n;
  float pn[3], pnc[3], p32[3];
  bool inside = true;
  if (F64) {
    double p[3] = {0.0, 0.0, 0.0};
    if (valid) { const double *pp = (const double *)pts + t * 3; p[0] = pp[0]; p[1] = pp[1]; p[2] = pp[2]; }
    normalize64(p, sc.lo, sc.hi, pn);
    normalize64(p, sc.clo, sc.chi, pnc);
#pragma unroll
    for (int k = 0; k < 3; ++k) { p32[k] = __double2float_rn(p[k]); inside &= (p[k] < sc.hi[k]) && (p[k] > sc.lo[k]); }
  } else {
    if (valid) { const float *pp = (const float *)pts + t * 3; p32[0] = pp[0]; p32[1] = pp[1]; p32[2] = pp[2]; }
    else { p32[0] = p32[1] = p32[2] = 0.f; }
    normalize32(p32, sc.lo, sc.hi, pn);
    normalize32(p32, sc.clo, sc.chi, pnc);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      inside &= (p32[k] < __double2float_rn(sc.hi[k])) && (p32[k] > __double2float_rn(sc.lo[k]));
  }
  float4 raw = decode_stage<STAGE>(sc, sw, crow, pn, pnc, p32);
  if (apply_mask && !inside) raw.w = 100.f;                     // Renderer.py:58
  if (valid) reinterpret_cast<float4 *>(out4)[t] = raw;
}


template <int STAGE, int NT>
__global__ void __launch_bounds__(NT) render_fwd_kernel(FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CP = StageInfo<STAGE>::CP;
  constexpr int WPAD = (StageInfo<STAGE>::WMAX + 3) & ~3;
  float *sw = smem;
  float *crow = smem + WPAD + threadIdx.x * CP;
  double *zc = reinterpret_cast<double *>(smem + WPAD + NT * CP);
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
  float pn[3], pnc[3], p32[3];
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[k] = __dadd_rn((double)o[k], __dmul_rn((double)d[k], z));          // Renderer.py:173-174
    p32[k] = __double2float_rn(p[k]);
    inside &= (p[k] < a.sc.hi[k]) && (p[k] > a.sc.lo[k]);               // Renderer.py:43-47
  }
  normalize64(p, a.sc.lo, a.sc.hi, pn);
  if (STAGE == ENS_STAGE_COARSE) normalize64(p, a.sc.clo, a.sc.chi, pnc);
  float4 raw = decode_stage<STAGE>(a.sc, sw, crow, pn, pnc, p32);
  if (!inside) raw.w = 100.f;
  // ---- compositing (common.py:285-296) ----
  const float alpha = 1.f / (1.f + expf(-(10.f * raw.w)));
  __syncthreads();                      // zs is reused below; everyone is done with the decode
  zs[threadIdx.x] = z;
  sraw[threadIdx.x] = raw;
  salpha[threadIdx.x] = alpha;
  __syncthreads();
  if (valid && s == 0) {                // sequential cumprod, exactly the reference's order
    float T = 1.f;
    for (int k = 0; k < S; ++k) {
      sT[threadIdx.x + k] = T;
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, salpha[threadIdx.x + k]), 1e-10f));
    }
  }
  __syncthreads();
  const float w = __fmul_rn(alpha, sT[threadIdx.x]);
  __syncthreads();
  salpha[threadIdx.x] = w;              // reuse as weights
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

// =============================================================================================
// render backward
// =============================================================================================

// one decoder's backward for this thread's point; CTA-collective when WG.
template <int NT, int LEVEL, int CD, int NO, bool WG>
__device__ __forceinline__ void decoder_bwd(const BwdArgs &a, float *sw, float *crow_base, int cp, float *sG, float *sX,
                                            const float pn[3], const float p32[3], const float (&gout)[NO],
                                            bool want_rays, bool valid, int64_t pidx, double gp[3]) {
  constexpr int FEAT_OFF = 0;
  float *crow = crow_base + threadIdx.x * cp;
  __syncthreads();
  stage_weights(sw, a.sc.w[LEVEL], MlpPack<CD>::total());
  const Vox v = make_vox(pn, a.sc.dims[LEVEL]);
  gather32(a.sc.grid[LEVEL], a.sc.dims[LEVEL], v, crow + FEAT_OFF);
  Vox vm;
  if (CD == 64) {
    vm = make_vox(pn, a.sc.dims[ENS_LEVEL_MIDDLE]);
    gather32(a.sc.grid[ENS_LEVEL_MIDDLE], a.sc.dims[ENS_LEVEL_MIDDLE], vm, crow + 32);
  }
  __syncthreads();
  uint32_t mask[5];
  float out[NO];
  float *hrow = WG ? (a.hscratch + pidx * 160) : nullptr;
  mlp_forward<CD, NO, WG>(sw, crow, p32[0], p32[1], p32[2], out, mask, hrow);
  float g_c[32], g_pe[3];
  WgCtx<NT, CD> wg{sG, sX, crow_base, a.gdec[LEVEL], hrow};
  // NOTE: sQ aliases the feature rows with stride 36 <= cp; they are dead by the time it is written
  mlp_backward<NT, CD, NO, WG>(sw, crow, cp, crow_base, p32[0], p32[1], p32[2], gout, mask, g_c, g_pe, wg);
  float gpn[3];
  gather32_bwd(a.sc.grid[LEVEL], valid ? a.ggrid[LEVEL] : nullptr, a.sc.dims[LEVEL], v, g_c, want_rays, gpn);
  if (want_rays) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // pn = ((p-lo)/(hi-lo))*2 - 1 -> .float():  g_p = (double(g_pn)*2)/(hi-lo)  + double(g_pe)
      gp[k] += __ddiv_rn((double)gpn[k] * 2.0, __dsub_rn(a.sc.hi[k], a.sc.lo[k])) + (double)g_pe[k];
    }
  }
}

template <int STAGE, int NT, bool WG>
__global__ void __launch_bounds__(NT) render_bwd_kernel(BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CP = StageInfo<STAGE>::CP;
  constexpr int WPAD = (StageInfo<STAGE>::WMAX + 3) & ~3;
  float *sw = smem;
  float *crow_base = smem + WPAD;
  float *sG = crow_base + NT * CP;
  float *sX = sG + (WG ? NT * 36 : 0);
  double *zc = reinterpret_cast<double *>(sX + (WG ? NT * 36 : 0));
  double *zs = zc + NT;
  double *sgw = zs + NT;                 // per-point dL/dw (double)
  float4 *sraw = reinterpret_cast<float4 *>(sgw + NT);
  float *salpha = reinterpret_cast<float *>(sraw + NT);
  float *sT = salpha + NT;
  float *sw_ = sT + NT;                  // weights
  float *ssuf = sw_ + NT;                // suffix sums
  double *sgp = reinterpret_cast<double *>(ssuf + NT);   // [NT][3] point gradients for the ray reduction

  const RayArgs &ra = a.ra;
  const int S = ra.S;
  const int rl = threadIdx.x / S, s = threadIdx.x % S;
  const int64_t ray = (int64_t)blockIdx.x * ra.rpc + rl;
  const bool valid = (rl < ra.rpc) && (ray < ra.R);
  const int64_t pidx = valid ? ray * S + s : ra.R * S;   // invalid lanes use the dump row of the scratch
  float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
  if (valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ra.rays_o[ray * 3 + k]; d[k] = ra.rays_d[ray * 3 + k]; }
  }
  const double z = place_sample(ra, a.sc, valid, ray, rl, s, o, d, zc, zs);
  double p[3];
  float pn[3], pnc[3], p32[3];
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[k] = __dadd_rn((double)o[k], __dmul_rn((double)d[k], z));
    p32[k] = __double2float_rn(p[k]);
    inside &= (p[k] < a.sc.hi[k]) && (p[k] > a.sc.lo[k]);
  }
  normalize64(p, a.sc.lo, a.sc.hi, pn);
  if (STAGE == ENS_STAGE_COARSE) normalize64(p, a.sc.clo, a.sc.chi, pnc);

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
    float gc[3] = {0.f, 0.f, 0.f};
    if (a.g_color) { gc[0] = a.g_color[ray * 3]; gc[1] = a.g_color[ray * 3 + 1]; gc[2] = a.g_color[ray * 3 + 2]; }
    for (int k = 0; k < S; ++k) {
      const float4 rk = sraw[threadIdx.x + k];
      const double zk = zs[threadIdx.x + k], dz = zk - dep;
      sgw[threadIdx.x + k] = (double)rk.x * gc[0] + (double)rk.y * gc[1] + (double)rk.z * gc[2] + gdt * zk + gv * dz * dz;
    }
    float acc = 0.f;
    for (int k = S - 1; k >= 0; --k) {
      ssuf[threadIdx.x + k] = acc;
      acc += sw_[threadIdx.x + k] * (float)sgw[threadIdx.x + k];
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
    const int64_t r0 = ray - rl;  (void)r0;
    if (a.g_color) {
      g_rgb[0] = wv * a.g_color[ray * 3]; g_rgb[1] = wv * a.g_color[ray * 3 + 1]; g_rgb[2] = wv * a.g_color[ray * 3 + 2];
    }
  }
  const bool want_rays = (a.g_rays_o != nullptr) || (a.g_rays_d != nullptr);
  double gp[3] = {0.0, 0.0, 0.0};

  if (STAGE == ENS_STAGE_COARSE) {
    float *crow = crow_base + threadIdx.x * CP;
    __syncthreads();
    stage_weights(sw, a.sc.w[ENS_LEVEL_COARSE], CoarsePack::total());
    const Vox v = make_vox(pnc, a.sc.dims[ENS_LEVEL_COARSE]);
    gather32(a.sc.grid[ENS_LEVEL_COARSE], a.sc.dims[ENS_LEVEL_COARSE], v, crow);
    __syncthreads();
    uint32_t mask[5];
    float *hrow = WG ? (a.hscratch + pidx * 160) : nullptr;
    (void)coarse_forward<WG>(sw, crow, mask, hrow);
    float g_c[32];
    WgCtx<NT, 32> wg{sG, sX, crow_base, a.gdec[ENS_LEVEL_COARSE], hrow};
    coarse_backward<NT, WG>(sw, crow_base, CP, g_occ, mask, g_c, wg);
    float gpn[3];
    gather32_bwd(a.sc.grid[ENS_LEVEL_COARSE], valid ? a.ggrid[ENS_LEVEL_COARSE] : nullptr, a.sc.dims[ENS_LEVEL_COARSE], v, g_c, want_rays, gpn);
    if (want_rays) {
#pragma unroll
      for (int k = 0; k < 3; ++k) gp[k] += __ddiv_rn((double)gpn[k] * 2.0, __dsub_rn(a.sc.chi[k], a.sc.clo[k]));
    }
  } else {
    const float go1[1] = {g_occ};
    decoder_bwd<NT, ENS_LEVEL_MIDDLE, 32, 1, WG>(a, sw, crow_base, CP, sG, sX, pn, p32, go1, want_rays, valid, pidx, gp);
    if (STAGE == ENS_STAGE_FINE || STAGE == ENS_STAGE_COLOR)
      decoder_bwd<NT, ENS_LEVEL_FINE, 64, 1, WG>(a, sw, crow_base, CP, sG, sX, pn, p32, go1, want_rays, valid, pidx, gp);
    if (STAGE == ENS_STAGE_COLOR) {
      const float go4[4] = {g_rgb[0], g_rgb[1], g_rgb[2], 0.f};          // output 3 is overwritten (decoder.py:341)
      decoder_bwd<NT, ENS_LEVEL_COLOR, 32, 4, WG>(a, sw, crow_base, CP, sG, sX, pn, p32, go4, want_rays, valid, pidx, gp);
    }
  }
  // ---- points -> rays: g_o = sum_s g_p, g_d = sum_s z_s g_p ----
  if (want_rays) {
    __syncthreads();
    sgp[threadIdx.x * 3 + 0] = gp[0]; sgp[threadIdx.x * 3 + 1] = gp[1]; sgp[threadIdx.x * 3 + 2] = gp[2];
    zc[threadIdx.x] = z;
    __syncthreads();
    if (valid && s < 3) {
      double so = 0.0, sd = 0.0;
      const int base = threadIdx.x - s;
      for (int k = 0; k < S; ++k) {
        const double g = sgp[(base + k) * 3 + s];
        so += g;
        sd += g * zc[base + k];
      }
      if (a.g_rays_o) a.g_rays_o[ray * 3 + s] = (float)so;
      if (a.g_rays_d) a.g_rays_d[ray * 3 + s] = (float)sd;
    }
  }
}

// =============================================================================================
// host launchers
// =============================================================================================
template <int STAGE, int NT>
static size_t fwd_smem_bytes() {
  return (size_t)(((StageInfo<STAGE>::WMAX + 3) & ~3) + NT * StageInfo<STAGE>::CP) * 4 + (size_t)NT * (8 + 8 + 16 + 4 + 4);
}
template <int STAGE, int NT, bool WG>
static size_t bwd_smem_bytes() {
  return (size_t)(((StageInfo<STAGE>::WMAX + 3) & ~3) + NT * StageInfo<STAGE>::CP + (WG ? 2 * NT * 36 : 0)) * 4 +
         (size_t)NT * (8 + 8 + 8 + 16 + 4 + 4 + 4 + 4 + 24);
}

template <int STAGE>
static int launch_eval(const DevScene &sc, const void *pts, int f64, int64_t n, int am, float *out4, cudaStream_t s) {
  const size_t smem = (size_t)(((StageInfo<STAGE>::WMAX + 3) & ~3) + 128 * StageInfo<STAGE>::CP) * 4;
  const unsigned g = (unsigned)((n + 127) / 128);
  if (f64) {
    ENS_CUDA_CALL(cudaFuncSetAttribute(eval_points_kernel<STAGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_points_kernel<STAGE, true><<<g, 128, smem, s>>>(sc, pts, n, am, out4);
  } else {
    ENS_CUDA_CALL(cudaFuncSetAttribute(eval_points_kernel<STAGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_points_kernel<STAGE, false><<<g, 128, smem, s>>>(sc, pts, n, am, out4);
  }
  ENS_CHECK_CUDA();
  return ENS_OK;
}


template <int STAGE>
static int launch_fwd(FwdArgs &a, cudaStream_t s) {
  constexpr int NT = NT_RENDER;
  a.ra.rpc = NT / a.ra.S;
  const size_t smem = fwd_smem_bytes<STAGE, NT>();
  ENS_CUDA_CALL(cudaFuncSetAttribute(render_fwd_kernel<STAGE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned g = (unsigned)((a.ra.R + a.ra.rpc - 1) / a.ra.rpc);
  render_fwd_kernel<STAGE, NT><<<g, NT, smem, s>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}

template <int STAGE, bool WG>
static int launch_bwd(BwdArgs &a, cudaStream_t s) {
  constexpr int NT = NT_RENDER;
  a.ra.rpc = NT / a.ra.S;
  const size_t smem = bwd_smem_bytes<STAGE, NT, WG>();
  ENS_CUDA_CALL(cudaFuncSetAttribute(render_bwd_kernel<STAGE, NT, WG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned g = (unsigned)((a.ra.R + a.ra.rpc - 1) / a.ra.rpc);
  render_bwd_kernel<STAGE, NT, WG><<<g, NT, smem, s>>>(a);
  ENS_CHECK_CUDA();
  return ENS_OK;
}


}  // namespace ens

using namespace ens;

// Kernel variant of the forward decode: "mma" (tensor pipe, default) or "fma" (ENS_FWD_VARIANT=fma; kept as
// the A/B reference the mma kernels were validated against).  Both are CUDA kernels for sm_100a.
static bool use_mma_forward() {
  const char *v = std::getenv("ENS_FWD_VARIANT");
  return !(v && std::strcmp(v, "fma") == 0);
}

static bool use_mma_backward() {
  const char *v = std::getenv("ENS_BWD_VARIANT");
  return !(v && std::strcmp(v, "fma") == 0);
}

extern "C" int ens_eval_points(const EnsScene *scene, int stage, const void *pts, int pts_is_f64, int64_t n,
                               int apply_bound_mask, float *out4, ens_stream_t stream) {
  int rc = check_scene(scene, stage);
  if (rc != ENS_OK) return rc;
  if (n < 0) return ENS_EINVAL;
  if (n == 0) return ENS_OK;                    // empty input: nothing to do (pointers of empty tensors are null)
  if (!pts || !out4) return ENS_EINVAL;
  const DevScene sc = make_dev_scene(scene);
  cudaStream_t s = (cudaStream_t)stream;
  {
    // variant of the forward-only decode: "tc" (tcgen05 / TMEM, default) | "mma" (mma.sync) | ENS_FWD_VARIANT=fma
    const char *v = std::getenv("ENS_EVAL_VARIANT");
    if (!(v && std::strcmp(v, "mma") == 0) && use_mma_forward()) {
      rc = tc_eval_points(sc, stage, pts, pts_is_f64, n, apply_bound_mask, out4, s);
      if (rc != ENS_EUNSUPPORTED) return rc;
    }
  }
  if (use_mma_forward()) {
    rc = mma_eval_points(sc, stage, pts, pts_is_f64, n, apply_bound_mask, out4, s);
    if (rc != ENS_EUNSUPPORTED) return rc;       // coarse stage: fma kernels below
  }
  switch (stage) {
    case ENS_STAGE_COARSE: return launch_eval<ENS_STAGE_COARSE>(sc, pts, pts_is_f64, n, apply_bound_mask, out4, s);
    case ENS_STAGE_MIDDLE: return launch_eval<ENS_STAGE_MIDDLE>(sc, pts, pts_is_f64, n, apply_bound_mask, out4, s);
    case ENS_STAGE_FINE: return launch_eval<ENS_STAGE_FINE>(sc, pts, pts_is_f64, n, apply_bound_mask, out4, s);
    default: return launch_eval<ENS_STAGE_COLOR>(sc, pts, pts_is_f64, n, apply_bound_mask, out4, s);
  }
}


extern "C" int ens_render_fwd(const EnsScene *scene, const EnsRenderCfg *cfg, int stage, const float *rays_o,
                              const float *rays_d, const float *gt_depth, const double *depth_max, int64_t n_rays,
                              double *depth, double *var, float *color, double *z_vals, float *weights, float *raw,
                              void *saved, int64_t saved_bytes, int saved_with_activations, void *scratch,
                              int64_t scratch_bytes, ens_stream_t stream) {
  int rc = check_scene(scene, stage);
  if (rc != ENS_OK) return rc;
  if (n_rays < 0) return ENS_EINVAL;
  if (n_rays == 0) return ENS_OK;               // empty batch: a no-op
  if (!rays_o || !rays_d || !depth || !var || !color) return ENS_EINVAL;
  const bool has_depth = gt_depth != nullptr && stage != ENS_STAGE_COARSE;
  if (has_depth && !depth_max) return ENS_EINVAL;
  int S, ns;
  rc = check_cfg(cfg, has_depth, stage, S, ns);
  if (rc != ENS_OK) return rc;
  if (n_rays == 0) return ENS_OK;
  FwdArgs a;
  a.sc = make_dev_scene(scene);
  fill_ray_args(a.ra, cfg, stage, rays_o, rays_d, gt_depth, depth_max, n_rays, S, ns);
  a.depth = depth; a.var = var; a.color = color; a.z_out = z_vals; a.w_out = weights; a.raw_out = raw;
  a.save_masks = nullptr; a.save_h = nullptr; a.n_tiles = 0;
  if (saved_with_activations < 0 || saved_with_activations > 3) return ENS_EINVAL;
  const bool want_h = saved_with_activations == 1;
  if (saved_with_activations == 3) {
    // tcgen05 forward that keeps the relu outputs r_0..r_4 of every decoder: the forward of the tcgen05 backward WITH
    // decoder gradients (mapping).  Anything that prevents this path is an error (no silent change of format).
    int64_t r_bytes = 0, nt32 = 0;
    const int64_t need = tc_saved3_bytes(n_rays, S, stage, &r_bytes, &nt32);
    if (need <= 0 || saved == nullptr || !use_mma_forward()) return ENS_EUNSUPPORTED;
    if (saved_bytes < need || (reinterpret_cast<uintptr_t>(saved) & 15)) return ENS_ESHAPE;
    a.save_r = reinterpret_cast<float *>(saved);
    a.save_masks = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(saved) + r_bytes);
    a.n_tiles = nt32;
    return tc_render_fwd(a, stage, scratch, scratch_bytes, (cudaStream_t)stream);
  }
  if (saved != nullptr) {
    int64_t n_tiles = 0, h_off = 0;
    const int64_t need = mma_fwd_saved_bytes(n_rays, S, stage, want_h, &n_tiles, &h_off);
    if (need > 0) {
      if (saved_bytes < need || (reinterpret_cast<uintptr_t>(saved) & 15)) return ENS_ESHAPE;
      a.save_masks = reinterpret_cast<uint32_t *>(saved);
      a.save_h = want_h ? reinterpret_cast<float *>(reinterpret_cast<char *>(saved) + h_off) : nullptr;
      a.n_tiles = n_tiles;
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (saved_with_activations == 2) {
    // masks as one word per point, written by the tcgen05 decode: the forward of a pose-only backward (tracking, event
    // render).  The caller tells ens_render_bwd the same kind, so there is no silent change of format: anything that
    // prevents this path is an error.
    if (a.save_masks == nullptr || !use_mma_forward() || !use_mma_backward()) return ENS_EUNSUPPORTED;
    return tc_render_fwd(a, stage, scratch, scratch_bytes, s);
  }
  if (use_mma_forward() && a.save_masks == nullptr && scratch != nullptr) {
    // no backward will follow and the caller gave scratch: placement + tcgen05 decode + compositing
    const char *v = std::getenv("ENS_EVAL_VARIANT");
    if (!(v && std::strcmp(v, "mma") == 0)) {
      rc = tc_render_fwd(a, stage, scratch, scratch_bytes, s);
      if (rc != ENS_EUNSUPPORTED) return rc;
    }
  }
  if (use_mma_forward()) {
    rc = mma_render_fwd(a, stage, s);
    if (rc != ENS_EUNSUPPORTED) return rc;       // coarse stage: fma kernels below
  }
  switch (stage) {
    case ENS_STAGE_COARSE: return launch_fwd<ENS_STAGE_COARSE>(a, s);
    case ENS_STAGE_MIDDLE: return launch_fwd<ENS_STAGE_MIDDLE>(a, s);
    case ENS_STAGE_FINE: return launch_fwd<ENS_STAGE_FINE>(a, s);
    default: return launch_fwd<ENS_STAGE_COLOR>(a, s);
  }
}

extern "C" int64_t ens_fwd_saved_bytes(int64_t n_rays, int n_samples_total, int stage, int want_decoder_grads) {
  if (!use_mma_forward() || !use_mma_backward()) return 0;
  return mma_fwd_saved_bytes(n_rays, n_samples_total, stage, want_decoder_grads, nullptr, nullptr);
}

extern "C" int64_t ens_fwd_saved_bytes_kind(int64_t n_rays, int n_samples_total, int stage, int kind) {
  if (kind == 3) return use_mma_forward() ? tc_saved3_bytes(n_rays, n_samples_total, stage, nullptr, nullptr) : 0;
  if (kind < 0 || kind > 3) return 0;
  return ens_fwd_saved_bytes(n_rays, n_samples_total, stage, kind == 1);
}

extern "C" int64_t ens_fwd_scratch_bytes(int64_t n_rays, int n_samples_total, int stage) {
  if (!use_mma_forward()) return 0;
  return tc_fwd_scratch_bytes(n_rays, n_samples_total, stage);
}

// tcgen05 backward (default for saved kinds 2 and 3); ENS_BWD_TC=0 keeps the mma.sync kernels for kind 2 (A/B check)
static bool use_tc_backward() {
  const char *v = std::getenv("ENS_BWD_TC");
  return !(v && v[0] == '0');
}

extern "C" int64_t ens_bwd_workspace_bytes(int64_t n_rays, int n_samples_total, int want_decoder_grads) {
  if (n_rays <= 0 || n_samples_total <= 0) return 0;
  const int64_t tc = tc_bwd_workspace_bytes(n_rays, n_samples_total, ENS_STAGE_COLOR, want_decoder_grads != 0);
  if (!want_decoder_grads) return tc;
  const int64_t fma = (n_rays * (int64_t)n_samples_total + 1) * 160 * (int64_t)sizeof(float);   // +1: dump row for idle lanes
  const int64_t mma = mma_bwd_workspace_bytes(n_rays, n_samples_total);
  const int64_t m = fma > mma ? fma : mma;
  return m > tc ? m : tc;
}

extern "C" int ens_render_bwd(const EnsScene *scene, const EnsRenderCfg *cfg, int stage, const float *rays_o,
                              const float *rays_d, const float *gt_depth, const double *depth_max, int64_t n_rays,
                              const float *raw, const double *g_depth, const double *g_var, const float *g_color,
                              const EnsGrads *grads, void *workspace, int64_t workspace_bytes, const void *saved,
                              int64_t saved_bytes, int saved_with_activations, ens_stream_t stream) {
  int rc = check_scene(scene, stage);
  if (rc != ENS_OK) return rc;
  if (n_rays < 0) return ENS_EINVAL;
  if (n_rays == 0) return ENS_OK;
  if (!rays_o || !rays_d || !raw || !grads) return ENS_EINVAL;
  const bool has_depth = gt_depth != nullptr && stage != ENS_STAGE_COARSE;
  if (has_depth && !depth_max) return ENS_EINVAL;
  int S, ns;
  rc = check_cfg(cfg, has_depth, stage, S, ns);
  if (rc != ENS_OK) return rc;
  if (n_rays == 0) return ENS_OK;
  BwdArgs a;
  a.sc = make_dev_scene(scene);
  fill_ray_args(a.ra, cfg, stage, rays_o, rays_d, gt_depth, depth_max, n_rays, S, ns);
  a.raw = raw; a.g_depth = g_depth; a.g_var = g_var; a.g_color = g_color;
  bool wg = false;
  // only the levels the stage touches receive gradient (decoder.py:312-342)
  const bool used[4][4] = {{true, false, false, false}, {false, true, false, false}, {false, true, true, false},
                           {false, true, true, true}};
  for (int l = 0; l < 4; ++l) {
    a.ggrid[l] = used[stage][l] ? grads->grid[l] : nullptr;
    a.gdec[l] = used[stage][l] ? grads->decoder[l] : nullptr;
  }
  // decoder grads are all-or-nothing per call for the levels the stage touches
  int n_dec = 0, n_used = 0;
  for (int l = 0; l < 4; ++l) if (used[stage][l]) { ++n_used; if (a.gdec[l]) ++n_dec; }
  if (n_dec != 0 && n_dec != n_used) return ENS_EINVAL;
  wg = n_dec > 0;
  a.g_rays_o = grads->rays_o; a.g_rays_d = grads->rays_d;
  a.hscratch = (float *)workspace;
  a.save_masks = nullptr; a.save_h = nullptr; a.n_tiles = 0; a.mask_fmt = 0;
  if (saved_with_activations < 0 || saved_with_activations > 3) return ENS_EINVAL;
  if (saved_with_activations == 3) {          // relu outputs kept by the tcgen05 forward: tcgen05 backward with decoder gradients
    int64_t r_bytes = 0, nt32 = 0;
    const int64_t need = tc_saved3_bytes(n_rays, S, stage, &r_bytes, &nt32);
    if (need <= 0 || saved == nullptr) return ENS_EUNSUPPORTED;
    if (saved_bytes < need) return ENS_ESHAPE;
    a.save_r = reinterpret_cast<const float *>(saved);
    a.save_masks = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(saved) + r_bytes);
    a.n_tiles = nt32;
    a.mask_fmt = 1;
    return tc_render_bwd(a, stage, wg, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  const bool have_h = saved_with_activations == 1;
  if (saved_with_activations == 2 && (saved == nullptr || !use_mma_backward())) return ENS_EUNSUPPORTED;
  if (saved != nullptr && use_mma_backward()) {
    int64_t n_tiles = 0, h_off = 0;
    const int64_t need = mma_fwd_saved_bytes(n_rays, S, stage, have_h, &n_tiles, &h_off);
    if (need > 0) {
      if (saved_bytes < need) return ENS_ESHAPE;
      a.save_masks = reinterpret_cast<const uint32_t *>(saved);
      a.save_h = have_h ? reinterpret_cast<const float *>(reinterpret_cast<const char *>(saved) + h_off) : nullptr;
      a.n_tiles = n_tiles;
      a.mask_fmt = saved_with_activations == 2 ? 1 : 0;
    } else if (saved_with_activations == 2) {
      return ENS_EUNSUPPORTED;
    }
  }
  const bool saved_covers = a.save_masks != nullptr && (!wg || a.save_h != nullptr);
  if (wg && !saved_covers) {
    if (!workspace || workspace_bytes < ens_bwd_workspace_bytes(n_rays, S, 1)) return ENS_ESHAPE;
  }
  // with saved activations the workspace is optional: given (and large enough) it enables the split backward
  if (wg && saved_covers && (!workspace || workspace_bytes < ens_bwd_workspace_bytes(n_rays, S, 1))) a.hscratch = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  if (saved_with_activations == 2 && !wg && use_tc_backward() && workspace != nullptr &&
      workspace_bytes >= tc_bwd_workspace_bytes(n_rays, S, stage, false)) {
    rc = tc_render_bwd(a, stage, false, workspace, workspace_bytes, s);
    if (rc != ENS_EUNSUPPORTED) return rc;
  }
  if (use_mma_backward()) {
    rc = mma_render_bwd(a, stage, wg, s);
    if (rc != ENS_EUNSUPPORTED) return rc;       // coarse stage: fma kernels below
  }
#define ENS_BWD_CASE(ST) \
  case ST: return wg ? launch_bwd<ST, true>(a, s) : launch_bwd<ST, false>(a, s);
  switch (stage) {
    ENS_BWD_CASE(ENS_STAGE_COARSE)
    ENS_BWD_CASE(ENS_STAGE_MIDDLE)
    ENS_BWD_CASE(ENS_STAGE_FINE)
    default: return wg ? launch_bwd<ENS_STAGE_COLOR, true>(a, s) : launch_bwd<ENS_STAGE_COLOR, false>(a, s);
  }
#undef ENS_BWD_CASE
}
