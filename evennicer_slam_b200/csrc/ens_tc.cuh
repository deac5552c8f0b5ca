// tcgen05 / UMMA / TMEM primitives shared by the tensor-core kernels (ens_decode_tc.cu, ens_bwd_tc.cu).
// Verified in isolation by tools/tc_probe.cu (K-major operands, TS form, 3xTF32) and tools/tc_probe2.cu / tc_probe3.cu
// (MN-major SWIZZLE_128B_BASE32B operands for the weight-gradient GEMMs, M = 64 lane map).
#pragma once
#include "ens_mma.cuh"

namespace ens {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: LBO = 128 B between the two K core matrices, SBO = (K/4)*128 B between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((128u >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128, K = 8, N from the instruction descriptor.  Executed by a
// whole (converged) warp with identical operands; elect.sync picks the one lane that issues, so the surrounding code
// stays warp-uniform (no per-instruction broadcast loops out of a divergent branch).
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE_%=;\n\t"
      "TC_WAIT_%=:\n\t"
      "nanosleep.u32 40;\n\t"                       // a spinning warp steals issue slots from the warp it shares the scheduler with
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE_%=;\n\t"
      "bra TC_WAIT_%=;\n\t"
      "TC_DONE_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
               "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// store 32 values as the A operand: the value itself (hi: the tensor core reads its top 19 bits) at `thi`, the TF32
// remainder at `tlo`
__device__ __forceinline__ void tmem_st32_split(uint32_t thi, uint32_t tlo, const float (&v)[32]) {
  uint32_t h[32], l[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    h[i] = __float_as_uint(v[i]);
    l[i] = __float_as_uint(v[i] - __uint_as_float(h[i] & 0xffffe000u));
  }
#define ENS_ST32(addr, r)                                                                                      \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                 \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "  \
               "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"                                \
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), \
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), \
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) \
               : "memory")
  ENS_ST32(thi, h);
  ENS_ST32(tlo, l);
#undef ENS_ST32
}
__device__ __forceinline__ void tmem_st_done() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// trilinear feature of one point into registers (same corner order / fma order as gather32 and gather_warp)
__device__ __forceinline__ void gather_regs(const float *__restrict__ grid, const int dims[3], const float pn[3],
                                            float (&f)[32]) {
  const Vox v = make_vox(pn, dims);
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int64_t lin; float w;
    corner(v, dims, c, lin, w);
    const float4 *src = reinterpret_cast<const float4 *>(grid + lin);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = __ldg(src + q);
      f[4 * q + 0] = fmaf(a.x, w, f[4 * q + 0]); f[4 * q + 1] = fmaf(a.y, w, f[4 * q + 1]);
      f[4 * q + 2] = fmaf(a.z, w, f[4 * q + 2]); f[4 * q + 3] = fmaf(a.w, w, f[4 * q + 3]);
    }
  }
}

// instruction descriptor: f32 accumulate, TF32 x TF32, both K-major, N columns, M = 128
__host__ __device__ constexpr uint32_t tc_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// issue  D (+)= A[128 x K] * W[32 x K]^T  in 3xTF32; A: hi at a_hi, lo at a_lo (TMEM columns), W: canonical smem
// matrix at float offset w_off (value) and w_off + TOT (remainder).  first_acc: 0 = overwrite D with the first MMA.
template <int K, int KMAT, int N = 32>
__device__ __forceinline__ void issue_gemm(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t sw_base, int w_off, int tot,
                                           uint32_t first_acc) {
  // K columns of a [N][KMAT] canonical matrix starting at float offset w_off (a column offset of 4c inside the
  // matrix is c*128 bytes and is folded into sw_base by the caller); SBO is that of the WHOLE matrix
  const uint64_t dh = umma_desc(sw_base + (uint32_t)w_off * 4u, (KMAT / 4) * 128u);
  const uint64_t dl = umma_desc(sw_base + (uint32_t)(w_off + tot) * 4u, (KMAT / 4) * 128u);
  constexpr uint32_t IDESC = tc_idesc(N);
  uint32_t acc = first_acc;
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) { umma_ts(d, a_lo + 8 * ks, dh + (uint64_t)(16 * ks), IDESC, acc); acc = 1; }
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) umma_ts(d, a_hi + 8 * ks, dl + (uint64_t)(16 * ks), IDESC, 1);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) umma_ts(d, a_hi + 8 * ks, dh + (uint64_t)(16 * ks), IDESC, 1);
}


// ---- SS form with MN-major operands (K = the POINT dimension): weight gradients ---------------------------------------
// For 32-bit operands the only MN-major shared-memory layout is SWIZZLE_128B_BASE32B (descriptor layout type 1):
// rows of 128 bytes (32 consecutive M/N elements) per k, 32-byte chunks XOR-ed with (address bits 7..8) = (k mod 4) when the
// tile base is 512-byte aligned IN THE SHARED ADDRESS SPACE (the hardware swizzles absolute address bits):
//     element (mn, k) at  (mn / 32) * LBO + k * 128 + (((mn % 32) / 8) ^ (k % 4)) * 32 + (mn % 8) * 4   bytes
// SBO = 512 B (next 4 k), LBO = stride between 32-element M/N blocks; one instruction covers 8 k = 1024 B.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;          // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// instruction descriptor, both operands MN-major
__host__ __device__ constexpr uint32_t tc_idesc_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// float offset of element (row = point k, column mn < 32) inside one [128][32] MN-major block (base 512-byte aligned)
__host__ __device__ constexpr int mn_off(int k, int mn) { return k * 32 + ((((mn >> 3) ^ (k & 3)) << 3) | (mn & 7)); }

}  // namespace ens
