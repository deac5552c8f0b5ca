// tcgen05 probe (sm_100a): D[128x32] = A[128x32] * W[32x32]^T on the 5th-gen tensor core, TF32,
//   test 1: A and W from shared memory (K-major, no swizzle, canonical 8x16B core matrices), D in TMEM
//   test 2: A written to TMEM with tcgen05.st and used as the A operand (TS form)
//   test 3: 3xTF32 (hi/lo split of both operands, three accumulating passes) against an fp64 reference
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: element (row r, col k) of an [R x K] fp32 operand lives at
//   (r/8) * SBO + (k/4) * 128 + (r%8) * 16 + (k%4) * 4      with SBO = (K/4) * 128 bytes
__host__ __device__ inline int canon_off_floats(int r, int k, int K) {
  return (r / 8) * (K / 4) * 32 + (k / 4) * 32 + (r % 8) * 4 + (k % 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;          // version = 1 (Blackwell)
  return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

#define TMEM_LD32(r, taddr)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "  \
               "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                        \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
               : "r"(taddr) : "memory")
#define TMEM_ST32(taddr, r)                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, " \
               "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"                               \
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), \
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), \
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) \
               : "memory")

constexpr int M = 128, N = 32, K = 32;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

// mode 0: SS single pass; mode 1: TS single pass; mode 2: TS 3xTF32
__global__ void __launch_bounds__(128) probe_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                    float *__restrict__ D, int mode, long long *cycles) {
  __shared__ __align__(128) float sA[M * K];        // canonical layout
  __shared__ __align__(128) float sWh[N * K];       // W (hi = as is: the tensor core drops the low 13 bits)
  __shared__ __align__(128) float sWl[N * K];       // W - trunc_tf32(W)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M * K; i += 128) { const int r = i / K, k = i % K; sA[canon_off_floats(r, k, K)] = A[i]; }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const float w = W[i];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    sWh[canon_off_floats(r, k, K)] = w;
    sWl[canon_off_floats(r, k, K)] = w - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> async proxy (UMMA reads)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
  // TMEM columns: [0,32) accumulator D, [32,64) A_hi, [64,96) A_lo
  const uint32_t tD = tbase, tAh = tbase + 32, tAl = tbase + 64;

  if (mode >= 1) {   // each thread = one row of A: write its 32 K-values (and the tf32 remainder) into TMEM
    uint32_t rh[32], rl[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a = A[tid * K + k];
      const float hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
      rh[k] = __float_as_uint(a);
      rl[k] = __float_as_uint(a - hi);
    }
    TMEM_ST32(tAh + lane_off, rh);
    TMEM_ST32(tAl + lane_off, rl);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  long long t0 = clock64();
  if (tid == 0) {
    const uint32_t SBO = (K / 4) * 128, LBO = 128;
    const uint64_t dA = make_desc(smem_u32(sA), LBO, SBO);
    const uint64_t dWh = make_desc(smem_u32(sWh), LBO, SBO);
    const uint64_t dWl = make_desc(smem_u32(sWl), LBO, SBO);
    uint32_t acc = 0;
    if (mode == 0) {
      for (int ks = 0; ks < K / 8; ++ks) { mma_ss(tD, dA + (uint64_t)(ks * 16), dWh + (uint64_t)(ks * 16), IDESC, acc); acc = 1; }
    } else if (mode == 1) {
      for (int ks = 0; ks < K / 8; ++ks) { mma_ts(tD, tAh + ks * 8, dWh + (uint64_t)(ks * 16), IDESC, acc); acc = 1; }
    } else {
      for (int ks = 0; ks < K / 8; ++ks) { mma_ts(tD, tAl + ks * 8, dWh + (uint64_t)(ks * 16), IDESC, acc); acc = 1; }
      for (int ks = 0; ks < K / 8; ++ks) mma_ts(tD, tAh + ks * 8, dWl + (uint64_t)(ks * 16), IDESC, 1);
      for (int ks = 0; ks < K / 8; ++ks) mma_ts(tD, tAh + ks * 8, dWh + (uint64_t)(ks * 16), IDESC, 1);
    }
    commit(&bar);
  }
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t1 = clock64();
  uint32_t r[32];
  TMEM_LD32(r, tD + lane_off);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int n = 0; n < 32; ++n) D[tid * N + n] = __uint_as_float(r[n]);
  if (tid == 0 && cycles) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(128u) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
  std::vector<float> A(M * K), W(N * K), D(M * N);
  srand(1);
  for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto &v : W) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dW, *dD; long long *dc;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    probe_kernel<<<1, 128>>>(dA, dW, dD, mode, dc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double e_tf = 0, e_full = 0, mx = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s_tf = 0, s_full = 0;
        for (int k = 0; k < K; ++k) {
          s_tf += (double)trunc_tf32(A[m * K + k]) * (double)trunc_tf32(W[n * K + k]);
          s_full += (double)A[m * K + k] * (double)W[n * K + k];
        }
        e_tf = fmax(e_tf, fabs(D[m * N + n] - s_tf));
        e_full = fmax(e_full, fabs(D[m * N + n] - s_full));
        mx = fmax(mx, fabs(s_full));
      }
    printf("mode %d: max|D - tf32_ref| = %.3e   max|D - fp64_ref| = %.3e   (max|ref| %.2f)   mma+commit+wait cycles %lld\n", mode, e_tf, e_full, mx, cyc);
  }
  return 0;
}
