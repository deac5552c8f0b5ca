// tcgen05 probe 3 (sm_100a): decode which shared-memory words ONE tcgen05.mma (kind::tf32, SS form) reads for a given
// pair of matrix descriptors, by one-hot tests:  A one-hot + B all ones -> the row (TMEM lane) a word belongs to;
// A all ones + B one-hot -> the column; A one-hot + B one-hot -> whether the two words share a k index.
// usage: tc_probe3 a_major b_major layout_type lbo sbo M N        (bytes; the same descriptor fields for A and B)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_LOOP_%=;\n\t"
      "DONE_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
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

constexpr int NA = 16384, NB = 4096;       // floats of the A region (64 KB) and the B region (16 KB)

struct Params { int a_major, b_major, type, lbo, sbo, M, N; };

// tests: (ia, ja) word indices of the one-hot elements (-1 = that operand is all ones).  out[t] = {count, lane, col}
__global__ void __launch_bounds__(128) onehot_kernel(Params p, const int *__restrict__ ia, const int *__restrict__ ja,
                                                     int ntests, int a_all, int b_all, int *__restrict__ out) {
  extern __shared__ __align__(128) float smem_raw[];
  float *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;   // 1024-aligned in the SHARED address space (the swizzle uses absolute address bits)
  float *sX = smem, *sG = smem + NA;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_cnt, s_lane, s_col;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < NA; i += 128) sX[i] = a_all ? 1.f : 0.f;
  for (int i = tid; i < NB; i += 128) sG[i] = b_all ? 1.f : 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_major << 15) | ((uint32_t)p.b_major << 16) |
                         ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
  uint32_t parity = 0;
  for (int t = 0; t < ntests; ++t) {
    {   // zero D so lanes an M = 64 instruction does not write read as 0
      uint32_t z[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) z[k] = 0u;
      TMEM_ST32(tbase + lane_off, z);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
      s_cnt = 0; s_lane = -1; s_col = -1;
      if (!a_all) sX[ia[t]] = 1.f;
      if (!b_all) sG[ja[t]] = 1.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mma_ss(tbase, make_desc(smem_u32(sX), p.lbo, p.sbo, p.type), make_desc(smem_u32(sG), p.lbo, p.sbo, p.type), idesc, 0);
      commit(&bar);
    }
    wait_bar(&bar, parity); parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    TMEM_LD32(r, tbase + lane_off);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int c = 0, first = -1;
#pragma unroll
    for (int n = 0; n < 32; ++n) if (n < p.N && __uint_as_float(r[n]) != 0.f) { ++c; if (first < 0) first = n; }
    if (c) { atomicAdd(&s_cnt, c); atomicMax(&s_lane, tid); atomicMax(&s_col, first); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      out[3 * t] = s_cnt; out[3 * t + 1] = s_lane; out[3 * t + 2] = s_col;
      if (!a_all) sX[ia[t]] = 0.f;
      if (!b_all) sG[ja[t]] = 0.f;
    }
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(64u) : "memory");
}

__global__ void __launch_bounds__(128) numeric_kernel(Params p, const float *__restrict__ xa, const float *__restrict__ xb,
                                                      float *__restrict__ D) {
  extern __shared__ __align__(128) float smem_raw[];
  float *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;   // 1024-aligned in the SHARED address space (the swizzle uses absolute address bits)
  float *sX = smem, *sG = smem + NA;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < NA; i += 128) sX[i] = xa[i];
  for (int i = tid; i < NB; i += 128) sG[i] = xb[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_major << 15) | ((uint32_t)p.b_major << 16) |
                         ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
  if (tid == 0) {
    mma_ss(tbase, make_desc(smem_u32(sX), p.lbo, p.sbo, p.type), make_desc(smem_u32(sG), p.lbo, p.sbo, p.type), idesc, 0);
    commit(&bar);
  }
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  TMEM_LD32(r, tbase + lane_off);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int n = 0; n < 32; ++n) D[tid * 32 + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(64u) : "memory");
}

static std::vector<int> run(const Params &p, const std::vector<int> &ia, const std::vector<int> &ja, int a_all, int b_all) {
  const int n = (int)std::max(ia.size(), ja.size());
  int *dia, *dja, *dout;
  CK(cudaMalloc(&dia, n * 4 + 4)); CK(cudaMalloc(&dja, n * 4 + 4)); CK(cudaMalloc(&dout, n * 12));
  if (!ia.empty()) CK(cudaMemcpy(dia, ia.data(), ia.size() * 4, cudaMemcpyHostToDevice));
  if (!ja.empty()) CK(cudaMemcpy(dja, ja.data(), ja.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(NA + NB) * 4 + 1024;
  CK(cudaFuncSetAttribute(onehot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  onehot_kernel<<<1, 128, smem>>>(p, dia, dja, n, a_all, b_all, dout);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int> out(3 * n);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(dia); cudaFree(dja); cudaFree(dout);
  return out;
}

int main(int argc, char **argv) {
  if (argc < 8) { printf("usage: %s a_major b_major type lbo sbo M N\n", argv[0]); return 1; }
  Params p{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7])};
  printf("== a_major %d b_major %d type %d lbo %d sbo %d M %d N %d\n", p.a_major, p.b_major, p.type, p.lbo, p.sbo, p.M, p.N);
  // phase 1: A word -> lane
  std::vector<int> ia(NA), none;
  for (int i = 0; i < NA; ++i) ia[i] = i;
  std::vector<int> o1 = run(p, ia, none, 0, 1);
  std::vector<std::vector<int>> rows(128);
  int used = 0;
  for (int i = 0; i < NA; ++i) if (o1[3 * i] > 0) { rows[o1[3 * i + 1]].push_back(i); ++used; }
  printf("A: %d words read\n", used);
  for (int m = 0; m < 128; ++m) {
    if (rows[m].empty()) continue;
    if (m < 10 || m % 16 == 0 || m == 31 || m == 32 || m == 33 || m == 63 || m == 65 || m == 127) {
      printf("  lane %3d:", m);
      for (int w : rows[m]) printf(" %d", w);
      printf("\n");
    }
  }
  // phase 2: B word -> column
  std::vector<int> jb(NB);
  for (int j = 0; j < NB; ++j) jb[j] = j;
  std::vector<int> o2 = run(p, none, jb, 1, 0);
  std::vector<std::vector<int>> cols(32);
  used = 0;
  for (int j = 0; j < NB; ++j) if (o2[3 * j] > 0) { cols[o2[3 * j + 2]].push_back(j); ++used; }
  printf("B: %d words read\n", used);
  for (int n = 0; n < p.N; ++n) {
    if (n < 10 || n == 16 || n == 31) {
      printf("  col %3d:", n);
      for (int w : cols[n]) printf(" %d", w);
      printf("\n");
    }
  }
  // phase 3: k pairing between the words of lane m0 (first lane with words) and column 0
  int m0 = 0; while (m0 < 128 && rows[m0].empty()) ++m0;
  if (m0 < 128 && !cols[0].empty()) {
    std::vector<int> pa, pb;
    for (int a : rows[m0]) for (int b : cols[0]) { pa.push_back(a); pb.push_back(b); }
    std::vector<int> o3 = run(p, pa, pb, 0, 0);
    printf("k pairing (lane %d word, col 0 word):", m0);
    for (size_t t = 0; t < pa.size(); ++t) if (o3[3 * t] > 0) printf(" (%d,%d)", pa[t], pb[t]);
    printf("\n");
  }
  {   // phase 4: numeric check of the decoded map with small-integer data
    std::vector<float> xa(NA), xb(NB), D(128 * 32);
    srand(7);
    for (auto &v : xa) v = (float)(rand() % 7 - 3);
    for (auto &v : xb) v = (float)(rand() % 7 - 3);
    float *dxa, *dxb, *dD;
    CK(cudaMalloc(&dxa, NA * 4)); CK(cudaMalloc(&dxb, NB * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dxa, xa.data(), NA * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dxb, xb.data(), NB * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(NA + NB) * 4 + 1024;
    CK(cudaFuncSetAttribute(numeric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    numeric_kernel<<<1, 128, smem>>>(p, dxa, dxb, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0, tot = 0;
    for (int m = 0; m < 128; ++m) {
      if (rows[m].size() != 8) continue;
      for (int n = 0; n < p.N; ++n) {
        if (cols[n].size() != 8) continue;
        float e = 0;
        for (int k = 0; k < 8; ++k) e += xa[rows[m][k]] * xb[cols[n][k]];
        ++tot;
        if (e != D[m * 32 + n]) { if (bad < 5) printf("  mismatch lane %d col %d: got %g expect %g\n", m, n, D[m * 32 + n], e); ++bad; }
      }
    }
    printf("numeric check (pairing by sorted word order): %d / %d mismatches\n", bad, tot);
  }
  return 0;
}
