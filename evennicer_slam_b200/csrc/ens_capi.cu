// Version / error strings / size queries of the C ABI (include/ens_render.h).
#include <cstdio>
#include "ens_common.cuh"

namespace ens {
static thread_local char g_last_error[256] = "";
void note_cuda_error(cudaError_t e, const char *file, int line) {
  std::snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s:%d)", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace ens

extern "C" const char *ens_last_error(void) { return ens::g_last_error; }

extern "C" int ens_version(void) { return ENS_ABI_VERSION; }

extern "C" const char *ens_strerror(int code) {
  switch (code) {
    case ENS_OK: return "ok";
    case ENS_EINVAL: return "invalid argument (null pointer, bad enum or negative size)";
    case ENS_ESHAPE: return "inconsistent shape or size";
    case ENS_ECUDA: return "CUDA error (ens_last_error() has the runtime's message)";
    case ENS_ENCCL: return "collective error";
    case ENS_EUNSUPPORTED: return "unsupported rendering mode (N_importance>0, occupancy=False, perturb>0 or lindisp)";
    default: return "unknown error";
  }
}

extern "C" int64_t ens_packed_decoder_floats(int level) {
  if (level < 0 || level > 3) return -1;
  return ens::packed_floats(level);
}
extern "C" int64_t ens_decoder_grad_floats(int level) {
  if (level < 0 || level > 3) return -1;
  return ens::grad_floats(level);
}
extern "C" int ens_decoder_num_tensors(int level) {
  if (level < 0 || level > 3) return -1;
  return level == ENS_LEVEL_COARSE ? 12 : 23;
}
