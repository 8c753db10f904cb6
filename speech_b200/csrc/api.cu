// Library introspection entry points of the C ABI (include/speech_b200.h).
#include "common.cuh"

#include "../../include/speech_b200.h"

extern "C" int sb_version(void) { return 100; }

extern "C" const char* sb_status_string(int status) {
  switch (status) {
    case SB_OK: return "ok";
    case SB_ERR_INVALID: return "invalid argument";
    case SB_ERR_CUDA: return "CUDA call failed";
    case SB_ERR_UNSUPPORTED: return "unsupported shape";
    case SB_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

extern "C" int sb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return SB_ERR_CUDA;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
  return SB_OK;
}
