// ABI bookkeeping: version, error strings, device check.
#include "common.cuh"

extern "C" int pemp_abi_version(void) { return PEMP_ABI_VERSION; }

extern "C" const char* pemp_strerror(int code) {
  switch (code) {
    case PEMP_OK: return "ok";
    case PEMP_E_SHAPE: return "PEMP_E_SHAPE: a dimension is non-positive or outside the supported range";
    case PEMP_E_ALIGN: return "PEMP_E_ALIGN: pointer alignment requirement not met";
    case PEMP_E_WORKSPACE: return "PEMP_E_WORKSPACE: workspace missing or too small";
    case PEMP_E_ARCH: return "PEMP_E_ARCH: current device is not compute capability 10.x (B200)";
    case PEMP_E_NULL: return "PEMP_E_NULL: required pointer is NULL";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown pemp error code";
}

extern "C" int pemp_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  return major == 10 ? PEMP_OK : PEMP_E_ARCH;
}
