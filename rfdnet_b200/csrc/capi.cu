// capi.cu -- library-level entry points and error plumbing of librfdnet_b200.so.
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace rfd {
thread_local char g_last_error[512] = {0};
std::atomic<long long> g_launch_count{0};

int set_cuda_error(cudaError_t e, const char *where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  (void)cudaGetLastError();  // clear the sticky-less error so later calls can proceed
  return RFD_ERR_CUDA;
}
}  // namespace rfd

extern "C" int rfd_abi_version(void) { return RFD_ABI_VERSION; }

extern "C" const char *rfd_status_string(int status) {
  switch (status) {
    case RFD_OK: return "ok";
    case RFD_ERR_INVALID_ARGUMENT: return "invalid argument (null pointer, negative size or unsupported combination)";
    case RFD_ERR_UNSUPPORTED_SIZE: return "size outside the range the sm_100a kernels support";
    case RFD_ERR_CUDA: return "CUDA runtime / launch failure (see rfd_last_error())";
    case RFD_ERR_NO_DEVICE: return "no compute-capability 10.x device";
    default: return "unknown status";
  }
}

extern "C" const char *rfd_last_error(void) { return rfd::g_last_error; }

extern "C" long long rfd_launch_count(void) { return rfd::g_launch_count.load(); }

extern "C" int rfd_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { rfd::set_cuda_error(e, "rfd_device_info"); return RFD_ERR_NO_DEVICE; }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) { rfd::set_cuda_error(e, "rfd_device_info"); return RFD_ERR_NO_DEVICE; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return p.major == 10 ? RFD_OK : RFD_ERR_NO_DEVICE;
}
