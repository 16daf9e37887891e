// Library-level C-ABI: error state, version, device probe.
#include "common.cuh"

namespace ss {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const char* file, int line, const std::string& msg) {
  g_last_error = std::string(file) + ":" + std::to_string(line) + ": " + msg;
  return 1;
}
}  // namespace ss

SS_API const char* ss_last_error(void) { return ss::g_last_error.c_str(); }

SS_API int ss_version(void) { return 100; }  // 0.1.0

// Fails (non-zero) unless a compute-capability-10.x device is visible: there is no CPU fallback.
SS_API int ss_require_device(int* sm_count_out) {
  int n = 0;
  SS_CUDA(cudaGetDeviceCount(&n));
  SS_REQUIRE(n > 0, "no CUDA device visible; seedstory_b200 has no CPU fallback");
  int dev = 0;
  SS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SS_CUDA(cudaGetDeviceProperties(&p, dev));
  SS_REQUIRE(p.major == 10, "kernels are built for sm_100a only");
  if (sm_count_out) *sm_count_out = p.multiProcessorCount;
  return 0;
}

SS_API int ss_stream_sync(void* stream) {
  SS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
