// Shared device/host helpers for the seedstory_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

// ---------------------------------------------------------------------------------------------
// Error convention of the C-ABI: every entry point returns 0 on success, non-zero on failure and
// leaves a thread-local message readable through ss_last_error().
// ---------------------------------------------------------------------------------------------
namespace ss {
void set_error(const std::string& msg);
int fail(const char* file, int line, const std::string& msg);
}  // namespace ss

#define SS_FAIL(msg) return ss::fail(__FILE__, __LINE__, (msg))
#define SS_REQUIRE(cond, msg)                         \
  do {                                                \
    if (!(cond)) return ss::fail(__FILE__, __LINE__, std::string("requirement failed: " #cond " — ") + (msg)); \
  } while (0)
#define SS_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return ss::fail(__FILE__, __LINE__, std::string(#expr ": ") + cudaGetErrorString(_e)); \
  } while (0)
#define SS_LAUNCH_CHECK() SS_CUDA(cudaGetLastError())

#define SS_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------
// dtype tags used across the C-ABI (matches include/seedstory_b200.h)
// ---------------------------------------------------------------------------------------------
enum { SS_F16 = 0, SS_BF16 = 1, SS_F32 = 2 };

template <typename T> struct ss_num;
template <> struct ss_num<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct ss_num<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of 8 x 16-bit values
struct __align__(16) vec8 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ vec8 ld_stream16(const void* p) {  // weights: read once, keep out of L1
  vec8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ vec8 ld_stream_rw16(const void* p) {
  // streamed data that an earlier kernel on the stream wrote (KV-cache pages): coherent load (no .nc — under
  // programmatic dependent launch the producer may still be running when this kernel starts), kept out of L1
  vec8 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ vec8 ld_cached16(const void* p) {
  vec8 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st16(void* p, const vec8& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T>
__device__ __forceinline__ void unpack8(const vec8& v, float* f) {
  const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = ss_num<T>::to_f(h[i]);
}
template <typename T>
__device__ __forceinline__ vec8 pack8(const float* f) {
  vec8 v;
  T* h = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = ss_num<T>::from_f(f[i]);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum over up to 1024 threads; `red` must hold 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below fp16/bf16 output resolution): two MUFU ops and a
// short FMA chain instead of erff's branchy polynomial — the GELU epilogues are instruction-bound otherwise.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  float t;  // approximate reciprocal (1 MUFU, ~1 ulp): the IEEE-rounded __frcp_rn costs a Newton sequence per element
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float r = 1.f - poly * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erf_fast(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with ss::launch_pdl may start while its predecessor
// in the stream is still draining; it must call pdl_wait() before touching anything the predecessor wrote
// (weights are constants and may be prefetched first) and should call pdl_trigger() as early as possible.
// Both are no-ops for kernels launched the ordinary way.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

namespace ss {
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace ss
