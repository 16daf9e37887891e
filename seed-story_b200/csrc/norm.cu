// Row normalisations: RMSNorm (Llama), LayerNorm (ViT / resamplers / UNet transformer blocks),
// L2-normalise over the token axis (ResamplerXLV2 input).
//
// RMSNorm follows the reference's rounding chain exactly
// (src/models_clm/modeling_llama_xformer.py:107-115): fp32 mean-square, x*rsqrt in fp32, round to
// fp16, then weight*x as an fp16 multiply.
#include "common.cuh"

template <int VEC_PER_THREAD>
__global__ void __launch_bounds__(256) rmsnorm_f16_kernel(const __half* __restrict__ x, int ldx,
                                                          const __half* __restrict__ w, __half* __restrict__ y,
                                                          int ldy, int K, float eps) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  const __half* xr = x + (size_t)row * ldx;
  __half* yr = y + (size_t)row * ldy;
  const int nvec = K >> 3;
  vec8 v[VEC_PER_THREAD];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      v[i] = ld_cached16(xr + vi * 8);
      float f[8];
      unpack8<__half>(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = __frsqrt_rn(ss / (float)K + eps);
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      float f[8];
      unpack8<__half>(v[i], f);
      vec8 wv = ld_cached16(w + vi * 8);
      const __half* wh = reinterpret_cast<const __half*>(&wv);
      vec8 o;
      __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) oh[j] = __hmul(wh[j], __float2half_rn(f[j] * rstd));
      st16(yr + vi * 8, o);
    }
  }
}

SS_API int ss_rmsnorm_f16(const void* x, int ldx, const void* weight, void* y, int ldy, int rows, int K, float eps,
                          void* stream) {
  SS_REQUIRE(K % 8 == 0 && K <= 256 * 8 * 6, "K must be a multiple of 8 and <= 12288");
  SS_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0, "row pitch must be 16-byte aligned");
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int nvec = K / 8;
  if (nvec <= 256 * 2)
    SS_CUDA(ss::launch_pdl(rmsnorm_f16_kernel<2>, dim3(rows), dim3(256), 0, s, (const __half*)x, ldx,
                           (const __half*)weight, (__half*)y, ldy, K, eps));
  else
    SS_CUDA(ss::launch_pdl(rmsnorm_f16_kernel<6>, dim3(rows), dim3(256), 0, s, (const __half*)x, ldx,
                           (const __half*)weight, (__half*)y, ldy, K, eps));
  SS_LAUNCH_CHECK();
  return 0;
}

// LayerNorm over the last dim: ONE WARP PER ROW (8 rows per CTA), the row lives in registers, fp32 statistics
// (mean, then centred second moment), affine, output rounded once.  Optional second output
// y2 = LN(x) + add[row % add_rows] (positional-embedding adds, e.g. src/models/qwen_visual.py:146-148).
template <typename T, int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ g,
                                                        const T* __restrict__ b, T* __restrict__ y, int ldy, int rows,
                                                        int K, float eps, const T* __restrict__ add, int add_rows,
                                                        T* __restrict__ y2, int ldy2) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const T* xr = x + (size_t)row * ldx;
  const int nvec = K >> 3;
  float f[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      unpack8<T>(ld_cached16(xr + vi * 8), f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)K;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = __frsqrt_rn(warp_sum(q) / (float)K + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      float gg[8], bb[8], o[8];
      unpack8<T>(ld_cached16(g + vi * 8), gg);
      if (b) unpack8<T>(ld_cached16(b + vi * 8), bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * gg[j] + (b ? bb[j] : 0.f);
      const vec8 ov = pack8<T>(o);  // rounded once to the storage type, as torch's LayerNorm kernel does
      st16(y + (size_t)row * ldy + vi * 8, ov);
      if (add) {
        float a[8], r[8];
        unpack8<T>(ld_cached16(add + (size_t)(row % add_rows) * K + vi * 8), a);
        unpack8<T>(ov, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[j];
        st16(y2 + (size_t)row * ldy2 + vi * 8, pack8<T>(r));
      }
    }
  }
}

template <typename T>
static int launch_layernorm(const void* x, int ldx, const void* gamma, const void* beta, void* y, int ldy, int rows, int K,
                            float eps, const void* add, int add_rows, void* y2, int ldy2, cudaStream_t s) {
  const dim3 grid((rows + 7) / 8), block(256);
  const int nvec = K / 8;
#define SS_LN_LAUNCH(MAXV)                                                                                          \
  SS_CUDA(ss::launch_pdl(layernorm_kernel<T, MAXV>, grid, block, 0, s, (const T*)x, ldx, (const T*)gamma, (const T*)beta, \
                         (T*)y, ldy, rows, K, eps, (const T*)add, add_rows, (T*)y2, ldy2))
  if (nvec <= 4 * 32) {
    SS_LN_LAUNCH(4);
  } else if (nvec <= 8 * 32) {
    SS_LN_LAUNCH(8);
  } else {
    SS_LN_LAUNCH(16);
  }
#undef SS_LN_LAUNCH
  return 0;
}

SS_API int ss_layernorm(int dtype, const void* x, int ldx, const void* gamma, const void* beta, void* y, int ldy,
                        int rows, int K, float eps, const void* add, int add_rows, void* y2, int ldy2, void* stream) {
  SS_REQUIRE(K % 8 == 0 && K <= 4096, "K must be a multiple of 8 and <= 4096");
  SS_REQUIRE(dtype == SS_F16 || dtype == SS_BF16, "dtype");
  SS_REQUIRE(add == nullptr || (y2 != nullptr && add_rows > 0), "add needs y2/add_rows");
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SS_F16)
    return launch_layernorm<__half>(x, ldx, gamma, beta, y, ldy, rows, K, eps, add, add_rows, y2, ldy2, s);
  return launch_layernorm<__nv_bfloat16>(x, ldx, gamma, beta, y, ldy, rows, K, eps, add, add_rows, y2, ldy2, s);
}

// F.normalize(x) with x [B, T, C]: L2 norm over dim=1 (the TOKEN axis — src/models_ipa/resampler.py:269),
// eps 1e-12: y = x / max(||x[:, :, c]||_2, eps).
__global__ void l2norm_tokens_kernel(const __half* __restrict__ x, __half* __restrict__ y, int T, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const __half* xb = x + (size_t)b * T * C;
  float s = 0.f;
  for (int t = 0; t < T; ++t) {
    const float v = __half2float(xb[(size_t)t * C + c]);
    s += v * v;
  }
  // torch computes the norm in the tensor dtype's accumulate type (fp32) and rounds it to fp16 before dividing
  const float nrm = fmaxf(__half2float(__float2half_rn(sqrtf(s))), 1e-12f);
  __half* yb = y + (size_t)b * T * C;
  for (int t = 0; t < T; ++t) yb[(size_t)t * C + c] = __float2half_rn(__half2float(xb[(size_t)t * C + c]) / nrm);
}

SS_API int ss_l2norm_tokens_f16(const void* x, void* y, int B, int T, int C, void* stream) {
  if (B == 0) return 0;
  dim3 grid(ceil_div(C, 128), B);
  l2norm_tokens_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, T, C);
  SS_LAUNCH_CHECK();
  return 0;
}
