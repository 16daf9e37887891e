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

// LayerNorm over the last dim, fp32 statistics (two-pass on registers), affine, output rounded once.
// Optional `add` row-broadcast term (positional embeddings added after the norm, e.g.
// src/models/qwen_visual.py:146-148) with period add_rows: y = LN(x)*g+b (+ add[row % add_rows]).
template <typename T, int VEC_PER_THREAD>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ g,
                                                        const T* __restrict__ b, T* __restrict__ y, int ldy, int K,
                                                        float eps, const T* __restrict__ add, int add_rows,
                                                        T* __restrict__ y2, int ldy2) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const T* xr = x + (size_t)row * ldx;
  const int nvec = K >> 3;
  float f[VEC_PER_THREAD][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      vec8 v = ld_cached16(xr + vi * 8);
      unpack8<T>(v, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  const float mean = block_sum(s, red) / (float)K;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = __frsqrt_rn(block_sum(q, red) / (float)K + eps);
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      float gg[8], bb[8], o[8];
      unpack8<T>(ld_cached16(g + vi * 8), gg);
      if (b) unpack8<T>(ld_cached16(b + vi * 8), bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * gg[j] + (b ? bb[j] : 0.f);
      // round once to the storage type, as torch's LayerNorm kernel does
      vec8 ov = pack8<T>(o);
      st16(y + (size_t)row * ldy + vi * 8, ov);
      if (add) {
        // second output: LN(x) + add, both operands already in storage precision (an fp16 add)
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

SS_API int ss_layernorm(int dtype, const void* x, int ldx, const void* gamma, const void* beta, void* y, int ldy,
                        int rows, int K, float eps, const void* add, int add_rows, void* y2, int ldy2, void* stream) {
  SS_REQUIRE(K % 8 == 0 && K <= 256 * 8 * 2, "K must be a multiple of 8 and <= 4096");
  SS_REQUIRE(dtype == SS_F16 || dtype == SS_BF16, "dtype");
  SS_REQUIRE(add == nullptr || (y2 != nullptr && add_rows > 0), "add needs y2/add_rows");
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SS_F16)
    layernorm_kernel<__half, 2><<<rows, 256, 0, s>>>((const __half*)x, ldx, (const __half*)gamma, (const __half*)beta,
                                                     (__half*)y, ldy, K, eps, (const __half*)add, add_rows,
                                                     (__half*)y2, ldy2);
  else
    layernorm_kernel<__nv_bfloat16, 2><<<rows, 256, 0, s>>>(
        (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, (__nv_bfloat16*)y, ldy,
        K, eps, (const __nv_bfloat16*)add, add_rows, (__nv_bfloat16*)y2, ldy2);
  SS_LAUNCH_CHECK();
  return 0;
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
