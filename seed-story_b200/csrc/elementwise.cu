// Bandwidth-bound glue kernels of the visual / de-tokenizer path: patchify, broadcast adds, row scatter,
// GroupNorm(+SiLU) on NHWC, nearest 2x upsample, channel concat, stride-2 im2col, CFG + Euler update,
// softmax rows, transposes, dtype casts, image post-processing.  All are vectorised 16-byte accesses
// over NHWC / token-major tensors (C innermost), grid-stride over 148*k CTAs.
#include "common.cuh"

namespace {
constexpr int EW_THREADS = 256;
inline int ew_grid(long long work_items) {
  long long g = (work_items + EW_THREADS - 1) / EW_THREADS;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace

// ---- ViT patchify: NCHW image -> [B*G*G, Kpad] rows, k = c*P*P + ky*P + kx (conv1 weight flatten order,
// src/models/qwen_visual.py:347,382) --------------------------------------------------------------
__global__ void im2col_patch_kernel(const __half* __restrict__ img, __half* __restrict__ out, int B, int C, int S,
                                    int P, int Kpad) {
  const int G = S / P;
  const long long total = (long long)B * G * G * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const long long row = i / Kpad;
    __half v = __float2half_rn(0.f);
    if (k < C * P * P) {
      const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
      const int gx = (int)(row % G), gy = (int)((row / G) % G), b = (int)(row / ((long long)G * G));
      v = img[(((long long)b * C + c) * S + gy * P + ky) * S + gx * P + kx];
    }
    out[i] = v;
  }
}
SS_API int ss_im2col_patch_f16(const void* img, void* out, int B, int C, int S, int P, int Kpad, void* stream) {
  SS_REQUIRE(S % P == 0 && Kpad >= C * P * P, "bad patch geometry");
  const long long total = (long long)B * (S / P) * (S / P) * Kpad;
  im2col_patch_kernel<<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const __half*)img, (__half*)out, B, C, S,
                                                                             P, Kpad);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- y[r, :] = x[r, :] + add[r % period, :]  (positional-embedding adds, fp16/bf16 add semantics) ----
template <typename T>
__global__ void add_bcast_kernel(const T* __restrict__ x, const T* __restrict__ add, T* __restrict__ y, long long rows,
                                 int C, int period) {
  const int vecs = C >> 3;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vecs;
    const int v = (int)(i % vecs);
    float a[8], b[8];
    unpack8<T>(ld_cached16(x + r * C + v * 8), a);
    unpack8<T>(ld_cached16(add + (r % period) * C + v * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    st16(y + r * C + v * 8, pack8<T>(a));
  }
}
SS_API int ss_add_bcast(int dtype, const void* x, const void* add, void* y, long long rows, int C, int period,
                        void* stream) {
  SS_REQUIRE(C % 8 == 0 && period > 0, "C % 8, period > 0");
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = ew_grid(rows * (C / 8));
  if (dtype == SS_F16)
    add_bcast_kernel<__half><<<g, EW_THREADS, 0, s>>>((const __half*)x, (const __half*)add, (__half*)y, rows, C, period);
  else
    add_bcast_kernel<__nv_bfloat16><<<g, EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)add,
                                                             (__nv_bfloat16*)y, rows, C, period);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- dst[dst_rows[i], :] = src[i, :]  (input_embeds[ids_cmp_mask] = image_embeds_lm[...], models.py:135) ----
__global__ void scatter_rows_kernel(const uint16_t* __restrict__ src, const int* __restrict__ dst_rows,
                                    uint16_t* __restrict__ dst, int ld_dst, int width) {
  const int i = blockIdx.x;
  const vec8* s = reinterpret_cast<const vec8*>(src + (size_t)i * width);
  vec8* d = reinterpret_cast<vec8*>(dst + (size_t)dst_rows[i] * ld_dst);
  for (int v = threadIdx.x; v < (width >> 3); v += blockDim.x) d[v] = s[v];
}
SS_API int ss_scatter_rows_16b(const void* src, const int* dst_rows, void* dst, int ld_dst, int n, int width,
                               void* stream) {
  SS_REQUIRE(width % 8 == 0 && ld_dst % 8 == 0, "width % 8");
  if (n == 0) return 0;
  scatter_rows_kernel<<<n, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, dst_rows, (uint16_t*)dst, ld_dst, width);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- GroupNorm on NHWC (diffusers ResnetBlock2D norm1/norm2, Transformer2D norm, VAE norms) --------
// pass 1: per-CTA partial (sum, sum of squares) per (image, group), reduced in a FIXED order (no float atomics:
// results are bit-reproducible run to run); pass 2: fold the partials into (mean, rstd); pass 3: normalise,
// affine, optional SiLU.
template <typename T>
__global__ void groupnorm_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, int HW, int C,
                                         int groups, int pix_per_cta, float* __restrict__ stats,
                                         unsigned int* __restrict__ tickets, float inv_cnt, float eps) {
  extern __shared__ float gsm[];  // [nl][2*C]
  __shared__ int is_last;
  pdl_trigger();
  pdl_wait();  // x comes from the preceding kernel
  const int n = blockIdx.y;
  const int vecs = C >> 3;
  const int nl = blockDim.x / vecs;  // pixel lanes
  const int v = threadIdx.x % vecs, lane = threadIdx.x / vecs;
  if (lane < nl) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    const T* base = x + ((size_t)n * HW) * C + v * 8;
#pragma unroll 4
    for (int p = p0 + lane; p < p1; p += nl) {
      float f[8];
      unpack8<T>(ld_cached16(base + (size_t)p * C), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] += f[j] * f[j];
      }
    }
    float* row = gsm + (size_t)lane * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      row[v * 8 + j] = s[j];
      row[C + v * 8 + j] = q[j];
    }
  }
  __syncthreads();
  const int cg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int l = 0; l < nl; ++l) {
      const float* row = gsm + (size_t)l * 2 * C;
      for (int c = g * cg; c < (g + 1) * cg; ++c) {
        s += row[c];
        q += row[C + c];
      }
    }
    float* out = partial + (((size_t)n * gridDim.x + blockIdx.x) * groups + g) * 2;
    out[0] = s;
    out[1] = q;
  }
  // The CTA that finishes LAST for image n folds the per-CTA partials into (mean, rstd) in a fixed order — the result
  // does not depend on which CTA does it, so it stays bit-reproducible.  (Replaces a separate finalize launch.)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[n], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // thread (slice, g): lanes of a warp read 32 consecutive groups of one chunk (256 contiguous bytes), 16 independent
  // loads in flight per thread — a serial walk paid one L2 latency per chunk (~30 us of tail on ~600 chunks); the
  // slice sums are then folded in slice order, so the result does not depend on which CTA runs this or on timing
  const int chunks = gridDim.x;
  const int nslices = blockDim.x / 32;
  float* red = gsm;  // [nslices][groups][2] (the pixel-lane buffer is dead by now)
  {
    const int sl = threadIdx.x / 32, gl = threadIdx.x % 32;
    for (int g0 = 0; g0 < groups; g0 += 32) {
      const int g = g0 + gl;
      float s = 0.f, q = 0.f;
      if (sl < nslices && g < groups) {
        const float2* base = reinterpret_cast<const float2*>(partial) + (size_t)n * chunks * groups + g;
        for (int c0 = sl; c0 < chunks; c0 += nslices * 16) {
          float2 v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int c = c0 + u * nslices;
            v[u] = c < chunks ? __ldcg(base + (size_t)c * groups) : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            s += v[u].x;
            q += v[u].y;
          }
        }
        red[((size_t)sl * groups + g) * 2] = s;
        red[((size_t)sl * groups + g) * 2 + 1] = q;
      }
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int sl = 0; sl < nslices; ++sl) {
      s += red[((size_t)sl * groups + g) * 2];
      q += red[((size_t)sl * groups + g) * 2 + 1];
    }
    const float mean = s * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    stats[2 * ((size_t)n * groups + g)] = mean;
    stats[2 * ((size_t)n * groups + g) + 1] = rsqrtf(var + eps);
  }
  if (threadIdx.x == 0) tickets[n] = 0u;  // ready for the next call on this workspace
}

// apply: same (chunk, image) geometry as the partial kernel, so every thread owns one 8-channel vector — gamma,
// beta and the (at most eight) group statistics live in registers and the pixel loop has no integer divisions
template <typename T>
__global__ void groupnorm_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ stats,
                                       const T* __restrict__ gamma, const T* __restrict__ beta, int HW, int C,
                                       int groups, int silu, int pix_per_cta) {
  pdl_trigger();
  pdl_wait();  // statistics (and x) come from the preceding kernels
  const int n = blockIdx.y;
  const int vecs = C >> 3, cg = C / groups;
  const int nl = blockDim.x / vecs;
  const int v = threadIdx.x % vecs, lane = threadIdx.x / vecs;
  if (lane >= nl) return;
  float sc[8], sh[8];
  {
    float gm[8], bt[8];
    unpack8<T>(ld_cached16(gamma + v * 8), gm);
    unpack8<T>(ld_cached16(beta + v * 8), bt);
    const int c0 = v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* st = stats + ((size_t)n * groups + (c0 + j) / cg) * 2;
      sc[j] = st[1];  // rstd
      sh[j] = st[0];  // mean; the pixel loop keeps the order ((x - mean) * rstd) * gamma + beta
    }
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    const size_t base = ((size_t)n * HW) * C + v * 8;
#pragma unroll 4
    for (int p = p0 + lane; p < p1; p += nl) {
      float f[8];
      unpack8<T>(ld_cached16(x + base + (size_t)p * C), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float o = (f[j] - sh[j]) * sc[j] * gm[j] + bt[j];
        if (silu) {
          o = ss_num<T>::to_f(ss_num<T>::from_f(o));  // GroupNorm output is rounded before the activation module
          o = o / (1.f + __expf(-o));
        }
        f[j] = o;
      }
      st16(y + base + (size_t)p * C, pack8<T>(f));
    }
  }
}

// workspace layout (32-bit words): [64 arrival tickets (uint32; must be ZERO before the first call, every call leaves
// them at zero — a fixed location, so calls of different shapes can share one workspace)][2*N*groups stats]
// [N*chunks*groups*2 partials]; ss_groupnorm_ws_floats() sizes it.
constexpr int GN_TICKETS = 64;
static inline void groupnorm_geometry(int N, int HW, int C, int* block, int* chunks, int* pix_per_cta) {
  const int vecs = C / 8;
  *block = vecs <= 256 ? vecs * (256 / vecs) : vecs;
  // a thread walks its pixels serially (4 loads in flight), so the walks must stay short: the 32x32x1280 level used to
  // run 64 CTAs x 32 dependent pixels (31 us for 5 MB)
  // 4 CTAs per SM over the batch (swept 2 / 4 / 8 / 12 on the UNet step: 21.35 / 21.10 / 21.27 / 21.42 ms)
  int ch = (148 * 4 + N - 1) / N;
  int ppc = (HW + ch - 1) / ch;
  if (ppc < 8) ppc = 8;
  *pix_per_cta = ppc;
  *chunks = (HW + ppc - 1) / ppc;
}

SS_API int ss_groupnorm_ws_floats(int N, int HW, int C, int groups) {
  int block, chunks, ppc;
  groupnorm_geometry(N, HW, C, &block, &chunks, &ppc);
  return GN_TICKETS + 2 * N * groups + 2 * N * chunks * groups;
}

SS_API int ss_groupnorm_nhwc(int dtype, const void* x, void* y, const void* gamma, const void* beta, float* stats_ws,
                             int N, int HW, int C, int groups, float eps, int silu, void* stream) {
  SS_REQUIRE(C % 8 == 0 && C % groups == 0 && C <= 4096, "C % 8, C % groups, C <= 4096");
  SS_REQUIRE(dtype == SS_F16 || dtype == SS_BF16, "dtype");
  if (N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int block, chunks, pix_per_cta;
  groupnorm_geometry(N, HW, C, &block, &chunks, &pix_per_cta);
  SS_REQUIRE(block <= 1024, "C too large for the GroupNorm stats kernel");
  const int vecs = C / 8, nl = block / vecs;
  const size_t smem = (size_t)nl * 2 * C * sizeof(float);
  SS_REQUIRE(smem <= 48 * 1024, "GroupNorm partial kernel shared memory");
  SS_REQUIRE(N <= GN_TICKETS, "GroupNorm batch > 64 images");
  unsigned int* tickets = reinterpret_cast<unsigned int*>(stats_ws);
  float* stats = stats_ws + GN_TICKETS;
  float* partial = stats + 2 * N * groups;
  const float inv_cnt = 1.f / ((float)HW * (float)(C / groups));
  if (dtype == SS_F16) {
    SS_CUDA(ss::launch_pdl(groupnorm_partial_kernel<__half>, dim3(chunks, N), dim3(block), smem, s, (const __half*)x,
                           partial, HW, C, groups, pix_per_cta, stats, tickets, inv_cnt, eps));
    SS_CUDA(ss::launch_pdl(groupnorm_apply_kernel<__half>, dim3(chunks, N), dim3(block), 0, s, (const __half*)x,
                           (__half*)y, (const float*)stats, (const __half*)gamma, (const __half*)beta, HW, C, groups, silu,
                           pix_per_cta));
  } else {
    SS_CUDA(ss::launch_pdl(groupnorm_partial_kernel<__nv_bfloat16>, dim3(chunks, N), dim3(block), smem, s,
                           (const __nv_bfloat16*)x, partial, HW, C, groups, pix_per_cta, stats, tickets, inv_cnt, eps));
    SS_CUDA(ss::launch_pdl(groupnorm_apply_kernel<__nv_bfloat16>, dim3(chunks, N), dim3(block), 0, s,
                           (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const float*)stats, (const __nv_bfloat16*)gamma,
                           (const __nv_bfloat16*)beta, HW, C, groups, silu, pix_per_cta));
  }
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- nearest 2x upsample, NHWC (Upsample2D before its conv) ------------------------------------
__global__ void upsample2x_kernel(const vec8* __restrict__ x, vec8* __restrict__ y, int N, int H, int W, int vecs) {
  const long long total = (long long)N * 2 * H * 2 * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int ox = (int)(p % (2 * W));
    p /= 2 * W;
    const int oy = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    y[i] = x[(((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * vecs + v];
  }
}
SS_API int ss_upsample2x_nhwc_16b(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  SS_REQUIRE(C % 8 == 0, "C % 8");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  upsample2x_kernel<<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const vec8*)x, (vec8*)y, N, H, W, C / 8);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- channel concat on NHWC rows: out[r] = [a[r] | b[r]] (UNet skip connections) ----------------
__global__ void concat_rows_kernel(const vec8* __restrict__ a, const vec8* __restrict__ b, vec8* __restrict__ out,
                                   long long rows, int va, int vb) {
  const int vo = va + vb;
  const long long total = rows * vo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vo);
    const long long r = i / vo;
    out[i] = v < va ? a[r * va + v] : b[r * vb + (v - va)];
  }
}
SS_API int ss_concat_channels_16b(const void* a, const void* b, void* out, long long rows, int Ca, int Cb,
                                  void* stream) {
  SS_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "C % 8");
  const long long total = rows * ((Ca + Cb) / 8);
  concat_rows_kernel<<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const vec8*)a, (const vec8*)b, (vec8*)out,
                                                                             rows, Ca / 8, Cb / 8);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- im2col for the two stride-2 3x3 Downsample2D convs (pad 1): cols[(n,oy,ox), (ky*3+kx)*C + c] ----
__global__ void im2col_s2_kernel(const vec8* __restrict__ x, vec8* __restrict__ cols, int N, int H, int W, int vecs) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * 9 * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int tap = (int)(p % 9);
    p /= 9;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    const int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;
    vec8 val{0u, 0u, 0u, 0u};
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = x[(((long long)n * H + iy) * W + ix) * vecs + v];
    cols[i] = val;
  }
}
SS_API int ss_im2col3x3_s2_nhwc_16b(const void* x, void* cols, int N, int H, int W, int C, void* stream) {
  SS_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "C % 8, even H/W");
  const long long total = (long long)N * (H / 2) * (W / 2) * 9 * (C / 8);
  im2col_s2_kernel<<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const vec8*)x, (vec8*)cols, N, H, W, C / 8);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- classifier-free guidance + Euler step (diffusers EulerDiscreteScheduler.step, eps-prediction) ----
// eps [2, HW, Cpad] NHWC (row 0 = uncond, row 1 = cond; first C channels valid), latents fp16 [HW, C]
// (NHWC of the [1,4,128,128] latent).  x <- x + (eu + g (ec - eu)) * (sigma_next - sigma) in fp32, rounded to
// fp16; also emits the next model input x / sqrt(sigma_next^2 + 1), duplicated for both CFG rows, channel-padded.
__global__ void cfg_euler_kernel(const __half* __restrict__ eps, int eps_ld, __half* __restrict__ lat,
                                 __half* __restrict__ next_in, int in_ld, int HW, int C, float guidance, float sigma,
                                 float sigma_next) {
  const int total = HW * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = i / C, c = i % C;
    // pipeline: noise_pred = uncond + g*(text - uncond) in fp16 tensors (each op rounded)
    const __half eu = eps[(size_t)p * eps_ld + c], ec = eps[((size_t)HW + p) * eps_ld + c];
    const __half diff = __float2half_rn(__half2float(ec) - __half2float(eu));
    const __half gd = __float2half_rn(guidance * __half2float(diff));
    const __half e = __float2half_rn(__half2float(eu) + __half2float(gd));
    // scheduler.step: upcast sample to fp32, pred_original = x - sigma*eps, derivative = (x - pred)/sigma, x += d*dt
    const float x = __half2float(lat[i]);
    const float ef = __half2float(e);
    const float pred = x - sigma * ef;
    const float deriv = (x - pred) / sigma;
    const __half xn = __float2half_rn(x + deriv * (sigma_next - sigma));
    lat[i] = xn;
    if (next_in) {
      const __half scaled = __float2half_rn(__half2float(xn) / sqrtf(sigma_next * sigma_next + 1.f));
      next_in[(size_t)p * in_ld + c] = scaled;
      next_in[((size_t)HW + p) * in_ld + c] = scaled;
    }
  }
}
SS_API int ss_cfg_euler_step_f16(const void* eps, int eps_ld, void* latents, void* next_in, int in_ld, int HW, int C,
                                 float guidance, float sigma, float sigma_next, void* stream) {
  cfg_euler_kernel<<<ew_grid((long long)HW * C), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const __half*)eps, eps_ld, (__half*)latents, (__half*)next_in, in_ld, HW, C, guidance, sigma, sigma_next);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- generic casts / scale: y = T2(x * scale) ------------------------------------------------------
template <typename TI, typename TO>
__global__ void cast_scale_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = ss_num<TO>::from_f(ss_num<TI>::to_f(x[i]) * scale);
}
SS_API int ss_cast_scale(int dtype_in, const void* x, int dtype_out, void* y, long long n, float scale, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int g = ew_grid(n);
  if (dtype_in == SS_F16 && dtype_out == SS_BF16)
    cast_scale_kernel<__half, __nv_bfloat16><<<g, EW_THREADS, 0, s>>>((const __half*)x, (__nv_bfloat16*)y, n, scale);
  else if (dtype_in == SS_BF16 && dtype_out == SS_F16)
    cast_scale_kernel<__nv_bfloat16, __half><<<g, EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__half*)y, n, scale);
  else if (dtype_in == SS_F16 && dtype_out == SS_F16)
    cast_scale_kernel<__half, __half><<<g, EW_THREADS, 0, s>>>((const __half*)x, (__half*)y, n, scale);
  else if (dtype_in == SS_BF16 && dtype_out == SS_BF16)
    cast_scale_kernel<__nv_bfloat16, __nv_bfloat16><<<g, EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                             n, scale);
  else
    SS_FAIL("unsupported cast");
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- row softmax (VAE mid-block attention scores), fp32 math, in place --------------------------------
template <typename T>
__global__ void softmax_rows_kernel(T* __restrict__ x, int ld, int n, float scale) {
  __shared__ float red[32];
  T* row = x + (size_t)blockIdx.x * ld;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, ss_num<T>::to_f(row[i]) * scale);
  m = block_max(m, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += __expf(ss_num<T>::to_f(row[i]) * scale - m);
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    row[i] = ss_num<T>::from_f(__expf(ss_num<T>::to_f(row[i]) * scale - m) * inv);
}
SS_API int ss_softmax_rows(int dtype, void* x, int ld, int rows, int n, float scale, void* stream) {
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SS_F16)
    softmax_rows_kernel<__half><<<rows, 512, 0, s>>>((__half*)x, ld, n, scale);
  else
    softmax_rows_kernel<__nv_bfloat16><<<rows, 512, 0, s>>>((__nv_bfloat16*)x, ld, n, scale);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- 2-D transpose of 16-bit elements: y[c, r] = x[r, c] ---------------------------------------------
__global__ void transpose_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int R, int C) {
  __shared__ uint16_t tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = x[(size_t)r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) y[(size_t)c * R + r] = tile[threadIdx.x][j];
  }
}
SS_API int ss_transpose_16b(const void* x, void* y, int R, int C, void* stream) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint16_t*)x, (uint16_t*)y, R, C);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- mean over tokens: y[b, :] = mean_t x[b, t, :]  (AttentionPool2d, resampler.py:92) ----------------
__global__ void mean_tokens_kernel(const __half* __restrict__ x, __half* __restrict__ y, int T, int C) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += __half2float(x[((size_t)b * T + t) * C + c]);
  y[(size_t)b * C + c] = __float2half_rn(s / (float)T);
}
SS_API int ss_mean_tokens_f16(const void* x, void* y, int B, int T, int C, void* stream) {
  if (B == 0) return 0;
  mean_tokens_kernel<<<dim3(ceil_div(C, 128), B), 128, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, T, C);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- VAE output -> uint8 HWC image: (x/2 + 0.5).clamp(0,1) * 255, round (diffusers VaeImageProcessor) ----
template <typename T>
__global__ void to_uint8_kernel(const T* __restrict__ x, int ldx, uint8_t* __restrict__ out, long long pixels, int C) {
  const long long total = pixels * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const int c = (int)(i % C);
    float v = ss_num<T>::to_f(x[p * ldx + c]) * 0.5f + 0.5f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = (uint8_t)rintf(v * 255.f);
  }
}
SS_API int ss_image_to_uint8(int dtype, const void* x, int ldx, void* out, long long pixels, int C, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SS_F16)
    to_uint8_kernel<__half><<<ew_grid(pixels * C), EW_THREADS, 0, s>>>((const __half*)x, ldx, (uint8_t*)out, pixels, C);
  else
    to_uint8_kernel<__nv_bfloat16><<<ew_grid(pixels * C), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, ldx,
                                                                              (uint8_t*)out, pixels, C);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---- y = act(x) -----------------------------------------------------------------------------------
template <typename T>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int op) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = ss_num<T>::to_f(x[i]);
    y[i] = ss_num<T>::from_f(op == 1 ? gelu_erf(v) : v / (1.f + expf(-v)));
  }
}
SS_API int ss_unary(int dtype, const void* x, void* y, long long n, int op, void* stream) {
  SS_REQUIRE(op == 1 || op == 2, "op: 1 gelu, 2 silu");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == SS_F16)
    unary_kernel<__half><<<ew_grid(n), EW_THREADS, 0, s>>>((const __half*)x, (__half*)y, n, op);
  else
    unary_kernel<__nv_bfloat16><<<ew_grid(n), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, op);
  SS_LAUNCH_CHECK();
  return 0;
}
