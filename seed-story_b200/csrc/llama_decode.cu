// Llama decode-step kernels (HBM-bound path, batch <= 8 sequences per rank):
//   * skinny GEMM  y[b, n] = sum_k x[b, k] W[n, k]      (weight-streaming, fused epilogues)
//   * RoPE + paged KV append                            (reference: modeling_llama_xformer.py:165-173, 236-244)
//   * paged single-query attention, split over KV pages (reference: :282-295, bottom-right causal mask
//                                                        degenerates to "see every cached token")
//   * image-token logits processor + greedy argmax      (reference: src/models_clm/generation.py:19-31)
//   * token-embedding gather, decode-state advance
//
// The skinny GEMM streams W once with 16-byte no-allocate loads and uses mma.sync.m16n8k16 purely as
// a convenient 16x8 dot-product engine (N = 8 batch columns); it is bandwidth-bound by construction.
#include <cstdlib>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// skinny GEMM
// ---------------------------------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_ROPE_APPEND = 3 };
constexpr int KV_PAGE = 64;

__device__ __forceinline__ __half hadd_t(__half a, __half b) {  // torch's half add: fp32 add, one rounding
  return __float2half_rn(__half2float(a) + __half2float(b));
}

__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int SG_WARPS_MAX = 8;
constexpr int SG_UNROLL = 8;   // k-blocks (32 wide) per register batch; two batches are in flight per warp
constexpr int SG_ROWS = 8;     // weight rows per CTA
constexpr int SG_XS_PAD = 32;  // halves of padding per staged activation row (64 B: rows land in alternate bank halves)

// Fused prologue / epilogue operands of the decode-layer kernels (all device pointers).
struct SkinnyFuse {
  // NORM prologue: x is the raw residual-stream row; the CTA recomputes LlamaRMSNorm (modeling_llama_xformer.py:107-115,
  // same rounding chain and the same reduction order as rmsnorm_f16_kernel) into shared memory before the dot products
  const __half* gamma;
  float eps;
  // EPI_ROPE_APPEND: W rows of the q and k sections are stored pair-interleaved per head (row 2i = dim i, row 2i+1 =
  // dim i + D/2), so the two halves of a rotary pair are neighbours in the CTA's 8 output rows
  __half* q_out;             // [B, H*D] natural order
  __half* kcache;            // this layer's K pages [page][H][64][D]
  __half* vcache;
  const long long* kv_base;  // [B] element offset of (page, head 0, slot % 64, dim 0) of each sequence's new token
  const __half* rope_cs;     // [B, D] cos / sin rows of each sequence's position (ss_decode_rope_meta)
  const __half* rope_sn;
  int H, D;
};

// One CTA = 8 rows of W.  The weight rows are the B operand (n = row), the <= 8 activation rows the A operand
// (m = batch index, rows 8..15 zero), so a lane streams ONE 16-byte piece of one weight row per 32-wide k-block;
// the 8 warps interleave over k-blocks, keep two register batches of loads in flight (software pipeline) and
// reduce their 8x8 partial results through shared memory.
template <int EPI, int SG_WARPS, int NV>   // NV: 0 = x is used as given; 2 / 4 = RMSNorm prologue for K <= 4096 / 8192
__global__ void __launch_bounds__(SG_WARPS * 32, NV > 0 ? 3 : 2) skinny_gemm_kernel(const __half* __restrict__ x, int ldx,
                                                                    const __half* __restrict__ W,
                                                                    __half* __restrict__ y, int ldy, int B, int N,
                                                                    int K, const __half* __restrict__ res, int ldr,
                                                                    const SkinnyFuse fz) {
  constexpr bool NORM = NV > 0;
  __shared__ float part[SG_WARPS_MAX][8][8];  // [warp][batch][row]
  __shared__ float red[32];
  extern __shared__ __align__(16) uint8_t sg_dyn[];
  __half* xs = reinterpret_cast<__half*>(sg_dyn);  // NORM: [B][K + SG_XS_PAD] normalised activations
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * SG_ROWS;
  const int r_w = min(row0 + g, N - 1);  // clamp: tail rows are recomputed, never stored
  const __half* wp = W + (size_t)r_w * K + t * 8;
  const int xs_ld = K + SG_XS_PAD;
  const __half* xg = NORM ? (xs + (size_t)min(g, B - 1) * xs_ld + t * 8) : (x + (size_t)min(g, B - 1) * ldx + t * 8);
  const bool xvalid = g < B;
  const int nkb = K >> 5;
  const int my_n = (nkb - warp + SG_WARPS - 1) / SG_WARPS;  // k-blocks owned by this warp: warp, warp+8, ...

  float c[4] = {0.f, 0.f, 0.f, 0.f};
  pdl_trigger();
  vec8 wa[SG_UNROLL], wb[SG_UNROLL];
  auto load_batch = [&](vec8* dst, int first) {
#pragma unroll
    for (int u = 0; u < SG_UNROLL; ++u) {
      const int i = first + u;
      dst[u] = (i < my_n) ? ld_stream16(wp + ((size_t)(warp + i * SG_WARPS) << 5)) : vec8{0u, 0u, 0u, 0u};
    }
  };
  auto compute_batch = [&](const vec8* src, int first) {
    vec8 xb[SG_UNROLL];
#pragma unroll
    for (int u = 0; u < SG_UNROLL; ++u) {
      const int i = first + u;
      if (NORM)
        xb[u] = (xvalid && i < my_n) ? *reinterpret_cast<const vec8*>(xg + ((size_t)(warp + i * SG_WARPS) << 5))
                                     : vec8{0u, 0u, 0u, 0u};
      else
        xb[u] = (xvalid && i < my_n) ? ld_cached16(xg + ((size_t)(warp + i * SG_WARPS) << 5)) : vec8{0u, 0u, 0u, 0u};
    }
#pragma unroll
    for (int u = 0; u < SG_UNROLL; ++u) {
      // same k permutation on both operands: piece j of the 16-byte vector feeds k-pairs (2t,2t+1)/(2t+8,2t+9)
      mma16816(c, xb[u].x, 0u, xb[u].y, 0u, src[u].x, src[u].y);
      mma16816(c, xb[u].z, 0u, xb[u].w, 0u, src[u].z, src[u].w);
    }
  };
  // weights are constants: they are requested before we wait for the producer of x, so the chip-wide prefetch covers
  // the kernel boundary.  Without a norm prologue BOTH register batches go out now (for K = 4096 that is all of this
  // CTA's 64 KB; measured 10.3 -> 9.4 us on o_proj, 22.2 -> 19.9 us on down_proj); with the prologue the second batch
  // would have to live across it (125 registers, two CTAs per SM: measured slower), so it follows the prologue.
  load_batch(wa, 0);
  if (!NORM) load_batch(wb, SG_UNROLL);
  pdl_wait();
  // EPI_ROPE_APPEND: everything the epilogue needs besides the dot product (cache slot, page, rotary factors) is
  // fetched now, so that the tail of the CTA is arithmetic + one store instead of a chain of dependent loads
  __half rope_cs = __float2half_rn(0.f), rope_sn = rope_cs;
  long long rope_dst = -1;  // element offset inside q_out (sec 0) / the K or V page pool (sec 1, 2); -1 = nothing to store
  int rope_sec = 0, rope_odd = 0;
  if (EPI == EPI_ROPE_APPEND && threadIdx.x < 64) {
    const int n = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int row = row0 + r;
    if (n < B && row < N) {
      const int HD = fz.H * fz.D, half_d = fz.D >> 1;
      const int sec = row / HD, within = row - sec * HD;
      const int h = within / fz.D, p = within - h * fz.D;
      const int d = (p >> 1) + (p & 1) * half_d;   // q / k rows are pair-interleaved, v rows natural
      rope_sec = sec;
      rope_odd = p & 1;
      if (sec == 0) {
        rope_dst = (long long)n * HD + h * fz.D + d;
      } else {
        rope_dst = fz.kv_base[n] + (long long)h * KV_PAGE * fz.D + (sec == 1 ? d : p);
      }
      if (sec != 2) {   // independent loads: no index chain in front of the rotary factors
        rope_cs = fz.rope_cs[n * fz.D + d];
        rope_sn = fz.rope_sn[n * fz.D + d];
      }
    }
  }
  if (NORM) {
    // RMSNorm of the <= 8 activation rows, recomputed per CTA (8 KB per row out of L2) while the first weight batch
    // is in flight: one pass, the row pieces stay in registers between the sum of squares and the scaling; arithmetic
    // and reduction order are those of rmsnorm_f16_kernel (256 threads, pieces tid, tid + 256, ...)
    const int nvec = K >> 3;
    vec8 gv[NV > 0 ? NV : 1], xv[NV > 0 ? NV : 1];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = threadIdx.x + i * (SG_WARPS * 32);
      gv[i] = (vi < nvec) ? ld_cached16(fz.gamma + vi * 8) : vec8{0u, 0u, 0u, 0u};
      xv[i] = (vi < nvec) ? ld_cached16(x + vi * 8) : vec8{0u, 0u, 0u, 0u};
    }
    for (int b = 0; b < B; ++b) {
      vec8 xn[NV > 0 ? NV : 1];  // next row's pieces: their latency hides behind this row's reduction
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = threadIdx.x + i * (SG_WARPS * 32);
        xn[i] = (b + 1 < B && vi < nvec) ? ld_cached16(x + (size_t)(b + 1) * ldx + vi * 8) : vec8{0u, 0u, 0u, 0u};
      }
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float f[8];
        unpack8<__half>(xv[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
      }
      ss = block_sum(ss, red);
      const float rstd = __frsqrt_rn(ss / (float)K + fz.eps);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = threadIdx.x + i * (SG_WARPS * 32);
        if (vi < nvec) {
          float f[8];
          unpack8<__half>(xv[i], f);
          const __half* wh = reinterpret_cast<const __half*>(&gv[i]);
          vec8 o;
          __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
          for (int j = 0; j < 8; ++j) oh[j] = __hmul(wh[j], __float2half_rn(f[j] * rstd));
          *reinterpret_cast<vec8*>(xs + (size_t)b * xs_ld + vi * 8) = o;
        }
        xv[i] = xn[i];
      }
    }
    __syncthreads();
  }
  if (NORM) load_batch(wb, SG_UNROLL);
  for (int first = 0; first < my_n; first += 2 * SG_UNROLL) {
    compute_batch(wa, first);
    load_batch(wa, first + 2 * SG_UNROLL);
    compute_batch(wb, first + SG_UNROLL);
    load_batch(wb, first + 3 * SG_UNROLL);
  }
  // C fragment: c0,c1 -> (batch g, rows 2t,2t+1); c2,c3 belong to the zero half of A
  part[warp][g][2 * t] = c[0];
  part[warp][g][2 * t + 1] = c[1];
  __syncthreads();

  if (EPI == EPI_SWIGLU) {
    // tile rows (2j, 2j+1) are (gate_j, up_j): the host interleaves gate/up rows pairwise, the same
    // packing the tensor-core GEMM's GLU epilogue consumes
    if (threadIdx.x < 32) {
      const int n = threadIdx.x >> 2, j = threadIdx.x & 3;
      float gate = 0.f, up = 0.f;
#pragma unroll
      for (int w = 0; w < SG_WARPS; ++w) {
        gate += part[w][n][2 * j];
        up += part[w][n][2 * j + 1];
      }
      const int out_col = blockIdx.x * 4 + j;
      if (n < B && out_col < (N >> 1)) {
        const __half gh = __float2half_rn(gate);
        const float gf = __half2float(gh);
        const __half act = __float2half_rn(gf / (1.f + expf(-gf)));  // silu evaluated in fp32, rounded (torch half kernel)
        y[(size_t)n * ldy + out_col] = __hmul(act, __float2half_rn(up));
      }
    }
  } else if (EPI == EPI_ROPE_APPEND) {
    // q/k/v projection of one decode token per sequence + apply_rotary_pos_emb (:165-173) + cache append (:241-242):
    // thread (n, r) owns output row row0 + r of sequence n; rows (2i, 2i+1) of a q/k head are the rotary pair
    // (d, d + D/2), so the partner value is one lane away.  Same fp16 arithmetic as rope_append_kernel.
    if (threadIdx.x < 64) {
      const int n = threadIdx.x >> 3, r = threadIdx.x & 7;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < SG_WARPS; ++w) acc += part[w][n][r];
      const __half o = __float2half_rn(acc);
      const __half other = __ushort_as_half(__shfl_xor_sync(0xffffffffu, __half_as_ushort(o), 1));
      if (rope_dst >= 0) {
        if (rope_sec == 2) {
          fz.vcache[rope_dst] = o;
        } else {
          // d < D/2: x_d cos_d + (-x_{d+D/2}) sin_d ;  d >= D/2: x_d cos_d + x_{d-D/2} sin_d
          const __half v = rope_odd ? hadd_t(__hmul(o, rope_cs), __hmul(other, rope_sn))
                                    : hadd_t(__hmul(o, rope_cs), __hmul(__hneg(other), rope_sn));
          if (rope_sec == 0)
            fz.q_out[rope_dst] = v;
          else
            fz.kcache[rope_dst] = v;
        }
      }
    }
  } else {
    if (threadIdx.x < 64) {
      const int n = threadIdx.x >> 3, r = threadIdx.x & 7;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < SG_WARPS; ++w) acc += part[w][n][r];
      const int row = row0 + r;
      if (n < B && row < N) {
        __half o = __float2half_rn(acc);
        if (EPI == EPI_RESIDUAL) o = __float2half_rn(__half2float(res[(size_t)n * ldr + row]) + __half2float(o));
        y[(size_t)n * ldy + row] = o;
      }
    }
  }
}

template <int EPI, bool NORM>
static int skinny_launch(const __half* xp, int ldx, const __half* Wp, __half* yp, int ldy, int B, int N, int K,
                         const __half* rp, int ldr, const SkinnyFuse& fz, cudaStream_t s) {
  const int grid = ceil_div(N, SG_ROWS);
  const size_t smem = NORM ? (size_t)B * (K + SG_XS_PAD) * sizeof(__half) : 0;
  auto go = [&](auto k) -> int {
    if (smem > 48 * 1024) {
      static bool raised = false;
      if (!raised) {
        SS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (8192 + SG_XS_PAD) * 2));
        raised = true;
      }
    }
    SS_CUDA(ss::launch_pdl(k, dim3(grid), dim3(256), smem, s, xp, ldx, Wp, yp, ldy, B, N, K, rp, ldr, fz));
    SS_LAUNCH_CHECK();
    return 0;
  };
  if (!NORM) return go(skinny_gemm_kernel<EPI, 8, 0>);
  if (K <= 4096) return go(skinny_gemm_kernel<EPI, 8, NORM ? 2 : 0>);
  return go(skinny_gemm_kernel<EPI, 8, NORM ? 4 : 0>);
}

SS_API int ss_skinny_gemm_f16(const void* x, int ldx, const void* W, void* y, int ldy, int B, int N, int K,
                              int epilogue, const void* residual, int ldr, void* stream) {
  SS_REQUIRE(B >= 1 && B <= 8, "skinny GEMM handles 1..8 rows");
  SS_REQUIRE(K % 32 == 0 && ldx % 8 == 0, "K must be a multiple of 32, ldx of 8");
  SS_REQUIRE(epilogue != EPI_SWIGLU || N % 8 == 0, "SwiGLU needs N % 8 == 0");
  cudaStream_t s = (cudaStream_t)stream;
  const __half *xp = (const __half*)x, *Wp = (const __half*)W, *rp = (const __half*)residual;
  __half* yp = (__half*)y;
  SkinnyFuse fz = {};
  switch (epilogue) {
    case EPI_NONE:
      return skinny_launch<EPI_NONE, false>(xp, ldx, Wp, yp, ldy, B, N, K, rp, ldr, fz, s);
    case EPI_RESIDUAL:
      SS_REQUIRE(residual != nullptr, "residual epilogue needs a residual pointer");
      return skinny_launch<EPI_RESIDUAL, false>(xp, ldx, Wp, yp, ldy, B, N, K, rp, ldr, fz, s);
    case EPI_SWIGLU:
      return skinny_launch<EPI_SWIGLU, false>(xp, ldx, Wp, yp, ldy, B, N, K, rp, ldr, fz, s);
    default:
      SS_FAIL("unknown epilogue");
  }
}

// RMSNorm fused into the projection that consumes it: y = epilogue(LlamaRMSNorm(x; gamma, eps) W^T)
// (input_layernorm -> q/k/v, post_attention_layernorm -> gate/up: modeling_llama_xformer.py:341-359).
SS_API int ss_skinny_gemm_rmsnorm_f16(const void* x, int ldx, const void* gamma, float eps, const void* W, void* y,
                                      int ldy, int B, int N, int K, int epilogue, void* stream) {
  SS_REQUIRE(B >= 1 && B <= 8, "skinny GEMM handles 1..8 rows");
  SS_REQUIRE(K % 32 == 0 && ldx % 8 == 0 && K <= 8192, "K must be a multiple of 32 (<= 8192), ldx of 8");
  SS_REQUIRE(epilogue == EPI_NONE || (epilogue == EPI_SWIGLU && N % 8 == 0), "epilogue must be none or SwiGLU (N % 8 == 0)");
  SkinnyFuse fz = {};
  fz.gamma = (const __half*)gamma;
  fz.eps = eps;
  cudaStream_t s = (cudaStream_t)stream;
  if (epilogue == EPI_SWIGLU)
    return skinny_launch<EPI_SWIGLU, true>((const __half*)x, ldx, (const __half*)W, (__half*)y, ldy, B, N, K, nullptr, 0, fz, s);
  return skinny_launch<EPI_NONE, true>((const __half*)x, ldx, (const __half*)W, (__half*)y, ldy, B, N, K, nullptr, 0, fz, s);
}

// The attention input side of a decode layer in ONE launch: input RMSNorm -> q/k/v projection -> rotary embedding
// -> q to q_out, k (post-RoPE) and v appended to the paged cache (modeling_llama_xformer.py:341, 228-244).
// Wqkv_il: [3*H*D, K]; inside every q and k head the rows are pair-interleaved (row 2i = dim i, 2i+1 = dim i + D/2).
SS_API int ss_decode_qkv_rope_append_f16(const void* x, int ldx, const void* gamma, float eps, const void* Wqkv_il,
                                         void* q_out, void* kcache, void* vcache, const long long* kv_base,
                                         const void* rope_cos, const void* rope_sin, int B, int H, int D, int K,
                                         void* stream) {
  SS_REQUIRE(B >= 1 && B <= 8, "decode handles 1..8 sequences");
  SS_REQUIRE(K % 32 == 0 && ldx % 8 == 0 && K <= 8192, "K must be a multiple of 32 (<= 8192), ldx of 8");
  SS_REQUIRE(D % 8 == 0 && D <= 256, "head dim must be a multiple of 8 and <= 256");
  SkinnyFuse fz = {};
  fz.gamma = (const __half*)gamma;
  fz.eps = eps;
  fz.q_out = (__half*)q_out;
  fz.kcache = (__half*)kcache;
  fz.vcache = (__half*)vcache;
  fz.kv_base = kv_base;
  fz.rope_cs = (const __half*)rope_cos;
  fz.rope_sn = (const __half*)rope_sin;
  fz.H = H;
  fz.D = D;
  return skinny_launch<EPI_ROPE_APPEND, true>((const __half*)x, ldx, (const __half*)Wqkv_il, nullptr, 0, B, 3 * H * D, K,
                                              nullptr, 0, fz, (cudaStream_t)stream);
}

// Per-step constants of the fused q/k/v kernel, one CTA per sequence: where the new token's K / V rows go
// (kv_base[b] = element offset of (page, head 0, slot % 64, dim 0); the page ids are the same in every layer) and the
// cos / sin rows of its position.  Runs once per decode step (after the state advance), not once per layer.
__global__ void decode_rope_meta_kernel(const int* __restrict__ tok_seq, const int* __restrict__ tok_pos,
                                        const int* __restrict__ tok_slot, const int* __restrict__ page_table,
                                        int max_pages, const __half* __restrict__ cos_t, const __half* __restrict__ sin_t,
                                        int H, int D, long long* __restrict__ kv_base, __half* __restrict__ rope_cs,
                                        __half* __restrict__ rope_sn) {
  const int b = blockIdx.x;
  const int seq = tok_seq[b], pos = tok_pos[b], slot = tok_slot[b];
  if (threadIdx.x == 0) {
    const int page = page_table[(size_t)seq * max_pages + slot / KV_PAGE];
    kv_base[b] = ((long long)page * H * KV_PAGE + (slot % KV_PAGE)) * D;
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    rope_cs[(size_t)b * D + d] = cos_t[(size_t)pos * D + d];
    rope_sn[(size_t)b * D + d] = sin_t[(size_t)pos * D + d];
  }
}

SS_API int ss_decode_rope_meta(const int* tok_seq, const int* tok_pos, const int* tok_slot, int B, const int* page_table,
                               int max_pages, const void* cos_table, const void* sin_table, int H, int D,
                               long long* kv_base, void* rope_cos, void* rope_sin, void* stream) {
  if (B == 0) return 0;
  decode_rope_meta_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(tok_seq, tok_pos, tok_slot, page_table, max_pages,
                                                                (const __half*)cos_table, (const __half*)sin_table, H, D,
                                                                kv_base, (__half*)rope_cos, (__half*)rope_sin);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RoPE + paged KV append.  One CTA per token, 128 threads... each thread owns pairs (d, d+64) of
// one head at a time.  All arithmetic in fp16 exactly as the reference: q*cos + rotate_half(q)*sin
// with fp16 tables (two fp16 products, one fp16 add).
//   qkv   [ntok, 3*H*D]  (q | k | v)
//   q_out [ntok, H*D]
//   K/V cache pages: [page][H][PAGE][D]
// ---------------------------------------------------------------------------------------------
// grid (ntok, H): one CTA per (token, head), D/2 threads: thread d owns the pair (d, d + D/2)
__global__ void __launch_bounds__(128) rope_append_kernel(const __half* __restrict__ qkv, int ld_qkv,
                                                          __half* __restrict__ q_out, __half* __restrict__ kcache,
                                                          __half* __restrict__ vcache, const int* __restrict__ tok_seq,
                                                          const int* __restrict__ tok_pos,
                                                          const int* __restrict__ tok_slot,
                                                          const int* __restrict__ page_table, int max_pages,
                                                          const __half* __restrict__ cos_t,
                                                          const __half* __restrict__ sin_t, int H, int D) {
  pdl_trigger();
  pdl_wait();
  const int tok = blockIdx.x, h = blockIdx.y;
  const int half_d = D >> 1;
  const int d = threadIdx.x;
  if (d >= half_d) return;
  const int seq = tok_seq[tok], pos = tok_pos[tok], slot = tok_slot[tok];
  const int page = page_table[(size_t)seq * max_pages + slot / KV_PAGE];
  const int in_page = slot % KV_PAGE;
  const __half* row = qkv + (size_t)tok * ld_qkv;
  const __half* cr = cos_t + (size_t)pos * D;
  const __half* sr = sin_t + (size_t)pos * D;
  const int HD = H * D;
  const __half c0 = cr[d], c1 = cr[d + half_d], s0 = sr[d], s1 = sr[d + half_d];
  const size_t dst = (((size_t)page * H + h) * KV_PAGE + in_page) * D;
  const __half q0 = row[h * D + d], q1 = row[h * D + d + half_d];
  const __half k0 = row[HD + h * D + d], k1 = row[HD + h * D + d + half_d];
  const __half v0 = row[2 * HD + h * D + d], v1 = row[2 * HD + h * D + d + half_d];
  q_out[(size_t)tok * HD + h * D + d] = hadd_t(__hmul(q0, c0), __hmul(__hneg(q1), s0));
  q_out[(size_t)tok * HD + h * D + d + half_d] = hadd_t(__hmul(q1, c1), __hmul(q0, s1));
  kcache[dst + d] = hadd_t(__hmul(k0, c0), __hmul(__hneg(k1), s0));
  kcache[dst + d + half_d] = hadd_t(__hmul(k1, c1), __hmul(k0, s1));
  vcache[dst + d] = v0;
  vcache[dst + d + half_d] = v1;
}

SS_API int ss_rope_kv_append_f16(const void* qkv, int ld_qkv, void* q_out, void* kcache, void* vcache,
                                 const int* tok_seq, const int* tok_pos, const int* tok_slot, int ntok,
                                 const int* page_table, int max_pages, const void* cos_table, const void* sin_table,
                                 int H, int D, void* stream) {
  SS_REQUIRE(D % 2 == 0 && D <= 256, "head dim must be even and <= 256");
  if (ntok == 0) return 0;
  SS_CUDA(ss::launch_pdl(rope_append_kernel, dim3(ntok, H), dim3(128), 0, (cudaStream_t)stream, (const __half*)qkv, ld_qkv,
                         (__half*)q_out, (__half*)kcache, (__half*)vcache, tok_seq, tok_pos, tok_slot, page_table,
                         max_pages, (const __half*)cos_table, (const __half*)sin_table, H, D));
  SS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Paged decode attention, D = 128, one query token per sequence (HBM-bound: K and V pages are streamed once).
// grid (H, B, S): CTA s of (head h, sequence b) takes pages s, s+S, ... of the sequence (lengths are read on the
// device, so the launch is CUDA-graph replayable while sequences grow).  Per 64-token page every thread issues its 8
// K and 8 V 16-byte loads BEFORE any arithmetic — one memory latency per page instead of one per pass — and the
// eight half-warps (one key row = 16 lanes x 16 bytes) keep private running maxima / sums / output rows, merged once
// at the end, so the page loop has no block-wide barrier.  The last CTA of a (b, h) to arrive (device counter)
// merges the S partial results: split and combine are ONE launch.
// ---------------------------------------------------------------------------------------------
constexpr int AD_THREADS = 128;
constexpr int AD_GROUPS = AD_THREADS / 16;  // half-warps
constexpr int AD_COUNTER_WORDS = 1024;      // arrival counters (one per (sequence, head)) at the front of the workspace

// NOTE on shared memory: this kernel deliberately uses only a few KB of STATIC shared memory.  A version that staged
// the pages in 64 KB of dynamic shared memory (bulk copies) pushed the SMs' L1 / shared split to "all shared" for the
// whole PDL-chained decode graph, and with the small L1 every weight-streaming GEMM of the step ran at ~60 % of its
// bandwidth (decode step 3.1 -> 5.0 ms).  Pages are therefore staged in registers.
__global__ void __launch_bounds__(AD_THREADS, 4) attn_decode_kernel(
    const __half* __restrict__ q, const __half* __restrict__ kcache, const __half* __restrict__ vcache,
    const int* __restrict__ seq_lens, const int* __restrict__ page_table, int max_pages, float* __restrict__ part,
    int* __restrict__ counters, __half* __restrict__ out, int H, int S, float scale) {
  constexpr int D = 128;
  __shared__ float osm[AD_GROUPS][D + 4];
  __shared__ float gm[AD_GROUPS], gl[AD_GROUPS];
  __shared__ int last_flag;
  pdl_trigger();
  pdl_wait();  // q and the newest K/V row come from the preceding kernel
  const int h = blockIdx.x, b = blockIdx.y, s = blockIdx.z;
  const int n = seq_lens[b];
  const int npages = (n + KV_PAGE - 1) / KV_PAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 4, l16 = lane & 15;
  const int grp = warp * 2 + sub;
  float* my_part = part + (((size_t)b * H + h) * S + s) * (D + 2);

  if (s < npages) {
    float qf[8];
    unpack8<__half>(ld_cached16(q + ((size_t)b * H + h) * D + l16 * 8), qf);
#pragma unroll
    for (int i = 0; i < 8; ++i) qf[i] *= scale;
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int pi = s; pi < npages; pi += S) {
      const int page = page_table[(size_t)b * max_pages + pi];
      const int cnt = min(KV_PAGE, n - pi * KV_PAGE);
      const __half* kp = kcache + ((size_t)page * H + h) * KV_PAGE * D + l16 * 8;
      const __half* vp = vcache + ((size_t)page * H + h) * KV_PAGE * D + l16 * 8;
      // the whole page (this thread: 8 K and 8 V pieces of 16 bytes) is requested before anything is consumed
      vec8 kv[8], vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = u * AD_GROUPS + grp;  // token of this half-warp in pass u
        kv[u] = (j < cnt) ? ld_stream_rw16(kp + (size_t)j * D) : vec8{0u, 0u, 0u, 0u};
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = u * AD_GROUPS + grp;
        vv[u] = (j < cnt) ? ld_stream_rw16(vp + (size_t)j * D) : vec8{0u, 0u, 0u, 0u};
      }
      asm volatile("" : "+r"(kv[0].x), "+r"(kv[1].x), "+r"(kv[2].x), "+r"(kv[3].x), "+r"(kv[4].x), "+r"(kv[5].x),
                        "+r"(kv[6].x), "+r"(kv[7].x), "+r"(vv[0].x), "+r"(vv[1].x), "+r"(vv[2].x), "+r"(vv[3].x),
                        "+r"(vv[4].x), "+r"(vv[5].x), "+r"(vv[6].x), "+r"(vv[7].x));
      float sc[8], pm = -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float kf[8];
        unpack8<__half>(kv[u], kf);
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) d += qf[i] * kf[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        sc[u] = (u * AD_GROUPS + grp < cnt) ? d : -INFINITY;
        pm = fmaxf(pm, sc[u]);
      }
      if (pm > -INFINITY) {  // (uniform per half-warp; the shuffles above ran converged)
        const float mn = fmaxf(m, pm);
        const float corr = __expf(m - mn);  // m = -inf on the first page: exp(-inf) = 0
        l *= corr;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float pj = __expf(sc[u] - mn);  // masked tokens: exp(-inf) = 0 (their V pieces are zeros)
          l += pj;
          float vf[8];
          unpack8<__half>(vv[u], vf);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += pj * vf[i];
        }
        m = mn;
      }
    }
    // merge the eight half-warps
#pragma unroll
    for (int i = 0; i < 8; ++i) osm[grp][l16 * 8 + i] = acc[i];
    if (l16 == 0) {
      gm[grp] = m;
      gl[grp] = l;
    }
    __syncthreads();
    {
      const int d = threadIdx.x;  // AD_THREADS == D
      float M = -INFINITY;
#pragma unroll
      for (int g2 = 0; g2 < AD_GROUPS; ++g2) M = fmaxf(M, gm[g2]);
      float o = 0.f, den = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < AD_GROUPS; ++g2) {
        const float w = (gm[g2] == -INFINITY) ? 0.f : __expf(gm[g2] - M);
        o += w * osm[g2][d];
        den += w * gl[g2];
      }
      my_part[d] = o;
      if (d == 0) {
        my_part[D] = M;
        my_part[D + 1] = den;
      }
    }
  }
  // ---- the last CTA of this (b, h) combines the partial results of the min(S, npages) active splits
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(&counters[b * H + h], 1);
    last_flag = (arrived == S - 1);
    if (last_flag) counters[b * H + h] = 0;  // ready for the next launch (stream order)
  }
  __syncthreads();
  if (!last_flag) return;
  __threadfence();
  {
    const int d = threadIdx.x;
    const int ns = min(S, npages);
    const float* p = part + ((size_t)b * H + h) * S * (D + 2);
    float M = -INFINITY;
    for (int s2 = 0; s2 < ns; ++s2) M = fmaxf(M, __ldcg(p + s2 * (D + 2) + D));
    float num = 0.f, den = 0.f;
#pragma unroll 4
    for (int s2 = 0; s2 < ns; ++s2) {
      const float w = __expf(__ldcg(p + s2 * (D + 2) + D) - M);
      num += w * __ldcg(p + s2 * (D + 2) + d);
      den += w * __ldcg(p + s2 * (D + 2) + D + 1);
    }
    out[((size_t)b * H + h) * D + d] = __float2half_rn(num / den);
  }
}

SS_API int ss_attn_decode_paged_f16(const void* q, const void* kcache, const void* vcache, const int* seq_lens,
                                    const int* page_table, int max_pages, void* out, float* workspace, int B, int H,
                                    int D, int splits, float scale, void* stream) {
  SS_REQUIRE(D == 128, "decode attention is specialised for head_dim 128");
  SS_REQUIRE(splits >= 1, "splits >= 1");
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  SS_REQUIRE(B * H <= AD_COUNTER_WORDS, "B * H exceeds the arrival-counter block of the workspace");
  int* counters = reinterpret_cast<int*>(workspace);  // first AD_COUNTER_WORDS words; partial results follow
  workspace += AD_COUNTER_WORDS;
  SS_CUDA(ss::launch_pdl(attn_decode_kernel, dim3(H, B, splits), dim3(AD_THREADS), 0, s, (const __half*)q,
                         (const __half*)kcache, (const __half*)vcache, seq_lens, page_table, max_pages, workspace,
                         counters, (__half*)out, H, splits, scale));
  SS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Image-token logits processor + greedy argmax (one CTA per sequence).
// Semantics of AutoImageTokenGenerationProcessor (generation.py:19-31) followed by argmax:
//   last id in img_ids[0 .. n-2]  -> scores[img_ids[idx+1]] = max(scores) + 10   (fp16 add)
//   otherwise                     -> scores[img_ids[1 .. n-1]] = 0.0
// then (optional) transformers' SuppressTokensLogitsProcessor: scores[suppress_ids] = -inf (applied AFTER the image
// processor, i.e. later in the `logits_processor=` list), then argmax with ties resolved to the lowest index.
// The logits row is modified in place exactly as the reference modifies `scores`.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) logits_argmax_kernel(__half* __restrict__ logits, int ld, int V,
                                                             const int* __restrict__ last_ids,
                                                             const int* __restrict__ img_ids, int n_img_ids,
                                                             const int* __restrict__ suppress_ids, int n_suppress,
                                                             int* __restrict__ next_ids) {
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int forced_s;
  const int b = blockIdx.x;
  __half* row = logits + (size_t)b * ld;
  if (threadIdx.x == 0) {
    int forced = -1;
    if (img_ids != nullptr) {
      const int last = last_ids[b];
      for (int i = 0; i + 1 < n_img_ids; ++i)
        if (img_ids[i] == last) {
          forced = img_ids[i + 1];
          break;
        }
    }
    forced_s = forced;
  }
  __syncthreads();
  const int forced = forced_s;
  if (img_ids != nullptr && forced < 0) {
    if (threadIdx.x >= 1 && threadIdx.x < n_img_ids) row[img_ids[threadIdx.x]] = __float2half_rn(0.f);
    __syncthreads();
  }
  if (forced < 0 && n_suppress > 0) {  // (forced: the image processor's max is taken over the unsuppressed row)
    if (threadIdx.x < n_suppress) row[suppress_ids[threadIdx.x]] = __float2half_rn(-INFINITY);
    __syncthreads();
  }
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = __half2float(row[i]);
    if (v > best || (v == best && i < besti)) {
      best = v;
      besti = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) {
      best = ov;
      besti = oi;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    red_v[wid] = best;
    red_i[wid] = besti;
  }
  __syncthreads();
  if (wid == 0) {
    best = red_v[lane];
    besti = red_i[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) {
        best = ov;
        besti = oi;
      }
    }
    if (lane == 0) {
      if (forced >= 0) {
        row[forced] = __hadd(__float2half_rn(best), __float2half_rn(10.f));
        next_ids[b] = forced;  // max+10 beats every other entry
        for (int i = 0; i < n_suppress; ++i) row[suppress_ids[i]] = __float2half_rn(-INFINITY);
      } else {
        next_ids[b] = besti;
      }
    }
  }
}

SS_API int ss_logits_process_argmax_f16(void* logits, int ld, int V, const int* last_ids, const int* img_ids,
                                        int n_img_ids, const int* suppress_ids, int n_suppress, int* next_ids, int B,
                                        void* stream) {
  SS_REQUIRE(n_img_ids <= 1024 && n_suppress <= 1024, "image-token / suppress list too long");
  if (B == 0) return 0;
  logits_argmax_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>((__half*)logits, ld, V, last_ids, img_ids, n_img_ids,
                                                             suppress_ids, suppress_ids ? n_suppress : 0, next_ids);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Embedding gather (rows of 16-bit values) and decode-state advance.
// ---------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const uint16_t* __restrict__ table, const int* __restrict__ ids,
                                   uint16_t* __restrict__ out, int ld_out, int width) {
  const int tok = blockIdx.x;
  const vec8* src = reinterpret_cast<const vec8*>(table + (size_t)ids[tok] * width);
  vec8* dst = reinterpret_cast<vec8*>(out + (size_t)tok * ld_out);
  for (int i = threadIdx.x; i < (width >> 3); i += blockDim.x) dst[i] = src[i];
}

SS_API int ss_gather_rows_16b(const void* table, const int* ids, void* out, int ld_out, int ntok, int width,
                              void* stream) {
  SS_REQUIRE(width % 8 == 0 && ld_out % 8 == 0, "row width must be a multiple of 8 elements");
  if (ntok == 0) return 0;
  gather_rows_kernel<<<ntok, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)table, ids, (uint16_t*)out, ld_out,
                                                            width);
  SS_LAUNCH_CHECK();
  return 0;
}

// After a decode step: last_id <- next_id; position, slot and length advance by one for live sequences;
// the emitted id is appended to the per-sequence output ring and the final-norm hidden row is kept.
__global__ void decode_advance_kernel(const int* __restrict__ next_ids, int* __restrict__ cur_ids,
                                      int* __restrict__ tok_pos, int* __restrict__ tok_slot,
                                      int* __restrict__ seq_lens, int* __restrict__ out_ids, int out_cap,
                                      int* __restrict__ n_out, int* __restrict__ done, int eos_id, int B,
                                      const int* __restrict__ schedule, int sched_cap) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (done[b]) return;
  int id = next_ids[b];
  const int k = n_out[b];
  // optional forced schedule (the reference exposes the same hook as `logits_processor=`, models.py:105,115):
  // entry >= 0 overrides the greedy choice for generated index k (used to force <img>/EOS with synthetic weights)
  if (schedule != nullptr && k < sched_cap && schedule[(size_t)b * sched_cap + k] >= 0)
    id = schedule[(size_t)b * sched_cap + k];
  if (k < out_cap) out_ids[(size_t)b * out_cap + k] = id;
  n_out[b] = k + 1;
  cur_ids[b] = id;
  tok_pos[b] += 1;
  tok_slot[b] += 1;
  seq_lens[b] += 1;
  if (id == eos_id) done[b] = 1;
}

SS_API int ss_decode_advance(const int* next_ids, int* cur_ids, int* tok_pos, int* tok_slot, int* seq_lens,
                             int* out_ids, int out_cap, int* n_out, int* done, int eos_id, int B,
                             const int* schedule, int sched_cap, void* stream) {
  if (B == 0) return 0;
  SS_REQUIRE(B <= 32, "at most 32 sequences per rank");
  decode_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(next_ids, cur_ids, tok_pos, tok_slot, seq_lens, out_ids,
                                                            out_cap, n_out, done, eos_id, B, schedule, sched_cap);
  SS_LAUNCH_CHECK();
  return 0;
}

// dst[b, idx[b], :] = src[b, :] with a device-side row index (keeps the per-step final-norm hidden state that
// ContinuousLVLM.generate slices at models.py:182-197 while the decode step stays CUDA-graph replayable).
__global__ void store_rows_indexed_kernel(const uint16_t* __restrict__ src, int ld_src, uint16_t* __restrict__ dst,
                                          int cap, const int* __restrict__ idx, int width) {
  const int b = blockIdx.x;
  const int r = idx[b];
  if (r < 0 || r >= cap) return;
  const vec8* s = reinterpret_cast<const vec8*>(src + (size_t)b * ld_src);
  vec8* d = reinterpret_cast<vec8*>(dst + ((size_t)b * cap + r) * width);
  for (int v = threadIdx.x; v < (width >> 3); v += blockDim.x) d[v] = s[v];
}
SS_API int ss_store_rows_indexed_16b(const void* src, int ld_src, void* dst, int cap, const int* idx, int B, int width,
                                     void* stream) {
  SS_REQUIRE(width % 8 == 0 && ld_src % 8 == 0, "width % 8");
  if (B == 0) return 0;
  store_rows_indexed_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, ld_src, (uint16_t*)dst, cap, idx,
                                                                width);
  SS_LAUNCH_CHECK();
  return 0;
}

// W' = W + scaling * B A  (peft LoRA merge: fp32 accumulate, ONE rounding to fp16 — SURVEY.md §7 "LoRA").
__global__ void lora_merge_kernel(const __half* __restrict__ W, const __half* __restrict__ A,
                                  const __half* __restrict__ Bm, __half* __restrict__ out, int N, int K, int r,
                                  float scaling) {
  extern __shared__ float brow[];  // B[n, :] for this row
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < r; i += blockDim.x) brow[i] = __half2float(Bm[(size_t)n * r + i]);
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < r; ++i) acc += brow[i] * __half2float(A[(size_t)i * K + k]);
    out[(size_t)n * K + k] = __float2half_rn(__half2float(W[(size_t)n * K + k]) + scaling * acc);
  }
}
SS_API int ss_lora_merge_f16(const void* W, const void* A, const void* B, void* out, int N, int K, int r,
                             float scaling, void* stream) {
  if (N == 0) return 0;
  lora_merge_kernel<<<N, 256, r * sizeof(float), (cudaStream_t)stream>>>((const __half*)W, (const __half*)A,
                                                                        (const __half*)B, (__half*)out, N, K, r, scaling);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// KV compaction for the window / multimodal attention-sink policy (src/inference/vis_george_sink.py:266-291):
// token slot i of the destination page list receives the K and V rows of source slot src_idx[i], for every
// layer.  Afterwards the sequence is again a dense [sink tokens..., live tokens...] run of pages, so the
// attention kernels need no per-token mask.  Pools: [layer][page][H][64][D].
// ---------------------------------------------------------------------------------------------
__global__ void kv_gather_tokens_kernel(uint16_t* __restrict__ kpool, uint16_t* __restrict__ vpool,
                                        long long layer_stride, const int* __restrict__ src_pages,
                                        const int* __restrict__ dst_pages, const int* __restrict__ src_idx, int H,
                                        int D) {
  const int i = blockIdx.x, layer = blockIdx.y;
  const int s = src_idx[i];
  const size_t src = (size_t)layer * layer_stride + ((size_t)src_pages[s / KV_PAGE] * H * KV_PAGE + (s % KV_PAGE)) * D;
  const size_t dst = (size_t)layer * layer_stride + ((size_t)dst_pages[i / KV_PAGE] * H * KV_PAGE + (i % KV_PAGE)) * D;
  const int vec_per_row = D >> 3;
  for (int t = threadIdx.x; t < H * vec_per_row; t += blockDim.x) {
    const int h = t / vec_per_row, v = t % vec_per_row;
    const size_t off = (size_t)h * KV_PAGE * D + v * 8;
    *reinterpret_cast<vec8*>(kpool + dst + off) = *reinterpret_cast<const vec8*>(kpool + src + off);
    *reinterpret_cast<vec8*>(vpool + dst + off) = *reinterpret_cast<const vec8*>(vpool + src + off);
  }
}
SS_API int ss_kv_gather_tokens_16b(void* kpool, void* vpool, int layers, long long layer_stride, const int* src_pages,
                                   const int* dst_pages, const int* src_idx, int n, int H, int D, void* stream) {
  SS_REQUIRE(D % 8 == 0, "head_dim % 8");
  if (n == 0) return 0;
  kv_gather_tokens_kernel<<<dim3(n, layers), 256, 0, (cudaStream_t)stream>>>((uint16_t*)kpool, (uint16_t*)vpool,
                                                                              layer_stride, src_pages, dst_pages,
                                                                              src_idx, H, D);
  SS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// past_key_values ingestion: one layer's K and V given as [H, n, D] tensors (row pitch in elements per head and
// per token, the layout of the reference's tuple-of-(K, V) cache, modeling_llama_xformer.py:241-244) are written into
// token slots 0..n-1 of the destination page list.  Used when a caller hands `generate(past_key_values=...)` a
// sliced / concatenated cache as src/inference/vis_george_sink.py:266-291 builds it.
// ---------------------------------------------------------------------------------------------
__global__ void kv_scatter_tokens_kernel(uint16_t* __restrict__ kpool, uint16_t* __restrict__ vpool,
                                         const uint16_t* __restrict__ ksrc, const uint16_t* __restrict__ vsrc,
                                         long long k_sh, long long k_st, long long v_sh, long long v_st,
                                         const int* __restrict__ dst_pages, int H, int D) {
  const int i = blockIdx.x;
  const size_t dst = ((size_t)dst_pages[i / KV_PAGE] * H * KV_PAGE + (i % KV_PAGE)) * D;
  const int vec_per_row = D >> 3;
  for (int t = threadIdx.x; t < H * vec_per_row; t += blockDim.x) {
    const int h = t / vec_per_row, v = t % vec_per_row;
    const size_t off = (size_t)h * KV_PAGE * D + v * 8;
    *reinterpret_cast<vec8*>(kpool + dst + off) = *reinterpret_cast<const vec8*>(ksrc + h * k_sh + i * k_st + v * 8);
    *reinterpret_cast<vec8*>(vpool + dst + off) = *reinterpret_cast<const vec8*>(vsrc + h * v_sh + i * v_st + v * 8);
  }
}
SS_API int ss_kv_scatter_tokens_16b(void* kpool_layer, void* vpool_layer, const void* k_src, const void* v_src,
                                    long long k_sh, long long k_st, long long v_sh, long long v_st,
                                    const int* dst_pages, int n, int H, int D, void* stream) {
  SS_REQUIRE(D % 8 == 0 && k_sh % 8 == 0 && k_st % 8 == 0 && v_sh % 8 == 0 && v_st % 8 == 0,
             "K/V rows must be 16-byte aligned (head_dim and pitches multiples of 8 elements)");
  SS_REQUIRE(((reinterpret_cast<uintptr_t>(k_src) | reinterpret_cast<uintptr_t>(v_src)) & 15) == 0, "K/V base alignment");
  if (n == 0) return 0;
  kv_scatter_tokens_kernel<<<n, 256, 0, (cudaStream_t)stream>>>((uint16_t*)kpool_layer, (uint16_t*)vpool_layer,
                                                               (const uint16_t*)k_src, (const uint16_t*)v_src, k_sh, k_st,
                                                               v_sh, v_st, dst_pages, H, D);
  SS_LAUNCH_CHECK();
  return 0;
}
