// Fused multi-head attention, fp16, head_dim 64 or 128, online softmax (scores never leave the SM).
// One kernel serves
//   * Llama prefill / chunked decode: causal from the bottom-right corner
//     (xops LowerTriangularFromBottomRightMask, modeling_llama_xformer.py:289-295), K/V read through
//     the page table (window + attention-sink pages),
//   * ViT self-attention (qwen_visual.py:184-235; head_dim 104 zero-padded to 128 by the host packer),
//   * the agent / attn-pool Resampler MHA (qwen_visual.py:138-150), PerceiverAttention and
//     AttentionPool2d (resampler.py:47-76, 90-118), the SDXL UNet self/cross attention (head_dim 64).
// Tensors are addressed with explicit (batch, token, head) strides so fused QKV buffers need no copies.
//
// v1 uses mma.sync m16n8k16 (64 query rows x 64 keys per CTA iteration); the tcgen05 version with S/P
// in TMEM is the planned replacement.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int FM_BM = 64, FM_BN = 64, FM_THREADS = 128;

struct FmhaParams {
  const __half *q, *k, *v;
  __half* o;
  long long q_sb, q_sl, q_sh, k_sb, k_sl, k_sh, v_sb, v_sl, v_sh, o_sb, o_sl, o_sh;
  int B, H, Lq, Lk;
  const int* kv_lens;     // optional device override of Lk per batch entry
  const int* page_table;  // optional [B, max_pages]; then k/v are page pools [page][H][64][D]
  int max_pages;
  float scale_log2;       // scale * log2(e)
  int causal;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// swizzled tile: row r, 16-byte chunk c of a [rows][D] fp16 tile
template <int D>
__device__ __forceinline__ __half* tile_ptr(__half* base, int r, int c) {
  return base + ((size_t)r * (D / 8) + (c ^ (r & 7))) * 8;
}

template <int D>
__device__ __forceinline__ void load_tile(__half* smem, const __half* g, long long row_stride, int row0, int nrows_valid,
                                          int tid) {
  // 64 rows x D/8 chunks
  constexpr int CH = D / 8;
#pragma unroll
  for (int i = 0; i < (64 * CH) / FM_THREADS; ++i) {
    const int idx = tid + i * FM_THREADS;
    const int r = idx / CH, c = idx % CH;
    const bool ok = (row0 + r) < nrows_valid;
    const __half* src = g + (long long)(ok ? row0 + r : 0) * row_stride + c * 8;
    cp_async16(tile_ptr<D>(smem, r, c), src, ok);
  }
}

template <int D, bool PAGED>
__global__ void __launch_bounds__(FM_THREADS) fmha_kernel(const FmhaParams p) {
  extern __shared__ __align__(128) uint8_t fm_smem[];
  __half* sQ = reinterpret_cast<__half*>(fm_smem);
  __half* sK = sQ + 64 * D;       // 2 buffers
  __half* sV = sK + 2 * 64 * D;   // 2 buffers

  pdl_trigger();
  pdl_wait();  // q/k/v (and kv_lens / page tables) come from preceding kernels
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * FM_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int Lk = p.kv_lens ? p.kv_lens[b] : p.Lk;
  const int Lq = p.Lq;
  const int shift = Lk - Lq;  // bottom-right alignment of the causal diagonal

  const __half* qg = p.q + b * p.q_sb + h * p.q_sh;
  const __half* kg = PAGED ? p.k : p.k + b * p.k_sb + h * p.k_sh;
  const __half* vg = PAGED ? p.v : p.v + b * p.v_sb + h * p.v_sh;
  const int* pt = PAGED ? p.page_table + (size_t)b * p.max_pages : nullptr;

  int n_end = Lk;
  if (p.causal) n_end = min(Lk, m0 + FM_BM + shift);
  const int ntiles = (n_end + FM_BN - 1) / FM_BN;

  auto issue_kv = [&](int tile, int buf) {
    if (PAGED) {
      const int page = pt[tile];
      const __half* kp = kg + ((size_t)page * p.H + h) * 64 * D;
      const __half* vp = vg + ((size_t)page * p.H + h) * 64 * D;
      load_tile<D>(sK + buf * 64 * D, kp, D, 0, Lk - tile * 64, tid);
      load_tile<D>(sV + buf * 64 * D, vp, D, 0, Lk - tile * 64, tid);
    } else {
      load_tile<D>(sK + buf * 64 * D, kg, p.k_sl, tile * FM_BN, Lk, tid);
      load_tile<D>(sV + buf * 64 * D, vg, p.v_sl, tile * FM_BN, Lk, tid);
    }
  };

  load_tile<D>(sQ, qg, p.q_sl, m0, Lq, tid);
  if (ntiles > 0) issue_kv(0, 0);
  cp_async_commit();

  float o_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float row_m[2] = {-INFINITY, -INFINITY}, row_l[2] = {0.f, 0.f};
  uint32_t qf[D / 16][4];

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) issue_kv(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (tile == 0) {
#pragma unroll
      for (int kt = 0; kt < D / 16; ++kt) {
        const int mi = lane >> 3;
        const int r = warp * 16 + (lane & 7) + ((mi & 1) ? 8 : 0);
        ldsm_x4(qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3], tile_ptr<D>(sQ, r, 2 * kt + (mi >> 1)));
      }
    }
    const __half* cK = sK + buf * 64 * D;
    const __half* cV = sV + buf * 64 * D;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < D / 16; ++kt) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key tiles
        const int mi = lane >> 3;
        const int r = jp * 16 + (lane & 7) + ((mi >> 1) ? 8 : 0);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, tile_ptr<D>(const_cast<__half*>(cK), r, 2 * kt + (mi & 1)));
        mma_16816(s[2 * jp], qf[kt], b0, b1);
        mma_16816(s[2 * jp + 1], qf[kt], b2, b3);
      }
    }
    // ---- mask + online softmax ----
    const int key0 = tile * FM_BN;
    const int qrow0 = m0 + warp * 16 + g;  // rows qrow0 and qrow0+8
    const bool need_mask = (key0 + FM_BN > Lk) || (p.causal && (key0 + FM_BN - 1 > m0 + warp * 16 + shift));
    if (need_mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + j * 8 + 2 * t + (e & 1);
          const int qr = qrow0 + ((e >> 1) ? 8 : 0);
          const bool ok = key < Lk && (!p.causal || key <= qr + shift);
          if (!ok) s[j][e] = -INFINITY;
        }
      }
    }
    float mx[2] = {row_m[0], row_m[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      msc[i] = (mx[i] == -INFINITY) ? 0.f : mx[i] * p.scale_log2;
      corr[i] = (row_m[i] == -INFINITY) ? 0.f : exp2f(row_m[i] * p.scale_log2 - msc[i]);
      row_m[i] = mx[i];
      row_l[i] *= corr[i];
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] * p.scale_log2 - msc[0]);
      s[j][1] = exp2f(s[j][1] * p.scale_log2 - msc[0]);
      s[j][2] = exp2f(s[j][2] * p.scale_log2 - msc[1]);
      s[j][3] = exp2f(s[j][3] * p.scale_log2 - msc[1]);
      ls[0] += s[j][0] + s[j][1];
      ls[1] += s[j][2] + s[j][3];
    }
    row_l[0] += ls[0];
    row_l[1] += ls[1];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {  // 16 keys each
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kt][0], s[2 * kt][1]);
      pa[1] = pack_h2(s[2 * kt][2], s[2 * kt][3]);
      pa[2] = pack_h2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      pa[3] = pack_h2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {  // pairs of 8-wide d tiles
        const int mi = lane >> 3;
        const int r = kt * 16 + (lane & 7) + ((mi & 1) ? 8 : 0);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, tile_ptr<D>(const_cast<__half*>(cV), r, 2 * np + (mi >> 1)));
        mma_16816(o_acc[2 * np], pa, b0, b1);
        mma_16816(o_acc[2 * np + 1], pa, b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  cp_async_wait<0>();

  // ---- finalize ----
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    row_l[i] += __shfl_xor_sync(0xffffffffu, row_l[i], 1);
    row_l[i] += __shfl_xor_sync(0xffffffffu, row_l[i], 2);
  }
  const float inv[2] = {row_l[0] > 0.f ? 1.f / row_l[0] : 0.f, row_l[1] > 0.f ? 1.f / row_l[1] : 0.f};
  __half* og = p.o + b * p.o_sb + h * p.o_sh;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qr = m0 + warp * 16 + g + i * 8;
    if (qr < Lq) {
      __half* orow = og + (long long)qr * p.o_sl;
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        const uint32_t v = pack_h2(o_acc[nt][2 * i] * inv[i], o_acc[nt][2 * i + 1] * inv[i]);
        *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * t) = v;
      }
    }
  }
}

template <int D, bool PAGED>
int launch_fmha(const FmhaParams& p, cudaStream_t s) {
  const int smem = 5 * 64 * D * 2;
  static bool attr = false;
  if (!attr) {
    SS_CUDA(cudaFuncSetAttribute(fmha_kernel<D, PAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  dim3 grid((p.Lq + FM_BM - 1) / FM_BM, p.H, p.B);
  SS_CUDA(ss::launch_pdl(fmha_kernel<D, PAGED>, grid, dim3(FM_THREADS), (size_t)smem, s, p));
  SS_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int ss_internal_fmha_tc_paged(const void* q, const void* kpool, const void* vpool, void* out, int B, int H, int Lq, int Lk,
                              int D, long long q_sb, long long q_sl, long long q_sh, long long o_sb, long long o_sl,
                              long long o_sh, const int* page_table, int max_pages, float scale, int causal,
                              cudaStream_t stream);
int ss_internal_fmha_tc(const void* q, const void* k, const void* v, void* out, int B, int H, int Lq, int Lk, int D,
                        long long q_sb, long long q_sl, long long q_sh, long long k_sb, long long k_sl, long long k_sh,
                        long long v_sb, long long v_sl, long long v_sh, long long o_sb, long long o_sl, long long o_sh,
                        float scale, int causal, cudaStream_t stream);

static constexpr int fmha_tc_min_lk() { return 65; }  // short key sets (UNet cross-attention: 64 context tokens) fit one mma.sync tile

// which kernel family served each ss_fmha_f16 call (test hook: a tcgen05 path that silently declines a layout
// would otherwise hide behind the mma.sync kernel's correct results)
static std::atomic<long long> g_fmha_tc_calls{0}, g_fmha_mma_calls{0};

SS_API int ss_fmha_path_counts(long long* tc_calls, long long* mma_calls, int reset) {
  if (tc_calls) *tc_calls = g_fmha_tc_calls.load();
  if (mma_calls) *mma_calls = g_fmha_mma_calls.load();
  if (reset) {
    g_fmha_tc_calls = 0;
    g_fmha_mma_calls = 0;
  }
  return 0;
}

SS_API int ss_fmha_f16(const void* q, const void* k, const void* v, void* out, int B, int H, int Lq, int Lk, int D,
                       long long q_sb, long long q_sl, long long q_sh, long long k_sb, long long k_sl, long long k_sh,
                       long long v_sb, long long v_sl, long long v_sh, long long o_sb, long long o_sl, long long o_sh,
                       const int* kv_lens, const int* page_table, int max_pages, float scale, int causal,
                       void* stream) {
  SS_REQUIRE(D == 64 || D == 128, "head_dim must be 64 or 128 (pad 104 -> 128 on the host)");
  SS_REQUIRE(q_sl % 8 == 0 && k_sl % 8 == 0 && v_sl % 8 == 0 && o_sl % 2 == 0, "token strides must keep 16-byte rows");
  SS_REQUIRE(q_sh % 8 == 0 && k_sh % 8 == 0 && v_sh % 8 == 0, "head strides must keep 16-byte rows");
  if (B == 0 || H == 0 || Lq == 0) return 0;
  if (page_table == nullptr && kv_lens == nullptr && Lq >= 64 && Lk >= fmha_tc_min_lk() && k_sb == v_sb) {
    // tcgen05 path (S/PV accumulators in TMEM); -1 = layout not expressible as row-matrix views
    const int rc = ss_internal_fmha_tc(q, k, v, out, B, H, Lq, Lk, D, q_sb, q_sl, q_sh, k_sb, k_sl, k_sh, v_sb, v_sl, v_sh,
                                       o_sb, o_sl, o_sh, scale, causal, (cudaStream_t)stream);
    if (rc >= 0) {
      if (rc == 0) g_fmha_tc_calls++;
      return rc;
    }
  }
  if (page_table != nullptr && kv_lens == nullptr && D == 128 && Lq >= 64 && Lk >= fmha_tc_min_lk()) {
    // paged K/V (the Llama prefill and the 66-token image-run chunk) on the tcgen05 kernel: pages through 3-D TMA maps
    const int rc = ss_internal_fmha_tc_paged(q, k, v, out, B, H, Lq, Lk, D, q_sb, q_sl, q_sh, o_sb, o_sl, o_sh, page_table,
                                             max_pages, scale, causal, (cudaStream_t)stream);
    if (rc >= 0) {
      if (rc == 0) g_fmha_tc_calls++;
      return rc;
    }
  }
  g_fmha_mma_calls++;
  FmhaParams p;
  p.q = (const __half*)q;
  p.k = (const __half*)k;
  p.v = (const __half*)v;
  p.o = (__half*)out;
  p.q_sb = q_sb; p.q_sl = q_sl; p.q_sh = q_sh;
  p.k_sb = k_sb; p.k_sl = k_sl; p.k_sh = k_sh;
  p.v_sb = v_sb; p.v_sl = v_sl; p.v_sh = v_sh;
  p.o_sb = o_sb; p.o_sl = o_sl; p.o_sh = o_sh;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.kv_lens = kv_lens;
  p.page_table = page_table;
  p.max_pages = max_pages;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  cudaStream_t s = (cudaStream_t)stream;
  if (page_table) {
    if (D == 64) return launch_fmha<64, true>(p, s);
    return launch_fmha<128, true>(p, s);
  }
  if (D == 64) return launch_fmha<64, false>(p, s);
  return launch_fmha<128, false>(p, s);
}
