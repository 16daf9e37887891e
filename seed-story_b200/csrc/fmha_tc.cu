// Fused attention on the 5th-gen tensor cores (tcgen05 + TMEM), fp16, head_dim 64 / 128.
//   S = Q K^T     tcgen05.mma  M=128 (queries) x N=128 (keys) x K=D, operands from TMA-staged shared memory,
//                 accumulator in TMEM (two S buffers: the MMA of tile j+1 overlaps the softmax of tile j)
//   softmax       256 threads: a query row is shared by a PAIR of threads (warps w and w+4 see the same TMEM lanes),
//                 each owning 64 of the tile's 128 keys and half of the output dims, so a thread reads its S slice
//                 once (tcgen05.ld 32x32b), exchanges the row max with its partner through shared memory, and
//                 writes its fp16 P slice (= one 128B-swizzled 64-key atom) for the next MMA
//   O += P V      tcgen05.mma  M=128 x N=D x K=128 with V as an MN-major B operand straight from the [key][d]
//                 layout TMA delivers.  O (and the row sums L = P 1, a 16-column MMA against a block of ones)
//                 ACCUMULATE IN TMEM across all key tiles: TMEM reads run at 64 B/clk/SM, so reading the 128x128
//                 fp32 S tile already costs 1024 clk per tile and a per-tile read-back of P V would add 50-100 %.
//                 The reference maximum of a row is only raised when the new tile maximum exceeds it by more
//                 than 2^8 (exact arithmetic is unchanged: P and L use the same reference); only then is the
//                 accumulator read, rescaled and written back — rare after the first tiles.
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..9 = softmax + output.
// Same mask semantics as fmha.cu: causal diagonal anchored bottom-right (xformers LowerTriangularFromBottomRightMask,
// modeling_llama_xformer.py:289-295), keys >= Lk masked.  Operands are row-matrix views (token rows, heads side by
// side in a row) so fused QKV buffers are consumed in place.
#include <cstdlib>
#include <cstring>

#include "tc.cuh"

namespace {

constexpr int FT_BM = 128, FT_BN = 128, FT_THREADS = 576;  // TMA + MMA warps, 16 softmax warps

struct FtParams {
  __half* o;
  long long o_sb, o_sl, o_sh;
  int H, Lq, Lk;
  int q_rows_per_batch, k_rows_per_batch;  // 0 => operand is broadcast over the batch
  int q_col_per_head, k_col_per_head, v_col_per_head;
  int q_col0, k_col0, v_col0;
  float scale_log2;
  int causal;
  int stages, poly;
  // paged K/V (the Llama cache: pages [page][H][64][D]): tmK / tmV are 3-D maps (D, 64 tokens, page * H + head) and a
  // 128-key tile is two pages from this sequence's page-table row
  const int* page_table;
  int paged, max_pages;
};

// 2^x for x <= 0 on the FMA / integer pipes: round-to-nearest split x = xi + xf (magic-number add), a degree-3
// minimax polynomial for 2^xf on [-0.5, 0.5] (max relative error 7.5e-5, below half an fp16 ulp — the result is
// rounded to fp16 right away), and xi added straight into the exponent field.  Inputs below -126 (masked keys are
// -inf) clamp to 2^-126, which rounds to 0 in fp16.
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the low mantissa bits of t now hold round(x)
  const float xf = x - (t - 12582912.f);
  float pl = fmaf(xf, 0.0551716685f, 0.2426111251f);
  pl = fmaf(pl, xf, 0.6932609677f);
  pl = fmaf(pl, xf, 0.9999280572f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}

template <int D>
struct FtSmem {
  static constexpr int ATOM = 128 * 128;            // [128 rows][64 x 16-bit] = 16 KB
  static constexpr int Q_BYTES = (D / 64) * ATOM;
  static constexpr int KV_BYTES = (D / 64) * ATOM;  // K tile or V tile
  static constexpr int P_BYTES = 2 * ATOM;          // one P tile; D = 64 has the shared memory for two
  static constexpr int P_BUFS = (D == 64) ? 2 : 1;
  // K/V stages: a tile's K/V can only be requested once the P V product two (STAGES) tiles back has retired, so two
  // stages leave the TMA latency exposed on every tile; head_dim 64 has the shared memory for four
  static constexpr int STAGES = (D == 64) ? 4 : 2;
  static constexpr int ONES_BYTES = 4096;           // [16][128] fp16 ones: B operand of the row-sum MMA
  static constexpr int TOTAL = Q_BYTES + STAGES * 2 * KV_BYTES + P_BUFS * P_BYTES + ONES_BYTES + 1024 + 256 + 2048 /*row-max exchange*/ +
                               1024 /*final per-group maxima (ping-pong kernel)*/;
};

// MN-major B operand (V as [key][d] rows of 128 bytes, 128B swizzle): SBO = 8 key rows * 128 B, LBO = distance between
// 64-wide d atoms (128 keys * 128 B)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int D>
__global__ void __launch_bounds__(FT_THREADS, 1) fmha_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmK,
                                                                const __grid_constant__ CUtensorMap tmV,
                                                                const FtParams p) {
  using S = FtSmem<D>;
  extern __shared__ uint8_t ft_smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + S::Q_BYTES;                       // stage s: K at s*2*KV, V right after
  uint8_t* sP = sKV + S::STAGES * 2 * S::KV_BYTES;
  uint8_t* sOnes = sP + S::P_BUFS * S::P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + S::ONES_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // [4]
  uint64_t* kv_empty = bars + 5;    // [4]
  uint64_t* s_full = bars + 9;      // [2]
  uint64_t* s_empty = bars + 11;    // [2]
  uint64_t* p_full = bars + 13;     // 1
  uint64_t* pv_full = bars + 14;    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
  float* xchg = reinterpret_cast<float*>(bars + 20);  // [4 parts][128 rows] row-max exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * FT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int Lq = p.Lq, Lk = p.Lk;
  const int shift = Lk - Lq;
  int n_end = Lk;
  if (p.causal) n_end = min(Lk, m0 + FT_BM + shift);
  const int ntiles = (n_end + FT_BN - 1) / FT_BN;

  for (int i = threadIdx.x; i < S::ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;
  tc::fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      tc::mbar_init(q_full, 1);
      for (int i = 0; i < 4; ++i) {
        tc::mbar_init(&kv_full[i], 1);
        tc::mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&s_full[i], 1);
        tc::mbar_init(&s_empty[i], 16);
      }
      tc::mbar_init(p_full, 16);
      tc::mbar_init(&pv_full[0], 1);
      tc::mbar_init(&pv_full[1], 1);
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tmem_ptr_smem, 512);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();  // Q/K/V come from the preceding kernel; O may still be read by an earlier one
  // The whole 512-column TMEM of the SM is allocated (one CTA per SM), so the base is column 0 / lane 0.  Using the
  // constant keeps every TMEM address warp-uniform: with an address loaded from shared memory the compiler wrapped
  // each tcgen05.mma of the single issuing thread in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, and with ~20 small
  // MMAs per key tile that issue overhead was the critical path of the kernel.
  if (*tmem_ptr_smem != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;
  // TMEM columns: S0 | S1 (128 each) | O accumulator (D columns at 256) | L = row sums of P (16 columns at 384)
  const int nst = min(S::STAGES, p.stages);  // K/V ring depth (p.stages is fixed by the host launcher)
  constexpr int PB = S::P_BUFS;
  const uint32_t tmem_S0 = tmem_base, tmem_O = tmem_base + 256, tmem_L = tmem_base + 384;

  if (ntiles == 0) {
    // nothing visible (only possible for degenerate causal shapes): write zeros
    if (warp >= 2 && warp < 6) {
      const int r = (warp & 3) * 32 + lane;
      if (m0 + r < Lq) {
        __half* orow = p.o + b * p.o_sb + (long long)(m0 + r) * p.o_sl + h * p.o_sh;
        for (int d = 0; d < D; ++d) orow[d] = __float2half_rn(0.f);
      }
    }
  } else if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int qrow0 = b * p.q_rows_per_batch + m0;
      tc::mbar_expect_tx(q_full, S::Q_BYTES);
#pragma unroll
      for (int a = 0; a < D / 64; ++a)
        tc::tma_load_2d(sQ + a * S::ATOM, &tmQ, q_full, p.q_col0 + h * p.q_col_per_head + a * 64, qrow0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % nst;
        const uint32_t ph = (j / nst) & 1;
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sK = sKV + s * 2 * S::KV_BYTES;
        uint8_t* sV = sK + S::KV_BYTES;
        const int krow0 = b * p.k_rows_per_batch + j * FT_BN;
        tc::mbar_expect_tx(&kv_full[s], 2 * S::KV_BYTES);
        if (p.paged) {
          // two 64-token pages per tile; a page past the end of the sequence is replaced by the last valid one (its
          // keys are >= Lk and masked; real data keeps the P V product free of stale-memory NaNs)
          const int npages = (Lk + 63) >> 6;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int pi = min(2 * j + half, npages - 1);
            const int row = p.page_table[(size_t)b * p.max_pages + pi] * p.H + h;
#pragma unroll
            for (int a = 0; a < D / 64; ++a) {
              tc::tma_load_3d(sK + a * S::ATOM + half * (64 * 128), &tmK, &kv_full[s], a * 64, 0, row);
              tc::tma_load_3d(sV + a * S::ATOM + half * (64 * 128), &tmV, &kv_full[s], a * 64, 0, row);
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < D / 64; ++a) {
            tc::tma_load_2d(sK + a * S::ATOM, &tmK, &kv_full[s], p.k_col0 + h * p.k_col_per_head + a * 64, krow0);
            tc::tma_load_2d(sV + a * S::ATOM, &tmV, &kv_full[s], p.v_col0 + h * p.v_col_per_head + a * 64, krow0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp runs converged, one elected lane issues =================
    {
      constexpr uint32_t idesc_qk = tc::make_idesc(0, FT_BM, FT_BN);               // A, B K-major
      constexpr uint32_t idesc_pv = tc::make_idesc(0, FT_BM, D) | (1u << 16);      // B (V) MN-major
      const uint32_t aQ = tc::smem_u32(sQ), aP = tc::smem_u32(sP), aOnes = tc::smem_u32(sOnes);
      constexpr uint32_t idesc_l = tc::make_idesc(0, FT_BM, 16);
      auto issue_qk = [&](int j) {
        const int s = j % nst;
        tc::mbar_wait(&kv_full[s], (j / nst) & 1);
        tc::mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t aK = tc::smem_u32(sKV + s * 2 * S::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * S::ATOM + (kk & 3) * 32;
          tc::mma_f16_ss_warp(tmem_S0 + (j & 1) * 128, tc::make_desc_sw128(aQ + off), tc::make_desc_sw128(aK + off), idesc_qk,
                         kk != 0);
        }
        tc::mma_commit_warp(&s_full[j & 1]);
      };
      tc::mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_qk(j + 1);  // S_{j+1} is computed while the softmax works on S_j
        tc::mbar_wait(p_full, j & 1);  // P_j is in shared memory
        tc::fence_after_sync();
        const int s = j % nst;
        const uint32_t aV = tc::smem_u32(sKV + s * 2 * S::KV_BYTES + S::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < FT_BN / 16; ++kk) {
          const uint64_t da = tc::make_desc_sw128(aP + (j % PB) * S::P_BYTES + (kk >> 2) * S::ATOM + (kk & 3) * 32);
          const uint64_t db = make_desc_mn_sw128(aV + kk * 16 * 128, S::ATOM);
          tc::mma_f16_ss_warp(tmem_O, da, db, idesc_pv, (j | kk) != 0);
          // row sums of P_j by the tensor core: P (128 x 16 keys) times a 16 x 16 block of ones
          tc::mma_f16_ss_warp(tmem_L, da, tc::make_desc_sw128(aOnes + (kk >> 2) * 2048 + (kk & 3) * 32), idesc_l,
                              (j | kk) != 0);
        }
        tc::mma_commit_warp(&kv_empty[s]);  // K_j / V_j no longer needed
        tc::mma_commit_warp(&pv_full[j & 1]);
      }
    }
  } else {
    // ================= softmax + output: FOUR threads per query row =================
    // 16 warps (four per scheduler) hide the TMEM-load / exchange / barrier latencies of this phase far better
    // than 8 did; warps w, w+4, w+8, w+12 see the same TMEM lane quarter and own 32 keys + D/4 output dims each.
    const int q = warp & 3;             // TMEM lane quarter
    const int part = (warp - 2) >> 2;   // which 32 keys of the tile / which quarter of the output dims
    const int r = q * 32 + lane;
    const int qrow = m0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    constexpr int DQ = D / 4;
    float m_ref = -INFINITY;  // reference maximum of this row (raw score units), shared by its four threads
    const int poly = p.poly;
    const float raise_thresh = 8.f / p.scale_log2;  // raise the reference only for a > 2^8 jump
    // every P V product issued so far has retired (commits are cumulative)
    auto wait_pv = [&](int jj) {
      tc::mbar_wait(&pv_full[jj & 1], (jj >> 1) & 1);
      tc::fence_after_sync();
    };

    const uint32_t tS0 = tmem_S0 + lane_off + part * 32;

    for (int j = 0; j < ntiles; ++j) {
      const int key0 = j * FT_BN + part * 32;
      const bool need_mask = (key0 + 32 > Lk) || (p.causal && (key0 + 31 > m0 + q * 32 + shift));
      const int key_lim = p.causal ? min(Lk - 1, qrow + shift) : (Lk - 1);  // last visible key for this row
      tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc::fence_after_sync();
      uint32_t cur[32];  // my 32 scores (fp32 bits), read once
      tc::tmem_ld_32x32(tS0 + (j & 1) * 128, cur);
      tc::tmem_ld_wait();
      // S_j is in registers: the MMA warp may overwrite this buffer with S_{j+2}
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_empty[j & 1]);
      float mx = -INFINITY;
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (key0 + i > key_lim) cur[i] = 0xff800000u;  // -inf
          mx = fmaxf(mx, __uint_as_float(cur[i]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(cur[i]));
      }
      // exchange the quarter-row maxima with the three partner threads (same row, other keys)
      xchg[part * 128 + r] = mx;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      mx = fmaxf(fmaxf(xchg[r], xchg[128 + r]), fmaxf(xchg[256 + r], xchg[384 + r]));
      if (j == 0) {
        m_ref = mx;  // O_0 = P_0 V_0 overwrites the accumulator: nothing to rescale
      } else {
        const bool raise = mx > m_ref + raise_thresh;  // also true when m_ref is still -inf and mx is finite
        if (__any_sync(0xffffffffu, raise)) {
          // rare after the first tiles: bring this warp's 32 rows x D/4 output dims (and 4 of the 16 row-sum
          // columns) to the new reference.  All four partner warps take the same decision (same row maxima).
          const float corr = raise ? exp2f((m_ref - mx) * p.scale_log2) : 1.f;
          if (raise) m_ref = mx;
          wait_pv(j - 1);
#pragma unroll
          for (int c = 0; c < DQ; c += 16) {
            uint32_t raw[16];
            tc::tmem_ld_32x16(tmem_O + lane_off + part * DQ + c, raw);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
            tc::tmem_st_32x16(tmem_O + lane_off + part * DQ + c, raw);
          }
          {
            uint32_t raw[4];
            tc::tmem_ld_32x4(tmem_L + lane_off + part * 4, raw);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
            tc::tmem_st_32x4(tmem_L + lane_off + part * 4, raw);
          }
          tc::tmem_st_wait();
          tc::fence_before_sync();  // ordered before the p_full arrival that releases P V_j
        }
      }
      const float msc = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;
      // the P buffer about to be overwritten was last read by P V_{j-PB}
      if (j >= PB) wait_pv(j - PB);
      // probabilities of my 32 keys -> four 16-byte pieces of P atom `part / 2` (fp16, 128B-swizzled rows of
      // 128 bytes).  P's row sums come from the tensor core (tmem_L), so no per-element adds are spent here.
      uint8_t* prow = sP + (j % PB) * S::P_BYTES + (part >> 1) * S::ATOM + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xa = fmaf(__uint_as_float(cur[g * 8 + 2 * i]), p.scale_log2, -msc);
          const float xb = fmaf(__uint_as_float(cur[g * 8 + 2 * i + 1]), p.scale_log2, -msc);
          // fp32 MUFU.EX2 (the packed-half form issues as two MUFU.EX2.F16 and is no faster); `poly` of every
          // 4 pairs may take an FMA-pipe polynomial instead (p.poly, fixed at 0 by the host launcher = all on the special-function unit)
          float ea, eb;
          if (i < poly) {
            ea = exp2_fma(xa);
            eb = exp2_fma(xb);
          } else {
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(xa));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(xb));
          }
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(eb), "f"(ea));  // {hi: eb, lo: ea}
        }
        const int piece = (part & 1) * 4 + g;
        *reinterpret_cast<uint4*>(prow + ((piece ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      tc::fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // partners have read my max before the next tile's
      if (lane == 0) tc::mbar_arrive(p_full);
    }
    // all products retired: read the accumulator once, normalise, store
    wait_pv(ntiles - 1);
    const float l_tot = __uint_as_float(tc::tmem_ld_32x1(tmem_L + lane_off));
    __half* orow = p.o + b * p.o_sb + (long long)qrow * p.o_sl + h * p.o_sh + part * DQ;
#pragma unroll
    for (int c = 0; c < DQ; c += 16) {
      uint32_t raw[16];
      tc::tmem_ld_32x16(tmem_O + lane_off + part * DQ + c, raw);
      tc::tmem_ld_wait();
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      if (qrow < Lq) {
#pragma unroll
        for (int c8 = 0; c8 < 16; c8 += 8) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(raw[c8 + i]) * inv;
          *reinterpret_cast<vec8*>(orow + c + c8) = pack8<__half>(f);
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}


// =============================================================================================
// head_dim 64, "ping-pong" variant: the 16 softmax warps form TWO independent groups.  Group g owns the key tiles
// j = g (mod 2) together with S buffer g, P buffer g and its own O_g / L_g accumulators in TMEM, and keeps its own
// reference maximum; the two partial results are merged once at the end (as a split-KV decode would).  While one
// group reads its S tile out of TMEM (64 B/clk/SM) the other is on the special-function unit, and neither waits
// for the other's P V product — in the single-group kernel those phases ran in lockstep across all 16 warps.
// A group is 8 warps = 2 threads per query row (64 keys + 32 output dims each).
// TMEM columns: S_0 | S_1 (128 each) | O_0 (256) | O_1 (320) | L_0 (384) | L_1 (400).
// =============================================================================================
__global__ void __launch_bounds__(FT_THREADS, 1) fmha_tc_pp_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const FtParams p) {
  constexpr int D = 64;
  using S = FtSmem<D>;
  extern __shared__ uint8_t ft_smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + S::Q_BYTES;  // stage s: K at s*2*KV, V right after
  uint8_t* sP = sKV + S::STAGES * 2 * S::KV_BYTES;
  uint8_t* sOnes = sP + S::P_BUFS * S::P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + S::ONES_BYTES);
  uint64_t* q_full = bars;        // 1
  uint64_t* kv_full = bars + 1;   // [4]
  uint64_t* kv_empty = bars + 5;  // [4]
  uint64_t* s_full = bars + 9;    // [2] one per group
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* pv_full = bars + 15;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 17);
  float* xchg = reinterpret_cast<float*>(bars + 20);  // [2 groups][2 halves][128 rows] row-max exchange
  float* mfin = xchg + 512;                           // [2 groups][128 rows] final reference maxima

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * FT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int Lq = p.Lq, Lk = p.Lk;
  const int shift = Lk - Lq;
  int n_end = Lk;
  if (p.causal) n_end = min(Lk, m0 + FT_BM + shift);
  const int ntiles = (n_end + FT_BN - 1) / FT_BN;
  const int nst = min(S::STAGES, p.stages);

  for (int i = threadIdx.x; i < S::ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;
  tc::fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      tc::mbar_init(q_full, 1);
      for (int i = 0; i < 4; ++i) {
        tc::mbar_init(&kv_full[i], 1);
        tc::mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&s_full[i], 1);
        tc::mbar_init(&s_empty[i], 8);
        tc::mbar_init(&p_full[i], 8);
        tc::mbar_init(&pv_full[i], 1);
      }
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tmem_ptr_smem, 512);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();  // Q/K/V come from the preceding kernel; O may still be read by an earlier one
  if (*tmem_ptr_smem != 0u) __trap();  // all 512 columns are ours: base 0 keeps every TMEM address warp-uniform
  constexpr uint32_t tmem_S0 = 0u, tmem_O = 256u, tmem_L = 384u;

  if (ntiles == 0) {
    if (warp >= 2 && warp < 6) {
      const int r = (warp & 3) * 32 + lane;
      if (m0 + r < Lq) {
        __half* orow = p.o + b * p.o_sb + (long long)(m0 + r) * p.o_sl + h * p.o_sh;
        for (int d = 0; d < D; ++d) orow[d] = __float2half_rn(0.f);
      }
    }
  } else if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int qrow0 = b * p.q_rows_per_batch + m0;
      tc::mbar_expect_tx(q_full, S::Q_BYTES);
      tc::tma_load_2d(sQ, &tmQ, q_full, p.q_col0 + h * p.q_col_per_head, qrow0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % nst;
        tc::mbar_wait(&kv_empty[s], ((j / nst) & 1) ^ 1);
        uint8_t* sK = sKV + s * 2 * S::KV_BYTES;
        uint8_t* sV = sK + S::KV_BYTES;
        const int krow0 = b * p.k_rows_per_batch + j * FT_BN;
        tc::mbar_expect_tx(&kv_full[s], 2 * S::KV_BYTES);
        tc::tma_load_2d(sK, &tmK, &kv_full[s], p.k_col0 + h * p.k_col_per_head, krow0);
        tc::tma_load_2d(sV, &tmV, &kv_full[s], p.v_col0 + h * p.v_col_per_head, krow0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp runs converged, one elected lane issues =================
    constexpr uint32_t idesc_qk = tc::make_idesc(0, FT_BM, FT_BN);           // A, B K-major
    constexpr uint32_t idesc_pv = tc::make_idesc(0, FT_BM, D) | (1u << 16);  // B (V) MN-major
    constexpr uint32_t idesc_l = tc::make_idesc(0, FT_BM, 16);
    const uint32_t aQ = tc::smem_u32(sQ), aOnes = tc::smem_u32(sOnes);
    auto issue_qk = [&](int j) {
      const int s = j % nst;
      tc::mbar_wait(&kv_full[s], (j / nst) & 1);
      tc::mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t aK = tc::smem_u32(sKV + s * 2 * S::KV_BYTES);
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
        tc::mma_f16_ss_warp(tmem_S0 + (j & 1) * 128, tc::make_desc_sw128(aQ + kk * 32), tc::make_desc_sw128(aK + kk * 32),
                            idesc_qk, kk != 0);
      tc::mma_commit_warp(&s_full[j & 1]);
    };
    tc::mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_qk(j + 1);
      const int g = j & 1;
      tc::mbar_wait(&p_full[g], (j >> 1) & 1);  // P_j (group g) is in shared memory
      tc::fence_after_sync();
      const int s = j % nst;
      const uint32_t aV = tc::smem_u32(sKV + s * 2 * S::KV_BYTES + S::KV_BYTES);
      const uint32_t acc0 = j >= 2;  // the group's first tile overwrites its accumulators
      // A = P_j straight from TMEM (the first 64 columns of S_g, 16 keys = 8 columns per MMA): P never touches
      // shared memory, which was the busiest resource of the kernel (P written once and read twice per tile)
#pragma unroll
      for (int kk = 0; kk < FT_BN / 16; ++kk) {
        const uint32_t ta = tmem_S0 + g * 128 + kk * 8;
        const uint64_t db = make_desc_mn_sw128(aV + kk * 16 * 128, S::ATOM);
        tc::mma_f16_ts_warp(tmem_O + g * 64, ta, db, idesc_pv, acc0 | (kk != 0));
        tc::mma_f16_ts_warp(tmem_L + g * 16, ta, tc::make_desc_sw128(aOnes + (kk >> 2) * 2048 + (kk & 3) * 32), idesc_l,
                            acc0 | (kk != 0));
      }
      tc::mma_commit_warp(&kv_empty[s]);  // K_j / V_j no longer needed
      tc::mma_commit_warp(&pv_full[g]);
    }
  } else {
    // ================= softmax: two groups of 8 warps, a pair of threads per query row in each =================
    const int q = warp & 3;              // TMEM lane quarter
    const int idx = (warp - 2) >> 2;     // 0..3
    const int g = idx & 1;               // group = parity of the key tiles it owns
    const int half = idx >> 1;           // which 64 keys of the tile / which half of the output dims
    const int r = q * 32 + lane;
    const int qrow = m0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int bar_pair = 1 + g * 4 + q;  // named barrier of the 64 threads sharing (group, lane quarter)
    float m_ref = -INFINITY;
    const float raise_thresh = 8.f / p.scale_log2;
    float* xg = xchg + g * 256;
    auto wait_pv = [&](int t) {  // the group's t-th P V product (and everything before it) has retired
      tc::mbar_wait(&pv_full[g], t & 1);
      tc::fence_after_sync();
    };

    int t = 0;
    for (int j = g; j < ntiles; j += 2, ++t) {
      const uint32_t tS = tmem_S0 + g * 128 + lane_off + half * 64;
      const int key0 = j * FT_BN + half * 64;
      const bool need_mask = (key0 + 64 > Lk) || (p.causal && (key0 + 63 > m0 + q * 32 + shift));
      const int key_lim = p.causal ? min(Lk - 1, qrow + shift) : (Lk - 1);
      tc::mbar_wait(&s_full[g], t & 1);
      tc::fence_after_sync();
      uint32_t cur[64];  // my 64 scores (fp32 bits), read once
      tc::tmem_ld_32x32(tS, cur);
      tc::tmem_ld_32x32(tS + 32, cur + 32);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_empty[g]);  // the MMA warp may overwrite S_g with the group's next tile
      float mx = -INFINITY;
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          if (key0 + i > key_lim) cur[i] = 0xff800000u;  // -inf
          mx = fmaxf(mx, __uint_as_float(cur[i]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(cur[i]));
      }
      xg[half * 128 + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_pair) : "memory");
      mx = fmaxf(mx, xg[(half ^ 1) * 128 + r]);
      if (t == 0) {
        m_ref = mx;
      } else {
        const bool raise = mx > m_ref + raise_thresh;
        if (__any_sync(0xffffffffu, raise)) {
          const float corr = raise ? exp2f((m_ref - mx) * p.scale_log2) : 1.f;
          if (raise) m_ref = mx;
          wait_pv(t - 1);
#pragma unroll
          for (int c = 0; c < 32; c += 16) {
            uint32_t raw[16];
            tc::tmem_ld_32x16(tmem_O + g * 64 + lane_off + half * 32 + c, raw);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
            tc::tmem_st_32x16(tmem_O + g * 64 + lane_off + half * 32 + c, raw);
          }
#pragma unroll
          for (int c = 0; c < 8; c += 4) {
            uint32_t raw[4];
            tc::tmem_ld_32x4(tmem_L + g * 16 + lane_off + half * 8 + c, raw);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
            tc::tmem_st_32x4(tmem_L + g * 16 + lane_off + half * 8 + c, raw);
          }
          tc::tmem_st_wait();
          tc::fence_before_sync();
        }
      }
      const float msc = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;
      // P_j (fp16 pairs, key 2c in the low half of column c) overwrites the first 64 columns of S_g: my 64 keys are
      // columns half*32 .. +31.  Both threads of the row hold their scores in registers (the exchange barrier above),
      // and the tensor pipe is in order, so the group's next Q K^T cannot overtake the P V that reads this.
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float xa = fmaf(__uint_as_float(cur[2 * i]), p.scale_log2, -msc);
        const float xb = fmaf(__uint_as_float(cur[2 * i + 1]), p.scale_log2, -msc);
        float ea, eb;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(xa));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(xb));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(eb), "f"(ea));  // {hi: eb, lo: ea}
      }
      tc::tmem_st_32x32(tmem_S0 + g * 128 + lane_off + half * 32, pk);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      asm volatile("bar.sync %0, 64;" ::"r"(bar_pair) : "memory");  // partner has read my max before the next tile's
      if (lane == 0) tc::mbar_arrive(&p_full[g]);
    }
    // ---- merge the two groups' partial results ----
    // the last product overall has retired => so has every earlier one of either group (commits are cumulative)
    tc::mbar_wait(&pv_full[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc::fence_after_sync();
    if (half == 0) mfin[g * 128 + r] = m_ref;
    asm volatile("bar.sync %0, 128;" ::"r"(9 + q) : "memory");
    const bool has_b = ntiles >= 2;  // group 1 owns no tile when there is a single key tile (its TMEM is undefined)
    const float mA = mfin[r], mB = has_b ? mfin[128 + r] : -INFINITY;
    const float mm = fmaxf(mA, mB);
    const float wA = (mA == -INFINITY) ? 0.f : exp2f((mA - mm) * p.scale_log2);
    const float wB = (mB == -INFINITY) ? 0.f : exp2f((mB - mm) * p.scale_log2);
    const int d0 = half * 32 + g * 16;  // this thread's 16 output dims
    uint32_t oa[16], ob[16];
    tc::tmem_ld_32x16(tmem_O + lane_off + d0, oa);
    float l_tot = __uint_as_float(tc::tmem_ld_32x1(tmem_L + lane_off)) * wA;
    if (has_b) {
      tc::tmem_ld_32x16(tmem_O + 64 + lane_off + d0, ob);
      const float lb = __uint_as_float(tc::tmem_ld_32x1(tmem_L + 16 + lane_off));
      tc::tmem_ld_wait();
      l_tot += lb * wB;
    } else {
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) ob[i] = 0u;
    }
    if (qrow < Lq) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      __half* orow = p.o + b * p.o_sb + (long long)qrow * p.o_sl + h * p.o_sh + d0;
#pragma unroll
      for (int c8 = 0; c8 < 16; c8 += 8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          f[i] = (__uint_as_float(oa[c8 + i]) * wA + __uint_as_float(ob[c8 + i]) * wB) * inv;
        *reinterpret_cast<vec8*>(orow + c8) = pack8<__half>(f);
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(0u, 512);
}

// tensor-map helper (defined in gemm_tc.cu)
}  // namespace

int ss_internal_get_tmap(CUtensorMap* out, const void* ptr, int dtype, int rank, const uint64_t* dims,
                         const uint64_t* strides, const uint32_t* box);

namespace {
template <int D>
int launch_ft(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FtParams& p, int B,
              cudaStream_t s) {
  using S = FtSmem<D>;
  dim3 grid((p.Lq + FT_BM - 1) / FT_BM, p.H, B);
  if constexpr (D == 64) {
    // head_dim 64 (UNet self-attention, resamplers): the ping-pong kernel (two softmax groups, P in TMEM)
    static bool attr_pp = false;
    if (!attr_pp) {
      SS_CUDA(cudaFuncSetAttribute(fmha_tc_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
      attr_pp = true;
    }
    SS_CUDA(ss::launch_pdl(fmha_tc_pp_kernel, grid, dim3(FT_THREADS), (size_t)S::TOTAL, s, tq, tk, tv, p));
  } else {
    static bool attr = false;
    if (!attr) {
      SS_CUDA(cudaFuncSetAttribute(fmha_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
      attr = true;
    }
    SS_CUDA(ss::launch_pdl(fmha_tc_kernel<D>, grid, dim3(FT_THREADS), (size_t)S::TOTAL, s, tq, tk, tv, p));
  }
  SS_LAUNCH_CHECK();
  return 0;
}
}  // namespace

// Returns 0 on success, -1 when the operand layout is not expressible as row-matrix views (caller falls back to
// the mma.sync kernel), > 0 on error.
int ss_internal_fmha_tc(const void* q, const void* k, const void* v, void* out, int B, int H, int Lq, int Lk, int D,
                        long long q_sb, long long q_sl, long long q_sh, long long k_sb, long long k_sl, long long k_sh,
                        long long v_sb, long long v_sl, long long v_sh, long long o_sb, long long o_sl, long long o_sh,
                        float scale, int causal, cudaStream_t stream) {
  if (D != 64 && D != 128) return -1;
  // heads must sit side by side inside a token row, batches must be stacked row blocks (or broadcast)
  auto ok = [&](long long sb, long long sl, long long sh, int L) {
    return sl % 8 == 0 && sh % 8 == 0 && sh * (H - 1) + D <= sl && (sb == 0 || sb == (long long)L * sl);
  };
  if (!ok(q_sb, q_sl, q_sh, Lq) || !ok(k_sb, k_sl, k_sh, Lk) || !ok(v_sb, v_sl, v_sh, Lk)) return -1;
  if (k_sb != v_sb && !(k_sb == 0 || v_sb == 0)) return -1;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) return -1;
  if (o_sl % 8 != 0 || o_sh % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return -1;
  CUtensorMap tq, tk, tv;
  auto mk = [&](CUtensorMap* tm, const void* ptr, long long sb, long long sl, int L) {
    const uint64_t rows = (uint64_t)(sb == 0 ? L : (long long)B * L);
    uint64_t dims[2] = {(uint64_t)sl, rows}, str[1] = {(uint64_t)sl * 2};
    uint32_t box[2] = {64, 128};
    return ss_internal_get_tmap(tm, ptr, SS_F16, 2, dims, str, box);
  };
  if (int e = mk(&tq, q, q_sb, q_sl, Lq)) return e;
  if (int e = mk(&tk, k, k_sb, k_sl, Lk)) return e;
  if (int e = mk(&tv, v, v_sb, v_sl, Lk)) return e;
  FtParams p;
  memset(&p, 0, sizeof(p));
  p.o = (__half*)out;
  p.o_sb = o_sb; p.o_sl = o_sl; p.o_sh = o_sh;
  p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.q_rows_per_batch = q_sb == 0 ? 0 : Lq;
  p.k_rows_per_batch = k_sb == 0 ? 0 : Lk;
  p.q_col_per_head = (int)q_sh; p.k_col_per_head = (int)k_sh; p.v_col_per_head = (int)v_sh;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.stages = 4;  // K/V ring depth (head_dim 128 uses its own 2-stage layout)
  p.poly = 0;    // exponentials on the MUFU unit (the polynomial split measured no faster, profiles/r1_fmha_experiments.md)
  if (D == 64) return launch_ft<64>(tq, tk, tv, p, B, stream);
  return launch_ft<128>(tq, tk, tv, p, B, stream);
}

// Paged variant (head_dim 128): K / V live in a page pool [page][H][64][128]; sequence b's keys are the pages of row b of
// page_table.  Returns 0 on success, -1 when the layout does not fit (caller falls back to the mma.sync kernel).
int ss_internal_fmha_tc_paged(const void* q, const void* kpool, const void* vpool, void* out, int B, int H, int Lq, int Lk,
                              int D, long long q_sb, long long q_sl, long long q_sh, long long o_sb, long long o_sl,
                              long long o_sh, const int* page_table, int max_pages, float scale, int causal,
                              cudaStream_t stream) {
  if (D != 128) return -1;
  if (q_sl % 8 != 0 || q_sh % 8 != 0 || q_sh * (H - 1) + D > q_sl || !(q_sb == 0 || q_sb == (long long)Lq * q_sl)) return -1;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(kpool) | reinterpret_cast<uintptr_t>(vpool)) & 15) return -1;
  if (o_sl % 8 != 0 || o_sh % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return -1;
  CUtensorMap tq, tk, tv;
  {
    const uint64_t rows = (uint64_t)(q_sb == 0 ? Lq : (long long)B * Lq);
    uint64_t dims[2] = {(uint64_t)q_sl, rows}, str[1] = {(uint64_t)q_sl * 2};
    uint32_t box[2] = {64, 128};
    if (int e = ss_internal_get_tmap(&tq, q, SS_F16, 2, dims, str, box)) return e;
  }
  auto mkp = [&](CUtensorMap* tm, const void* pool) {
    // (dim within head, token within page, page * H + head); the pool's page count is not known here: the extent is
    // an upper bound, every access is to a page the table names
    uint64_t dims[3] = {(uint64_t)D, 64, (uint64_t)1 << 24}, str[2] = {(uint64_t)D * 2, (uint64_t)64 * D * 2};
    uint32_t box[3] = {64, 64, 1};
    return ss_internal_get_tmap(tm, pool, SS_F16, 3, dims, str, box);
  };
  if (int e = mkp(&tk, kpool)) return e;
  if (int e = mkp(&tv, vpool)) return e;
  FtParams p;
  memset(&p, 0, sizeof(p));
  p.o = (__half*)out;
  p.o_sb = o_sb; p.o_sl = o_sl; p.o_sh = o_sh;
  p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.q_rows_per_batch = q_sb == 0 ? 0 : Lq;
  p.q_col_per_head = (int)q_sh;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.stages = 4;
  p.poly = 0;
  p.page_table = page_table;
  p.paged = 1;
  p.max_pages = max_pages;
  return launch_ft<128>(tq, tk, tv, p, B, stream);
}
