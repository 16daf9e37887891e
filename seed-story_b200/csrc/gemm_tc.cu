// Dense contraction on the 5th-gen tensor cores:  C[M,N] = epilogue(A[M,K] · B[N,K]^T)
//   * operands: 16-bit (fp16 or bf16), both K-major (x[M,K] and nn.Linear weight [N,K] as stored)
//   * TMA (cp.async.bulk.tensor, 128B swizzle) stages 128 x 64 / BN x 64 tiles into a shared-memory ring
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator
//   * four epilogue warps read TMEM with tcgen05.ld and apply bias / activation / GLU / residual
//   * the same kernel is an implicit-GEMM 3x3 convolution (NHWC): the A tile for filter tap (ky,kx)
//     is a 4-D TMA box shifted by (ky-1, kx-1); out-of-bounds pixels are zero-filled by the TMA unit,
//     which IS the conv padding.
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "tc.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 bytes = one swizzle row
constexpr int GEMM_PERSIST_THREADS = 320;  // TMA, MMA + two epilogue groups of 4 warps

struct GemmParams {
  void* out;
  int ldo;
  const void* bias;      // [N] or null
  const void* bias2;     // [groups, ld_b2] or null; group = row / rows_per_group
  int rows_per_group;
  int ld_b2;
  const void* residual;  // [M, ldr] or null
  int ldr;
  int act;  // 0 none, 1 gelu(erf), 2 silu
  int glu;  // 0 none, 1 first*gelu(second), 2 silu(first)*second   (column pairs interleaved)
  float alpha;
  int M, N, K;
  int b_const;  // B is a weight matrix nothing on the stream writes: its first tiles may be fetched before the PDL wait
  // LayerNorm folded around the GEMM.  Consumer side: A holds the RAW rows x, B holds the ROW-CENTRED weights
  // W'' = gamma (.) W - rowmean(gamma (.) W), for which x W''^T = (x - mean(x)) (gamma (.) W)^T exactly (the mean
  // subtraction rides on the contraction), so the epilogue only multiplies by rstd and adds `bias` = beta W^T + b;
  // rstd of a row comes from `ln_slots` partial (sum, sum of squares) pairs laid out [slot][M] that the GEMM which
  // produced x left behind.  Producer side: stats_out receives those partials for this GEMM's (rounded) output rows,
  // two slots per N tile.
  const float* ln_stats;
  int ln_slots;
  float ln_eps;
  float* stats_out;
  // conv mode
  int H, W, Cin, bw, bh, tiles_x, tiles_y;
};

template <typename T>
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return gelu_erf(v);
  if (act == 2) return v / (1.f + expf(-v));
  return v;
}

template <typename T, int BN, bool REMOTE>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const CUtensorMap* tmC, uint8_t* stg, uint32_t taddr,
                                              uint64_t* full_bar, uint32_t full_parity, uint64_t* empty_bar, int n0,
                                              int m0, int q, int lane, int f_begin, int f_step) {
  const int r = q * 32 + lane;
  const int acc_per_fill = p.glu ? 128 : 64;
  constexpr int kTail = BN % 64;  // 32 for BN = 160, else 0 (a GLU tail is BN % 128 = 32 as well)
  const int nfills = (BN + acc_per_fill - 1) / acc_per_fill;
  const long long m = (long long)m0 + r;
  const bool row_ok = m < (long long)p.M;
  const T* res_row = (p.residual && row_ok) ? reinterpret_cast<const T*>(p.residual) + m * p.ldr : nullptr;
  const T* b2_row = (p.bias2 && row_ok)
                        ? reinterpret_cast<const T*>(p.bias2) + (m / p.rows_per_group) * (long long)p.ld_b2
                        : nullptr;
  const T* bias = reinterpret_cast<const T*>(p.bias);
  T* out_row = reinterpret_cast<T*>(p.out) + m * p.ldo;
  const int n_out_cols = p.glu ? (p.N >> 1) : p.N;

  // the residual of a 32-column chunk is requested one chunk ahead (the first one before the accumulator is
  // even complete), so its global-memory latency hides behind the main loop and the previous chunk
  vec8 nr[4];
  auto request = [&](int col0) {
#pragma unroll
    for (int gI = 0; gI < 4; ++gI) {
      const int col = col0 + gI * 8;
      nr[gI] = (res_row && col < p.N) ? ld_cached16(res_row + col) : vec8{0u, 0u, 0u, 0u};
    }
  };
  if (f_begin < nfills) request(n0 + f_begin * acc_per_fill);
  // folded LayerNorm, consumer side: this row's mean / rstd from the partial sums its producer left (summed in slot
  // order: deterministic); fetched while the main loop still runs
  float out_scale = p.alpha;  // accumulator -> value scale: alpha, times this row's rstd under a folded LayerNorm
  if (p.ln_stats) {
    float s1 = 0.f, s2 = 0.f;
    if (row_ok) {
      const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + m;
      // eight slot loads in flight at a time (one memory latency per batch instead of one per slot)
      for (int s0 = 0; s0 < p.ln_slots; s0 += 8) {
        float2 v2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v2[u] = (s0 + u < p.ln_slots) ? __ldcg(st + (size_t)(s0 + u) * p.M) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          s1 += v2[u].x;
          s2 += v2[u].y;
        }
      }
    }
    const float inv_k = 1.f / (float)p.K;
    const float mean = s1 * inv_k;
    const float var = fmaxf(s2 * inv_k - mean * mean, 0.f);
    out_scale *= rsqrtf(var + p.ln_eps);
  }
  float st_sum = 0.f, st_sq = 0.f;  // producer side: statistics of this thread's row over the columns it stores
  tc::mbar_wait(full_bar, full_parity);
  tc::fence_after_sync();

#pragma unroll 1
  for (int f = f_begin; f < nfills; f += f_step) {
    const int c0 = f * acc_per_fill;
    const int fill_cols = min(acc_per_fill, BN - c0);
    const bool direct = kTail != 0 && fill_cols < acc_per_fill;  // tail: registers -> global, no staging
    const bool last_fill = f + f_step >= nfills;
    if (!direct) {
      if (lane == 0) tc::tma_store_wait_read<0>();  // the previous store from this buffer has been read out
      __syncwarp();
    }
#pragma unroll 1
    for (int cc = 0; cc < fill_cols; cc += 32) {
      const int c = c0 + cc;
      const int col0 = n0 + c;
      vec8 vb[4], vr[4], vb2[4];
#pragma unroll
      for (int gI = 0; gI < 4; ++gI) {
        vr[gI] = nr[gI];
        const int col = col0 + gI * 8;
        const bool col_ok = col < p.N;
        vb[gI] = (bias && col_ok) ? ld_cached16(bias + col) : vec8{0u, 0u, 0u, 0u};
        vb2[gI] = (b2_row && col_ok) ? ld_cached16(b2_row + col) : vec8{0u, 0u, 0u, 0u};
      }
      {  // next chunk of this warp: same fill, or the first chunk of its next fill
        int nf = f, ncc = cc + 32;
        if (ncc >= fill_cols) nf = f + f_step, ncc = 0;
        if (nf < nfills) request(n0 + nf * acc_per_fill + ncc);
      }
      uint32_t raw[32];
      tc::tmem_ld_32x32(taddr + (uint32_t)c, raw);
      tc::tmem_ld_wait();
      if (last_fill && cc + 32 >= fill_cols) {  // last TMEM read of this tile: hand the accumulator back
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (REMOTE)
            tc::mbar_arrive_remote(empty_bar, 0);  // the pair's accumulator-drained barrier lives in the leader CTA
          else
            tc::mbar_arrive(empty_bar);
        }
      }
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * out_scale;
#pragma unroll
      for (int gI = 0; gI < 4; ++gI) {
        float* vv = v + gI * 8;
        float bf[8];
        unpack8<T>(vb[gI], bf);
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] = ss_num<T>::to_f(ss_num<T>::from_f(vv[i] + bf[i]));
        if (p.bias2) {
          unpack8<T>(vb2[gI], bf);
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] = ss_num<T>::to_f(ss_num<T>::from_f(vv[i] + bf[i]));
        }
        if (p.glu == 0) {
          if (p.act) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = ss_num<T>::to_f(ss_num<T>::from_f(act_apply<T>(vv[i], p.act)));
          }
          if (p.residual) {
            unpack8<T>(vr[gI], bf);
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] += bf[i];
          }
          if (p.stats_out && col0 + gI * 8 < p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float r = ss_num<T>::to_f(ss_num<T>::from_f(vv[i]));  // the value that is stored
              st_sum += r;
              st_sq = fmaf(r, r, st_sq);
            }
          }
          if (direct) {
            if (row_ok && col0 + gI * 8 < p.N) st16(out_row + col0 + gI * 8, pack8<T>(vv));
          } else {
            // 16-byte piece j of this row inside the 128-byte staging row, 128B-swizzled like the TMA expects
            const int j = (cc >> 3) + gI;
            *reinterpret_cast<vec8*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = pack8<T>(vv);
          }
        } else {
          T o4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = vv[2 * i], b = vv[2 * i + 1];
            float o;
            if (p.glu == 1)
              o = a * ss_num<T>::to_f(ss_num<T>::from_f(gelu_erf(b)));
            else
              o = ss_num<T>::to_f(ss_num<T>::from_f(a / (1.f + __expf(-a)))) * b;
            o4[i] = ss_num<T>::from_f(o);
          }
          // 8 accumulator columns -> 4 outputs = 8 bytes; output column within the fill = (cc + gI*8) / 2
          const int ocol = (cc + gI * 8) >> 1;               // 0..63
          if (direct) {
            const int gcol = ((n0 + c0) >> 1) + ocol;
            if (row_ok && gcol < n_out_cols)
              *reinterpret_cast<uint2*>(out_row + gcol) = *reinterpret_cast<const uint2*>(o4);
          } else {
            const int j = ocol >> 3, within = (ocol & 7) * 2;  // 16-byte piece, byte offset inside it
            *reinterpret_cast<uint2*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4) + within) =
                *reinterpret_cast<const uint2*>(o4);
          }
        }
      }
    }
    if (direct) continue;
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const int out_col = p.glu ? ((n0 + c0) >> 1) : (n0 + c0);
      if (out_col < n_out_cols && (long long)m0 + q * 32 < (long long)p.M)
        tc::tma_store_2d(tmC, stg, out_col, m0 + q * 32);
      tc::tma_store_commit();
    }
  }
  if (p.stats_out && row_ok) {
    // two slots per N tile: the two epilogue groups' halves of the tile, or (whole tile by one group) the sums and zeros
    float2* so = reinterpret_cast<float2*>(p.stats_out) + m;
    const int slot = (n0 / BN) * 2;
    if (f_step == 2) {
      so[(size_t)(slot + f_begin) * p.M] = make_float2(st_sum, st_sq);
    } else {
      so[(size_t)slot * p.M] = make_float2(st_sum, st_sq);
      so[(size_t)(slot + 1) * p.M] = make_float2(0.f, 0.f);
    }
  }
}

template <int BN>
struct PersistLayout {
  // a stage holds KATOMS k-atoms of 64 elements: wide stages amortise the per-stage barrier round trip of the
  // single MMA-issuing thread when the N tile (and with it the tensor-core time per atom) is small
  static constexpr int KATOMS = (BN >= 160) ? 1 : 2;
  static constexpr int A_BYTES = BM * BK * 2;   // per atom
  static constexpr int B_BYTES = BN * BK * 2;   // per atom
  static constexpr int STAGE_BYTES = KATOMS * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 160 ? 5 : (BN == 128 ? 3 : 4));
  // TMEM: two accumulators; the 160-wide tile (every SDXL channel count is a multiple of 160) keeps them on
  // 256-column boundaries, and allocations must be powers of two
  static constexpr int TMEM_STRIDE = (BN == 160) ? 256 : BN;
  static constexpr int TMEM_COLS = 2 * TMEM_STRIDE;
  static constexpr int STAGING_BYTES = 8 /*epilogue warps*/ * 32 * 128;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <typename T, int BN, bool CONV>
__global__ void __launch_bounds__(GEMM_PERSIST_THREADS, 1) gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const __grid_constant__ CUtensorMap tmC,
                                                                         const GemmParams p, int n_tiles_n,
                                                                         int total_tiles) {
  using L = PersistLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::STAGES * L::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* tmem_full_bar = empty_bar + L::STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = CONV ? (p.Cin / BK) : ((p.K + BK - 1) / BK);
  const int num_atoms = CONV ? 9 * kchunks : kchunks;                  // 64-wide k atoms per tile
  const int num_kb = (num_atoms + L::KATOMS - 1) / L::KATOMS;          // pipeline stages per tile

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::prefetch_tmap(&tmC);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < L::STAGES; ++s) {
        tc::mbar_init(&full_bar[s], 1);
        tc::mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        tc::mbar_init(&tmem_full_bar[b], 1);
        tc::mbar_init(&tmem_empty_bar[b], 4);  // one arrival per epilogue warp
      }
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tmem_ptr_smem, L::TMEM_COLS);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile -> coordinates
  auto tile_coords = [&](int tile, int& n0, int& m0, int& img, int& y0, int& x0) {
    const int tn = tile % n_tiles_n;
    int tm = tile / n_tiles_n;
    n0 = tn * BN;
    if (CONV) {
      const int tx = tm % p.tiles_x;
      tm /= p.tiles_x;
      const int ty = tm % p.tiles_y;
      img = tm / p.tiles_y;
      y0 = ty * p.bh;
      x0 = tx * p.bw;
      m0 = (img * p.H + y0) * p.W + x0;  // output rows of a tile are contiguous in the flat [N*H*W, Cout] view
    } else {
      m0 = tm * BM;
      img = y0 = x0 = 0;
    }
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      // The B operand of a linear layer / convolution is a weight matrix no kernel writes, so the first pipeline
      // fill of B is requested BEFORE waiting on the preceding kernel: the HBM latency of the weights then overlaps
      // that kernel's tail.  A (and everything the epilogue reads) is produced by earlier kernels in the stream.
      int pre = 0;
      if (p.b_const && (int)blockIdx.x < total_tiles) {
        int n0, m0, img, y0, x0;
        tile_coords(blockIdx.x, n0, m0, img, y0, x0);
        pre = min(L::STAGES, num_kb);
        for (int kb = 0; kb < pre; ++kb) {
          uint8_t* stage = smem + kb * L::STAGE_BYTES;
          const int na = min(L::KATOMS, num_atoms - kb * L::KATOMS);
          tc::mbar_expect_tx(&full_bar[kb], na * (L::A_BYTES + L::B_BYTES));
          for (int a = 0; a < na; ++a) {
            const int atom = kb * L::KATOMS + a;
            uint8_t* sb = stage + a * (L::A_BYTES + L::B_BYTES) + L::A_BYTES;
            if (CONV) {
              const int tap = atom / kchunks, c0 = (atom % kchunks) * BK;
              tc::tma_load_2d(sb, &tmB, &full_bar[kb], tap * p.Cin + c0, n0);
            } else {
              tc::tma_load_2d(sb, &tmB, &full_bar[kb], atom * BK, n0);
            }
          }
        }
      }
      pdl_wait();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n0, m0, img, y0, x0;
        tile_coords(tile, n0, m0, img, y0, x0);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % L::STAGES;
          const uint32_t ph = (it / L::STAGES) & 1;
          const bool b_done = (int)it < pre;  // this stage's barrier is armed and its B tile already in flight
          uint8_t* stage = smem + s * L::STAGE_BYTES;
          const int na = min(L::KATOMS, num_atoms - kb * L::KATOMS);
          if (!b_done) {
            tc::mbar_wait(&empty_bar[s], ph ^ 1);
            tc::mbar_expect_tx(&full_bar[s], na * (L::A_BYTES + L::B_BYTES));
          }
          for (int a = 0; a < na; ++a) {
            const int atom = kb * L::KATOMS + a;
            uint8_t* sa = stage + a * (L::A_BYTES + L::B_BYTES);
            uint8_t* sb = sa + L::A_BYTES;
            if (CONV) {
              const int tap = atom / kchunks, c0 = (atom % kchunks) * BK;
              const int ky = tap / 3, kx = tap % 3;
              tc::tma_load_4d(sa, &tmA, &full_bar[s], c0, x0 + kx - 1, y0 + ky - 1, img);
              if (!b_done) tc::tma_load_2d(sb, &tmB, &full_bar[s], tap * p.Cin + c0, n0);
            } else {
              tc::tma_load_2d(sa, &tmA, &full_bar[s], atom * BK, m0);
              if (!b_done) tc::tma_load_2d(sb, &tmB, &full_bar[s], atom * BK, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(sizeof(T) == 2 && !std::is_same<T, __half>::value, BM, BN);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1, use = lt >> 1;
        tc::mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + buf * L::TMEM_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % L::STAGES;
          const uint32_t ph = (it / L::STAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::fence_after_sync();
          const int na = min(L::KATOMS, num_atoms - kb * L::KATOMS);
          for (int a = 0; a < na; ++a) {
            const uint32_t sa = tc::smem_u32(smem + s * L::STAGE_BYTES + a * (L::A_BYTES + L::B_BYTES));
            const uint64_t da = tc::make_desc_sw128(sa);
            const uint64_t db = tc::make_desc_sw128(sa + L::A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::mma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | a | k) != 0);
          }
          tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ================= epilogue: two groups of 4 warps (warps 2..5 and 6..9) =================
    // group g drains TMEM accumulator g, i.e. every other tile of this CTA, so two epilogues are in flight
    // while the MMA warp works on the next tile.
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;   // 0 or 1
    uint8_t* stg = staging + ((grp * 4 + q) * (32 * 128));
    pdl_wait();  // residual / bias2 come from earlier kernels; our stores must not overtake their readers
    const int nfills = (BN + (p.glu ? 128 : 64) - 1) / (p.glu ? 128 : 64);  // see epilogue_tile()
    // one-tile CTAs (small problems): nothing to overlap with, so the two groups take alternate fills of that tile
    const bool split_cols = (total_tiles <= (int)gridDim.x) && nfills >= 2;
    const int f_begin = split_cols ? grp : 0, f_step = split_cols ? 2 : 1;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      if (!split_cols && (int)(lt & 1) != grp) continue;
      int n0, m0, img, y0, x0;
      tile_coords(tile, n0, m0, img, y0, x0);
      const uint32_t buf = lt & 1, use = lt >> 1;
      epilogue_tile<T, BN, false>(p, &tmC, stg, tmem_base + buf * L::TMEM_STRIDE + ((uint32_t)(q * 32) << 16),
                                  &tmem_full_bar[buf], use & 1, &tmem_empty_bar[buf], n0, m0, q, lane, f_begin, f_step);
    }
    // the staging buffers only have to outlive the READ side of the last tile stores; the writes complete with the grid
    if (lane == 0) tc::tma_store_wait_read<0>();
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, L::TMEM_COLS);
}


// =============================================================================================
// v3: CTA-pair kernel (cta_group::2).  A cluster of two CTAs (one TPC) owns a 256 x 256 output tile: each CTA
// stages ITS 128 rows of A and ITS 128 rows (half) of the B tile, the leader CTA issues tcgen05.mma with M = 256
// that reads both CTAs' shared memory, and each CTA's TMEM receives its own 128 accumulator rows.  Per CTA the
// shared-memory traffic per MMA drops from (128 + 256) to (128 + 128) operand rows, which is what lets the
// tensor pipe run past the single-CTA ceiling; stages shrink to 32 KB, so the ring is 6 deep.
// Barrier protocol: both producers signal the LEADER's full barrier (TMA .cta_group::2 + remote arrive);
// tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs; both epilogues arrive on the
// leader's "accumulator drained" barrier.
// =============================================================================================
template <int PBN>
struct PairLayout {
  static constexpr int BN = PBN;                     // pair tile width: 256, or 160 (every SDXL width is a multiple)
  static constexpr int A_BYTES = BM * BK * 2;        // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // 16 / 10 KB: this CTA's half of the N tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 7;
  static constexpr int STAGING_BYTES = 8 * 32 * 128;
  static constexpr int TMEM_STRIDE = 256, TMEM_COLS = 512;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
};

template <typename T, bool CONV, int PBN>
__global__ void __launch_bounds__(GEMM_PERSIST_THREADS, 1) gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmB,
                                                                      const __grid_constant__ CUtensorMap tmC,
                                                                      const GemmParams p, int n_tiles_n,
                                                                      int total_tiles) {
  using L = PairLayout<PBN>;
  constexpr int BN = L::BN;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::STAGES * L::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* tmem_full_bar = empty_bar + L::STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int kchunks = CONV ? (p.Cin / BK) : ((p.K + BK - 1) / BK);
  const int num_kb = CONV ? 9 * kchunks : kchunks;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 2);   // leader's copy is the one used: both CTAs' producers arrive on it
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full_bar[b], 1);
      tc::mbar_init(&tmem_empty_bar[b], 8);  // 4 epilogue warps of the owning group in EACH CTA (leader's copy)
    }
    tc::fence_barrier_init();
  }
  tc::cluster_sync();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  if (warp == 1) tc::tmem_alloc_2sm(tmem_ptr_smem, L::TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // pair tile -> this CTA's sub-tile coordinates
  auto tile_coords = [&](int tile, int& n0, int& m0, int& img, int& y0, int& x0) {
    const int tn = tile % n_tiles_n;
    int st = (tile / n_tiles_n) * 2 + (int)rank;  // 128-row sub-tile index of this CTA
    n0 = tn * BN;
    if (CONV) {
      const int tx = st % p.tiles_x;
      st /= p.tiles_x;
      const int ty = st % p.tiles_y;
      img = st / p.tiles_y;
      y0 = ty * p.bh;
      x0 = tx * p.bw;
      m0 = (img * p.H + y0) * p.W + x0;
    } else {
      m0 = st * BM;
      img = y0 = x0 = 0;
    }
  };

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      pdl_wait();
      uint32_t it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        int n0, m0, img, y0, x0;
        tile_coords(tile, n0, m0, img, y0, x0);
        const int nb0 = n0 + (int)rank * (BN / 2);  // my half of the B tile
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % L::STAGES;
          const uint32_t ph = (it / L::STAGES) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (CONV) {
            const int tap = kb / kchunks, c0 = (kb % kchunks) * BK;
            const int ky = tap / 3, kx = tap % 3;
            tc::tma_load_4d_2sm(sa, &tmA, &full_bar[s], c0, x0 + kx - 1, y0 + ky - 1, img);
            tc::tma_load_2d_2sm(sb, &tmB, &full_bar[s], tap * p.Cin + c0, nb0);
          } else {
            tc::tma_load_2d_2sm(sa, &tmA, &full_bar[s], kb * BK, m0);
            tc::tma_load_2d_2sm(sb, &tmB, &full_bar[s], kb * BK, nb0);
          }
          if (rank == 0)
            tc::mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);  // bytes of both CTAs land on the leader's barrier
          else
            tc::mbar_arrive_remote(&full_bar[s], 0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc(sizeof(T) == 2 && !std::is_same<T, __half>::value, 2 * BM, BN);
      uint32_t it = 0, lt = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++lt) {
        const uint32_t buf = lt & 1, use = lt >> 1;
        tc::mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + buf * L::TMEM_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % L::STAGES;
          const uint32_t ph = (it / L::STAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t da = tc::make_desc_sw128(sa);
          const uint64_t db = tc::make_desc_sw128(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::mma_f16_ss_2sm(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          tc::mma_commit_2sm(&empty_bar[s]);
        }
        tc::mma_commit_2sm(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ================= epilogue (both CTAs): two groups of 4 warps, each CTA drains its own 128 rows =================
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    uint8_t* stg = staging + ((grp * 4 + q) * (32 * 128));
    pdl_wait();
    const int nfills = (BN + (p.glu ? 128 : 64) - 1) / (p.glu ? 128 : 64);
    const bool split_cols = (total_tiles <= num_clusters) && nfills >= 2;
    const int f_begin = split_cols ? grp : 0, f_step = split_cols ? 2 : 1;
    uint32_t lt = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++lt) {
      if (!split_cols && (int)(lt & 1) != grp) continue;
      int n0, m0, img, y0, x0;
      tile_coords(tile, n0, m0, img, y0, x0);
      const uint32_t buf = lt & 1, use = lt >> 1;
      // "accumulator ready" is multicast to both CTAs' barriers; "accumulator drained" lives in the leader CTA
      epilogue_tile<T, BN, true>(p, &tmC, stg, tmem_base + buf * L::TMEM_STRIDE + ((uint32_t)(q * 32) << 16),
                                 &tmem_full_bar[buf], use & 1, &tmem_empty_bar[buf], n0, m0, q, lane, f_begin, f_step);
    }
    // the staging buffers only have to outlive the READ side of the last tile stores; the writes complete with the grid
    if (lane == 0) tc::tma_store_wait_read<0>();
  }

  tc::fence_before_sync();
  tc::cluster_sync();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, L::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point resolved at run time) + cache
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int dtype, rank;
  uint64_t dims[4];
  uint64_t strides[3];
  uint32_t box[4];
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return h;
  }
};
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
std::mutex g_tmap_mu;

// dims/strides innermost first; strides in bytes for dims 1..rank-1
int get_tmap(CUtensorMap* out, const void* ptr, int dtype, int rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  key.dtype = dtype;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i + 1 < rank) key.strides[i] = strides[i];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  SS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides[i];
  }
  SS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i) SS_REQUIRE(strides[i] % 16 == 0, "TMA strides must be multiples of 16 bytes");
  CUtensorMap tm;
  CUresult r = fn(&tm, dtype == SS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                  const_cast<void*>(ptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SS_FAIL("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() > 65536) g_tmaps.clear();
    g_tmaps[key] = tm;
  }
  *out = tm;
  return 0;
}

}  // namespace

// shared with fmha_tc.cu
int ss_internal_get_tmap(CUtensorMap* out, const void* ptr, int dtype, int rank, const uint64_t* dims,
                         const uint64_t* strides, const uint32_t* box) {
  return get_tmap(out, ptr, dtype, rank, dims, strides, box);
}

namespace {

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T, int BN, bool CONV>
int launch_persist(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const GemmParams& p,
                   int n_tiles_n, long long total_tiles, cudaStream_t s) {
  using L = PersistLayout<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    SS_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel<T, BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 L::TOTAL));
    attr_set = true;
  }
  const int grid = (int)(total_tiles < sm_count() ? total_tiles : sm_count());
  SS_CUDA(ss::launch_pdl(gemm_tc_persist_kernel<T, BN, CONV>, dim3(grid), dim3(GEMM_PERSIST_THREADS), (size_t)L::TOTAL, s,
                         ta, tb, tcm, p, n_tiles_n, (int)total_tiles));
  return 0;
}

template <bool CONV>
int dispatch_persist(int dtype, int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm,
                     const GemmParams& p, int n_tiles_n, long long total_tiles, cudaStream_t s) {
  if (dtype == SS_F16) {
    if (bn == 64) return launch_persist<__half, 64, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
    if (bn == 128) return launch_persist<__half, 128, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
    if (bn == 160) return launch_persist<__half, 160, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
    return launch_persist<__half, 256, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
  } else {
    if (bn == 64) return launch_persist<__nv_bfloat16, 64, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
    if (bn == 128) return launch_persist<__nv_bfloat16, 128, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
    if (bn == 160) return launch_persist<__nv_bfloat16, 160, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
    return launch_persist<__nv_bfloat16, 256, CONV>(ta, tb, tcm, p, n_tiles_n, total_tiles, s);
  }
}

template <typename T, bool CONV, int PBN>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const GemmParams& p, int n_tiles_n,
                long long total_pair_tiles, cudaStream_t s) {
  using L = PairLayout<PBN>;
  static bool attr_set = false;
  if (!attr_set) {
    SS_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<T, CONV, PBN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 L::TOTAL));
    attr_set = true;
  }
  const int max_clusters = sm_count() / 2;
  const int clusters = (int)(total_pair_tiles < max_clusters ? total_pair_tiles : max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(GEMM_PERSIST_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  SS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<T, CONV, PBN>, ta, tb, tcm, p, n_tiles_n,
                             (int)total_pair_tiles));
  return 0;
}

template <bool CONV>
int dispatch_pair(int dtype, int pbn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm,
                  const GemmParams& p, int n_tiles_n, long long pair_tiles, cudaStream_t s) {
  (void)pbn;  // one pair tile width (256); a 160-wide pair tile measured no faster than the single-CTA 160 tile
  if (dtype == SS_F16) return launch_pair<__half, CONV, 256>(ta, tb, tcm, p, n_tiles_n, pair_tiles, s);
  return launch_pair<__nv_bfloat16, CONV, 256>(ta, tb, tcm, p, n_tiles_n, pair_tiles, s);
}

// Tile selection by a wave-quantisation cost model, in units of "accumulator columns of tensor-core time":
// a launch costs waves x (BN x smem-boundness + a fixed per-tile term).  The 160-wide tile exists because every
// SDXL channel count (320, 640, 1280, 2560, 5120, 10240) is a multiple of 160 but not always of 256, and
// M = 2048 rows x N = 1280 is 80 tiles of 256 (half the SMs idle) but 128 tiles of 160.
long long tile_cost(long long m_tiles, int N, int bn) {
  const long long tiles = m_tiles * ((N + bn - 1) / bn);
  const long long waves = (tiles + sm_count() - 1) / sm_count();
  const int eff = bn == 64 ? 150 : (bn == 128 ? 108 : 100);  // narrow tiles are shared-memory-bandwidth bound
  return waves * (bn * eff + 24 * 100);
}

long long pair_cost(long long m_tiles, int N, int pbn) {
  const long long tiles = ((m_tiles + 1) / 2) * ((N + pbn - 1) / pbn);
  const long long clusters = sm_count() / 2;
  const long long waves = (tiles + clusters - 1) / clusters;
  // per CTA the pair moves (128 + pbn/2) operand rows per K atom instead of (128 + bn): the short-K GEMMs are
  // L2->SM bandwidth bound, so the pair's main loop is rated 5 % (256) / 15 % (160) faster than a single-CTA tile
  return waves * (pbn * (pbn == 256 ? 95 : 85) + 24 * 100);
}

// N tile for the persistent kernel; a GLU epilogue pairs columns inside a 128-column fill, so it needs >= 128
int pick_bn_persist(long long m_tiles, int N, int glu, int force_bn) {
  if (force_bn == 64 || force_bn == 128 || force_bn == 160 || force_bn == 256)
    return (glu && force_bn == 64) ? 128 : force_bn;
  if (N <= 64 && !glu) return 64;
  int best = 256;
  long long best_cost = tile_cost(m_tiles, N, 256);
  const int cands[2] = {160, 128};
  for (int c : cands) {
    if (c == 160 && glu) continue;  // the 160 tile's register-stored tail measured slower under a GLU epilogue
    const long long k = tile_cost(m_tiles, N, c);
    if (k < best_cost) best = c, best_cost = k;
  }
  return best;
}

// CTA-pair tile width for this problem: 0 (use the single-CTA persistent kernel) or 256.  force_bn 1256 forces the
// pair kernel (test hook).
int pick_pair(long long m_tiles, int N, int glu, int force_bn) {
  (void)glu;
  if (force_bn == 1256) return N >= 256 ? 256 : 0;
  if (force_bn == 64 || force_bn == 128 || force_bn == 160) return 0;
  if (N < 256) return 0;
  const long long single = tile_cost(m_tiles, N, pick_bn_persist(m_tiles, N, glu, 0));
  const int pad256 = (N + 255) / 256 * 256;
  const long long pair_tiles256 = ((m_tiles + 1) / 2) * (pad256 / 256);
  if (pad256 * 100 <= N * 108 && pair_tiles256 >= 60) {
    if (force_bn == 256) return 256;
    if (pair_cost(m_tiles, N, 256) <= single) return 256;
  }
  return 0;
}

int get_out_tmap(CUtensorMap* out, const void* C, int dtype, long long M, int n_out, int ldc) {
  uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)M}, str[1] = {(uint64_t)ldc * 2};
  uint32_t box[2] = {64, 32};
  return get_tmap(out, C, dtype, 2, dims, str, box);
}

}  // namespace

// Row-statistics slots a GEMM with `stats_out` writes per output row: two per N tile of the kernel the dispatcher
// picks for this shape (one per epilogue group).
static int stat_slots_for(long long m_tiles, int N, int glu, int force_bn) {
  if (const int pbn = pick_pair(m_tiles, N, glu, force_bn)) return 2 * ((N + pbn - 1) / pbn);
  const int bn = pick_bn_persist(m_tiles, N, glu, force_bn);
  return 2 * ((N + bn - 1) / bn);
}

SS_API int ss_gemm_row_stat_slots(int M, int N) {
  return stat_slots_for((M + BM - 1) / BM, N, 0, 0);
}

static int gemm_tn_impl(int dtype, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                        const void* bias, const void* bias2, int rows_per_group, const void* residual, int ldr, int act,
                        int glu, float alpha, int force_bn, int flags, const float* ln_stats, int ln_slots,
                        float ln_eps, float* stats_out, void* stream) {
  SS_REQUIRE(dtype == SS_F16 || dtype == SS_BF16, "dtype must be f16 or bf16");
  SS_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  SS_REQUIRE(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0, "N, K, lda, ldb, ldc % 8");
  SS_REQUIRE(glu == 0 || (act == 0 && residual == nullptr), "GLU epilogue excludes act/residual");
  SS_REQUIRE(bias2 == nullptr || rows_per_group > 0, "bias2 needs rows_per_group");
  SS_REQUIRE(ln_stats == nullptr || (ln_slots > 0 && bias2 == nullptr), "folded LayerNorm needs its statistics slots");
  SS_REQUIRE(stats_out == nullptr || glu == 0, "row statistics are produced by non-GLU epilogues");
  const long long m_tiles = (M + BM - 1) / BM;
  const int bn = pick_bn_persist(m_tiles, N, glu, force_bn);
  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    if (int e = get_tmap(&ta, A, dtype, 2, dims, str, box)) return e;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = C;
  p.ldo = ldc;
  p.bias = bias;
  p.bias2 = bias2;
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  p.ld_b2 = N;
  p.residual = residual;
  p.ldr = ldr;
  p.act = act;
  p.glu = glu;
  p.alpha = alpha;
  p.M = M;
  p.N = N;
  p.K = K;
  p.b_const = (flags & 1 /* SS_GEMM_B_CONST */) ? 1 : 0;
  p.ln_stats = ln_stats;
  p.ln_slots = ln_slots;
  p.ln_eps = ln_eps;
  p.stats_out = stats_out;
  CUtensorMap tcm;
  if (int e = get_out_tmap(&tcm, C, dtype, M, glu ? N / 2 : N, ldc)) return e;
  if (const int pbn = pick_pair(m_tiles, N, glu, force_bn)) {
    CUtensorMap tb2;
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {BK, (uint32_t)pbn / 2};
    if (int e = get_tmap(&tb2, B, dtype, 2, dims, str, box)) return e;
    const int n_tiles_n = (N + pbn - 1) / pbn;
    const long long pair_tiles = ((m_tiles + 1) / 2) * n_tiles_n;
    return dispatch_pair<false>(dtype, pbn, ta, tb2, tcm, p, n_tiles_n, pair_tiles, (cudaStream_t)stream);
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    if (int e = get_tmap(&tb, B, dtype, 2, dims, str, box)) return e;
  }
  const int n_tiles_n = (N + bn - 1) / bn;
  return dispatch_persist<false>(dtype, bn, ta, tb, tcm, p, n_tiles_n, m_tiles * n_tiles_n, (cudaStream_t)stream);
}

// C[M,N] = epi(alpha * A[M,K] B[N,K]^T); see include/seedstory_b200.h for the argument contract.
SS_API int ss_gemm_tn(int dtype, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                      const void* bias, const void* bias2, int rows_per_group, const void* residual, int ldr, int act,
                      int glu, float alpha, int force_bn, int flags, void* stream) {
  return gemm_tn_impl(dtype, A, lda, B, ldb, C, ldc, M, N, K, bias, bias2, rows_per_group, residual, ldr, act, glu, alpha,
                      force_bn, flags, nullptr, 0, 0.f, nullptr, stream);
}

// The same GEMM with a LayerNorm folded around it (include/seedstory_b200.h): `ln_stats` applies LN to A's rows from
// the row statistics an earlier GEMM left (B must hold the row-centred gamma-scaled weights, `bias` = beta W^T + b);
// `stats_out` makes THIS GEMM leave the statistics of its output rows.
SS_API int ss_gemm_tn_ln(int dtype, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                         const void* bias, const void* residual, int ldr, int act, int glu, int flags,
                         const float* ln_stats, int ln_slots, float ln_eps, float* stats_out, void* stream) {
  return gemm_tn_impl(dtype, A, lda, B, ldb, C, ldc, M, N, K, bias, nullptr, 0, residual, ldr, act, glu, 1.f, 0, flags,
                      ln_stats, ln_slots, ln_eps, stats_out, stream);
}

// 3x3 stride-1 pad-1 convolution, NHWC activations [Nimg,H,W,Cin], weights [Cout, 9*Cin] with
// k = (ky*3+kx)*Cin + c, output NHWC [Nimg,H,W,Cout].
SS_API int ss_conv3x3_nhwc(int dtype, const void* x, const void* w, void* y, int Nimg, int H, int W, int Cin, int Cout,
                           const void* bias, const void* bias2 /*[Nimg, ld_bias2]*/, int ld_bias2,
                           const void* residual, int act, int force_bn, void* stream) {
  SS_REQUIRE(dtype == SS_F16 || dtype == SS_BF16, "dtype must be f16 or bf16");
  SS_REQUIRE(Cin % BK == 0, "Cin must be a multiple of 64 for the tensor-core conv");
  SS_REQUIRE(Cout % 8 == 0, "Cout must be a multiple of 8");
  int bw = W >= 128 ? 128 : W, bh = BM / bw;
  SS_REQUIRE(bw * bh == BM && W % bw == 0 && H % bh == 0, "image must tile into 128-pixel boxes");
  const long long m_tiles = (long long)Nimg * (H / bh) * (W / bw);
  const int bn = pick_bn_persist(m_tiles, Cout, 0, force_bn);
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {BK, (uint32_t)bw, (uint32_t)bh, 1};
    if (int e = get_tmap(&ta, x, dtype, 4, dims, str, box)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout}, str[1] = {(uint64_t)9 * Cin * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    if (int e = get_tmap(&tb, w, dtype, 2, dims, str, box)) return e;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = y;
  p.ldo = Cout;
  p.bias = bias;
  p.bias2 = bias2;
  p.rows_per_group = H * W;
  p.ld_b2 = ld_bias2 > 0 ? ld_bias2 : Cout;
  p.residual = residual;
  p.ldr = Cout;
  p.act = act;
  p.alpha = 1.f;
  p.M = Nimg * H * W;
  p.N = Cout;
  p.K = 9 * Cin;
  p.b_const = 1;  // convolution weights
  p.H = H;
  p.W = W;
  p.Cin = Cin;
  p.bw = bw;
  p.bh = bh;
  p.tiles_x = W / bw;
  p.tiles_y = H / bh;
  {
    CUtensorMap tcm;
    if (int e = get_out_tmap(&tcm, y, dtype, (long long)Nimg * H * W, Cout, Cout)) return e;
    if (const int pbn = pick_pair(m_tiles, Cout, 0, force_bn)) {
      CUtensorMap tb2;
      uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout}, str[1] = {(uint64_t)9 * Cin * 2};
      uint32_t box[2] = {BK, (uint32_t)pbn / 2};
      if (int e = get_tmap(&tb2, w, dtype, 2, dims, str, box)) return e;
      const int n_tiles_n2 = (Cout + pbn - 1) / pbn;
      const long long pair_tiles = ((m_tiles + 1) / 2) * n_tiles_n2;
      return dispatch_pair<true>(dtype, pbn, ta, tb2, tcm, p, n_tiles_n2, pair_tiles, (cudaStream_t)stream);
    }
    const int n_tiles_n = (Cout + bn - 1) / bn;
    return dispatch_persist<true>(dtype, bn, ta, tb, tcm, p, n_tiles_n, m_tiles * n_tiles_n, (cudaStream_t)stream);
  }
}
