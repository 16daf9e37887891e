/*
 * seedstory_b200 — C-ABI of the B200-native SEED-Story inference hot path.
 *
 * The reference (TencentARC/SEED-Story) has no FFI of its own: its "plugin API" is hydra `_target_`
 * paths plus duck-typed attribute access.  The host-side mirror of that API lives in
 * seed-story_b200/src/… (Python, same module paths and signatures); everything arithmetic behind it is
 * reached through the entry points declared here.  Each entry point cites the reference code whose
 * arithmetic it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - every function returns 0 on success; on failure it returns non-zero and ss_last_error() holds a
 *     thread-local message.  There is no CPU fallback anywhere.
 *   - all data pointers are DEVICE pointers (plain `void*` / `int*`), sizes are element counts,
 *     `stream` is a cudaStream_t passed as `void*` (NULL = default stream).  No torch types.
 *   - dtype tags: 0 = fp16, 1 = bf16.  Matrices are row-major; `ld*` are row pitches in elements.
 *   - KV cache pages hold 64 tokens: one page of one layer is [heads][64][head_dim] 16-bit values.
 */
#ifndef SEEDSTORY_B200_H
#define SEEDSTORY_B200_H

#ifdef __cplusplus
#define SS_EXPORT extern "C"
#else
#define SS_EXPORT
#endif

#define SS_DTYPE_F16 0
#define SS_DTYPE_BF16 1

#define SS_KV_PAGE_TOKENS 64

/* ---- library ------------------------------------------------------------------------------- */
SS_EXPORT const char* ss_last_error(void);
SS_EXPORT int ss_version(void);
/* fails unless an sm_100 device is current; writes its SM count */
SS_EXPORT int ss_require_device(int* sm_count_out);
SS_EXPORT int ss_stream_sync(void* stream);

/* ---- row normalisations --------------------------------------------------------------------- */
/* LlamaRMSNorm.forward — src/models_clm/modeling_llama_xformer.py:107-115 (fp32 variance, x*rsqrt
 * rounded to fp16, then fp16 weight multiply). */
SS_EXPORT int ss_rmsnorm_f16(const void* x, int ldx, const void* weight, void* y, int ldy, int rows, int K, float eps,
                             void* stream);
/* nn.LayerNorm as used at src/models/qwen_visual.py:143-147,353-354 and src/models_ipa/resampler.py:40-41;
 * optional second output y2 = LN(x) + add[row % add_rows] (the positional-embedding adds of
 * Resampler.forward, qwen_visual.py:146-148). */
SS_EXPORT int ss_layernorm(int dtype, const void* x, int ldx, const void* gamma, const void* beta, void* y, int ldy,
                           int rows, int K, float eps, const void* add, int add_rows, void* y2, int ldy2,
                           void* stream);
/* F.normalize(x) over dim=1 of [B,T,C] — src/models_ipa/resampler.py:269 */
SS_EXPORT int ss_l2norm_tokens_f16(const void* x, void* y, int B, int T, int C, void* stream);

/* ---- Llama decode step (batch <= 8) ---------------------------------------------------------- */
/* y[b,n] = sum_k x[b,k] W[n,k] for the q/k/v/o/gate/up/down/lm_head projections of a decode step —
 * modeling_llama_xformer.py:228-230, 296, 191, 759.  epilogue: 0 none, 1 y = residual + y (fp16 add,
 * :353/:359), 2 SwiGLU on interleaved (gate_j, up_j) row pairs -> y[b, j], j < N/2 (:191). */
SS_EXPORT int ss_skinny_gemm_f16(const void* x, int ldx, const void* W, void* y, int ldy, int B, int N, int K,
                                 int epilogue, const void* residual, int ldr, void* stream);
/* apply_rotary_pos_emb (:165-173) on q and k, then append k (post-RoPE) and v to the paged cache
 * (replaces the torch.cat at :241-242).  qkv rows are [q | k | v], each H*D wide. */
SS_EXPORT int ss_rope_kv_append_f16(const void* qkv, int ld_qkv, void* q_out, void* kcache, void* vcache,
                                    const int* tok_seq, const int* tok_pos, const int* tok_slot, int ntok,
                                    const int* page_table, int max_pages, const void* cos_table,
                                    const void* sin_table, int H, int D, void* stream);
/* The same projection with LlamaRMSNorm (:107-115, identical rounding chain) recomputed in the kernel prologue:
 * y = epilogue(RMSNorm(x; gamma, eps) W^T) — input_layernorm -> q/k/v (:341, 228-230) and post_attention_layernorm ->
 * gate/up (:357-358, 191) of a decode step without a separate normalisation launch.  epilogue: 0 none, 2 SwiGLU. */
SS_EXPORT int ss_skinny_gemm_rmsnorm_f16(const void* x, int ldx, const void* gamma, float eps, const void* W, void* y,
                                         int ldy, int B, int N, int K, int epilogue, void* stream);
/* The attention input side of one decode layer in one launch (LlamaDecoderLayer.forward :341 + LlamaAttention.forward
 * :228-244): RMSNorm -> q/k/v projection -> apply_rotary_pos_emb -> q to q_out [B, H*D], k (post-RoPE) and v appended
 * to the paged cache.  Wqkv_il [3*H*D, K] = rows [q | k | v] where inside every q and k head the rows are
 * pair-interleaved (row 2i = dim i, row 2i+1 = dim i + D/2: a rotary pair sits in neighbouring rows); v rows natural.
 * kv_base [B], rope_cos / rope_sin [B, D]: the step constants ss_decode_rope_meta prepared (cache destination and
 * rotary factors of each sequence's new token), so the kernel's tail has no dependent index loads. */
SS_EXPORT int ss_decode_qkv_rope_append_f16(const void* x, int ldx, const void* gamma, float eps, const void* Wqkv_il,
                                            void* q_out, void* kcache, void* vcache, const long long* kv_base,
                                            const void* rope_cos, const void* rope_sin, int B, int H, int D, int K,
                                            void* stream);
/* Once per decode step (not per layer): kv_base[b] = element offset, inside one layer's page pool, of (page of
 * tok_slot[b], head 0, slot % 64, dim 0) — prepare_inputs_for_generation's position / cache bookkeeping (:796-852) —
 * and the cos / sin table rows of tok_pos[b] (:150-151). */
SS_EXPORT int ss_decode_rope_meta(const int* tok_seq, const int* tok_pos, const int* tok_slot, int B,
                                  const int* page_table, int max_pages, const void* cos_table, const void* sin_table,
                                  int H, int D, long long* kv_base, void* rope_cos, void* rope_sin, void* stream);
/* xops.memory_efficient_attention(q,k,v, LowerTriangularFromBottomRightMask) for q_len == 1 (:289-295)
 * over the pages retained in page_table (window + attention-sink pages); split over pages and merged by the last
 * CTA to arrive (one launch).  workspace: 1024 int32 arrival counters (B*H <= 1024; zero before the first call, the
 * kernel leaves them zero) followed by B*H*splits*(D+2) floats of partial results. */
SS_EXPORT int ss_attn_decode_paged_f16(const void* q, const void* kcache, const void* vcache, const int* seq_lens,
                                       const int* page_table, int max_pages, void* out, float* workspace, int B, int H,
                                       int D, int splits, float scale, void* stream);
/* AutoImageTokenGenerationProcessor.__call__ (src/models_clm/generation.py:19-31) + greedy argmax.
 * img_ids = [BOI, IMG_0..IMG_{n-3}, EOI]; NULL disables the processor. Logits are edited in place.
 * suppress_ids (optional, NULL = none): transformers' SuppressTokensLogitsProcessor placed after the image
 * processor in the `logits_processor=` list of src/models_clm/models.py:146-153 (scores[:, ids] = -inf). */
SS_EXPORT int ss_logits_process_argmax_f16(void* logits, int ld, int V, const int* last_ids, const int* img_ids,
                                           int n_img_ids, const int* suppress_ids, int n_suppress, int* next_ids,
                                           int B, void* stream);
/* embed_tokens lookup — modeling_llama_xformer.py:580-581 / models.py:127 */
SS_EXPORT int ss_gather_rows_16b(const void* table, const int* ids, void* out, int ld_out, int ntok, int width,
                                 void* stream);
/* greedy-loop bookkeeping on the device (append id, advance position/slot/length, EOS -> done) —
 * restates the sequence/position updates of HF greedy_search + prepare_inputs_for_generation (:827-844) */
SS_EXPORT int ss_decode_advance(const int* next_ids, int* cur_ids, int* tok_pos, int* tok_slot, int* seq_lens,
                                int* out_ids, int out_cap, int* n_out, int* done, int eos_id, int B,
                                const int* schedule, int sched_cap, void* stream);
/* keep the step's post-final-norm hidden row (src/models_clm/models.py:182-197 slices them afterwards) */
SS_EXPORT int ss_store_rows_indexed_16b(const void* src, int ld_src, void* dst, int cap, const int* idx, int B,
                                        int width, void* stream);
/* peft 0.4 LoRA Linear folded into the base weight: W' = W + (alpha/r) B A, fp32 accumulate, one rounding
 * (configs/clm_models/llama2chat7b_lora.yaml:7-27; wrap at src/models_clm/peft_models.py:49) */
SS_EXPORT int ss_lora_merge_f16(const void* W, const void* A, const void* B, void* out, int N, int K, int r,
                                float scaling, void* stream);

/* KV compaction for the image window / multimodal attention sink: destination slot i <- source slot src_idx[i]
 * in every layer (replaces the torch.cat slicing of src/inference/vis_george_sink.py:266-291 and the window cut of
 * src/inference/gen_george.py:235-239).  Source and destination page lists must be disjoint. */
SS_EXPORT int ss_kv_gather_tokens_16b(void* kpool, void* vpool, int layers, long long layer_stride,
                                      const int* src_pages, const int* dst_pages, const int* src_idx, int n, int H,
                                      int D, void* stream);

/* past_key_values ingestion (live sink-KV mode): one layer's K/V as [H, n, D] tensors with explicit head / token
 * pitches (elements) — the reference's tuple-of-(K, V) cache layout, modeling_llama_xformer.py:241-244, as sliced and
 * concatenated by src/inference/vis_george_sink.py:266-291 — written to token slots 0..n-1 of dst_pages. */
SS_EXPORT int ss_kv_scatter_tokens_16b(void* kpool_layer, void* vpool_layer, const void* k_src, const void* v_src,
                                       long long k_sh, long long k_st, long long v_sh, long long v_st,
                                       const int* dst_pages, int n, int H, int D, void* stream);

/* ---- dense contractions on tcgen05 ------------------------------------------------------------ */
/* C[M,N] = epi(alpha * A[M,K] B[N,K]^T): every nn.Linear on the prefill / ViT / resampler / UNet /
 * VAE paths (e.g. qwen_visual.py:191,233,258-260; resampler.py:58-76; modeling_llama_xformer.py:228-230).
 * epilogue order: +bias[N] -> round -> +bias2[row / rows_per_group, N] -> round -> act (1 gelu-erf,
 * 2 silu) -> round -> +residual -> round.  glu: 1 = first*gelu(second) (diffusers GEGLU), 2 =
 * silu(first)*second (LlamaMLP, :191) over interleaved column pairs, output width N/2.
 * force_bn: 0 = auto, else 64/128/160/256 (N tile of the single-CTA kernel; 256 may take the CTA-pair kernel);
 * 1256 forces the CTA-pair kernel (256-wide pair tile; test hook).
 * flags: SS_GEMM_B_CONST = B is a weight matrix that no work queued on `stream` writes (an nn.Linear weight): its
 * first tiles are then fetched while the preceding kernel is still draining.  Leave it clear when B is an
 * activation (e.g. the q k^T product of the VAE mid-block attention). */
#define SS_GEMM_B_CONST 1
SS_EXPORT int ss_gemm_tn(int dtype, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N,
                         int K, const void* bias, const void* bias2, int rows_per_group, const void* residual, int ldr,
                         int act, int glu, float alpha, int force_bn, int flags, void* stream);
/* The same GEMM with a LayerNorm folded around it (the LN -> Linear pairs of diffusers' BasicTransformerBlock that
 * src/models_ipa/adapter_modules.py:455-466 reaches: norm1 -> attn1.to_q/k/v, norm2 -> attn2.to_q, norm3 -> ff.net.0).
 * Consumer side (ln_stats != NULL): A holds the RAW rows x [M,K]; B holds the row-centred, gamma-scaled weights
 *   W''[n,k] = gamma[k] W[n,k] - mean_k(gamma[k] W[n,k]), for which x W''^T = (x - mean(x)) (gamma (.) W)^T, so that
 *   LN(x) W^T + b = rstd (x W''^T) + bias with bias[n] = sum_k beta[k] W[n,k] + b[n] (passed as `bias`).  rstd of row m
 *   is formed from ln_slots partial (sum, sum of squares) pairs ln_stats[slot][m][2] (fp32) left by the GEMM that
 *   wrote x.
 * Producer side (stats_out != NULL, non-GLU): this GEMM leaves those partials for ITS output rows (of the values as
 *   stored, after bias / activation / residual), ss_gemm_row_stat_slots(M, N) slots, layout [slot][M][2] fp32. */
SS_EXPORT int ss_gemm_tn_ln(int dtype, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N,
                            int K, const void* bias, const void* residual, int ldr, int act, int glu, int flags,
                            const float* ln_stats, int ln_slots, float ln_eps, float* stats_out, void* stream);
SS_EXPORT int ss_gemm_row_stat_slots(int M, int N);
/* 3x3 / stride 1 / pad 1 convolution on NHWC activations as an implicit GEMM (diffusers ResnetBlock2D
 * conv1/conv2, Up/Downsample convs, VAE decoder convs — SURVEY.md Appendix C).  w is [Cout, 9*Cin] with
 * k = (ky*3+kx)*Cin + c.  bias2 is the per-image time-embedding row [Nimg, Cout]. */
SS_EXPORT int ss_conv3x3_nhwc(int dtype, const void* x, const void* w, void* y, int Nimg, int H, int W, int Cin,
                              int Cout, const void* bias, const void* bias2, int ld_bias2, const void* residual,
                              int act, int force_bn, void* stream);

/* ---- fused attention ---------------------------------------------------------------------------- */
/* softmax(scale * Q K^T [+ bottom-right causal mask]) V, fp16, head_dim 64/128, explicit (batch, token, head)
 * strides in elements.  Replaces xops.memory_efficient_attention (modeling_llama_xformer.py:282-295, causal=1,
 * K/V through page_table), the ViT bmm/softmax/bmm (src/models/qwen_visual.py:208-217), nn.MultiheadAttention
 * inside Resampler (:146-149), PerceiverAttention / AttentionPool2d (src/models_ipa/resampler.py:68-73, 95-113)
 * and the diffusers Attention processors of the SDXL UNet.  kv_lens (device, optional) overrides Lk per batch
 * entry; with page_table != NULL k/v are page pools [page][H][64][D]. */
SS_EXPORT int ss_fmha_f16(const void* q, const void* k, const void* v, void* out, int B, int H, int Lq, int Lk, int D,
                          long long q_sb, long long q_sl, long long q_sh, long long k_sb, long long k_sl,
                          long long k_sh, long long v_sb, long long v_sl, long long v_sh, long long o_sb,
                          long long o_sl, long long o_sh, const int* kv_lens, const int* page_table, int max_pages,
                          float scale, int causal, void* stream);

/* test hook: how many ss_fmha_f16 calls were served by the tcgen05 kernels / by the mma.sync kernel since the last
 * reset (host-side counters; a layout the tcgen05 path declines falls back to mma.sync and shows up here) */
SS_EXPORT int ss_fmha_path_counts(long long* tc_calls, long long* mma_calls, int reset);

/* ---- bandwidth-bound glue ------------------------------------------------------------------------ */
/* conv1 patchify of the ViT (src/models/qwen_visual.py:347,382): NCHW image -> [B*G*G, Kpad] rows */
SS_EXPORT int ss_im2col_patch_f16(const void* img, void* out, int B, int C, int S, int P, int Kpad, void* stream);
/* y[r] = x[r] + add[r % period] — positional embedding adds (qwen_visual.py:387) */
SS_EXPORT int ss_add_bcast(int dtype, const void* x, const void* add, void* y, long long rows, int C, int period,
                           void* stream);
/* dst[dst_rows[i]] = src[i] — input_embeds[ids_cmp_mask] = image_embeds_lm[...] (src/models_clm/models.py:135) */
SS_EXPORT int ss_scatter_rows_16b(const void* src, const int* dst_rows, void* dst, int ld_dst, int n, int width,
                                  void* stream);
/* nn.GroupNorm(32, C) (+ optional fused SiLU) on NHWC — diffusers ResnetBlock2D / Transformer2DModel / VAE norms
 * (SURVEY.md Appendix C).  Deterministic (fixed-order) reduction; stats_ws holds ss_groupnorm_ws_floats() floats and
 * must be zero-filled ONCE before its first use (its first 64 words are arrival counters every call leaves at zero). */
SS_EXPORT int ss_groupnorm_ws_floats(int N, int HW, int C, int groups);
SS_EXPORT int ss_groupnorm_nhwc(int dtype, const void* x, void* y, const void* gamma, const void* beta,
                                float* stats_ws, int N, int HW, int C, int groups, float eps, int silu, void* stream);
/* Upsample2D's F.interpolate(scale_factor=2, mode="nearest") on NHWC */
SS_EXPORT int ss_upsample2x_nhwc_16b(const void* x, void* y, int N, int H, int W, int C, void* stream);
/* torch.cat([hidden, skip], dim=1) of the UNet up blocks, on NHWC rows */
SS_EXPORT int ss_concat_channels_16b(const void* a, const void* b, void* out, long long rows, int Ca, int Cb,
                                     void* stream);
/* im2col for Downsample2D's stride-2 3x3 conv (pad 1); feeds ss_gemm_tn */
SS_EXPORT int ss_im2col3x3_s2_nhwc_16b(const void* x, void* cols, int N, int H, int W, int C, void* stream);
/* classifier-free guidance + EulerDiscreteScheduler.step (eps-prediction) + scale_model_input of the next step —
 * the per-step glue of StableDiffusionXLPipeline.__call__ reached from src/models_ipa/adapter_modules.py:455-466 */
SS_EXPORT int ss_cfg_euler_step_f16(const void* eps, int eps_ld, void* latents, void* next_in, int in_ld, int HW,
                                    int C, float guidance, float sigma, float sigma_next, void* stream);
/* y = act(x) elementwise: 1 gelu(erf), 2 silu (e.g. the SiLU before ResnetBlock2D.time_emb_proj) */
SS_EXPORT int ss_unary(int dtype, const void* x, void* y, long long n, int op, void* stream);
SS_EXPORT int ss_cast_scale(int dtype_in, const void* x, int dtype_out, void* y, long long n, float scale,
                            void* stream);
/* in-place softmax(scale * row) — VAE mid-block single-head attention over 16384 tokens */
SS_EXPORT int ss_softmax_rows(int dtype, void* x, int ld, int rows, int n, float scale, void* stream);
SS_EXPORT int ss_transpose_16b(const void* x, void* y, int R, int C, void* stream);
/* x.mean(dim=tokens) — AttentionPool2d (src/models_ipa/resampler.py:92) */
SS_EXPORT int ss_mean_tokens_f16(const void* x, void* y, int B, int T, int C, void* stream);
/* VaeImageProcessor.postprocess: (x/2+0.5).clamp(0,1)*255 round -> uint8 HWC */
SS_EXPORT int ss_image_to_uint8(int dtype, const void* x, int ldx, void* out, long long pixels, int C, void* stream);

#endif /* SEEDSTORY_B200_H */
