"""Thin torch-tensor front end over the C-ABI (device pointers + current stream in, nothing else).

torch is used here for device memory and streams only; every function launches hand-written sm_100a
kernels from libseedstory_b200.so and raises if the library or a CUDA device is missing.
"""
import ctypes
import weakref

import torch

from . import _capi

F16, BF16 = 0, 1
KV_PAGE = 64
PROFILE = None  # bench.py sets this to a list: (name, algorithmic_flops, start_event, end_event) per tcgen05 launch
RECORD = None   # bench.py sets this to a list: (name, algorithmic_flops, relaunch) per tcgen05 launch; `relaunch()`
                # re-issues the identical call (same buffers), so each launch type can be timed back-to-back inside a
                # CUDA graph — eager event brackets include the host's launch gaps for kernels of a few microseconds


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(e0, name, flops):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((name, flops, e0, e1))


def _dt(t):
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"seedstory_b200 kernels take fp16/bf16 tensors, got {t.dtype}")


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _capi.SeedStoryError("seedstory_b200 ops need CUDA tensors (there is no CPU fallback)")


def require_device():
    n = ctypes.c_int(0)
    _capi.call("ss_require_device", ctypes.byref(n))
    return n.value


def rmsnorm(x, weight, eps, out=None):
    _req_cuda(x, weight)
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    out = torch.empty_like(x2) if out is None else out
    _capi.call("ss_rmsnorm_f16", _p(x2), x2.stride(0), _p(weight), _p(out), out.stride(0), x2.shape[0], K,
               ctypes.c_float(eps), _stream())
    return out.view(x.shape)


def layernorm(x, gamma, beta, eps, add=None, out=None, out2=None):
    """y = LN(x); if `add` [add_rows, K] is given also returns y2 = y + add[row % add_rows]."""
    _req_cuda(x, gamma)
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    assert x2.stride(1) == 1
    out = torch.empty((x2.shape[0], K), dtype=x.dtype, device=x.device) if out is None else out
    if add is not None:
        out2 = torch.empty_like(out) if out2 is None else out2
        add_rows = add.shape[0]
    else:
        add_rows = 0
    _capi.call("ss_layernorm", _dt(x), _p(x2), x2.stride(0), _p(gamma), _p(beta), _p(out), out.stride(0), x2.shape[0],
               K, ctypes.c_float(eps), _p(add), add_rows, _p(out2), out2.stride(0) if out2 is not None else 0,
               _stream())
    shp = x.shape
    if add is not None:
        return out.view(shp), out2.view(shp)
    return out.view(shp)


def l2norm_tokens(x):
    _req_cuda(x)
    B, T, C = x.shape
    x = x.contiguous()
    y = torch.empty_like(x)
    _capi.call("ss_l2norm_tokens_f16", _p(x), _p(y), B, T, C, _stream())
    return y


EPI_NONE, EPI_RESIDUAL, EPI_SWIGLU = 0, 1, 2


def skinny_gemm(x, W, epilogue=EPI_NONE, residual=None, out=None):
    """x [B<=8, K] @ W[N, K]^T with the decode epilogues."""
    _req_cuda(x, W)
    B, K = x.shape
    N = W.shape[0]
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    out = torch.empty((B, n_out), dtype=x.dtype, device=x.device) if out is None else out
    _capi.call("ss_skinny_gemm_f16", _p(x), x.stride(0), _p(W), _p(out), out.stride(0), B, N, K, epilogue,
               _p(residual), residual.stride(0) if residual is not None else 0, _stream())
    return out


def skinny_gemm_rmsnorm(x, gamma, eps, W, epilogue=EPI_NONE, out=None):
    """epilogue(RMSNorm(x; gamma, eps) @ W^T) for x [B<=8, K]: the normalisation runs in the kernel prologue."""
    _req_cuda(x, W)
    B, K = x.shape
    N = W.shape[0]
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    out = torch.empty((B, n_out), dtype=x.dtype, device=x.device) if out is None else out
    _capi.call("ss_skinny_gemm_rmsnorm_f16", _p(x), x.stride(0), _p(gamma), ctypes.c_float(eps), _p(W), _p(out),
               out.stride(0), B, N, K, epilogue, _stream())
    return out


def interleave_rope_rows(wqkv, H, D):
    """[q | k | v] projection rows -> the layout ss_decode_qkv_rope_append_f16 consumes: inside every q and k head row
    2i holds dim i and row 2i+1 dim i + D/2 (v rows unchanged)."""
    HD = H * D
    K = wqkv.shape[1]
    qk = wqkv[:2 * HD].view(2 * H, 2, D // 2, K).transpose(1, 2).reshape(2 * HD, K)
    return torch.cat([qk, wqkv[2 * HD:]], 0).contiguous()


def decode_rope_meta(tok_seq, tok_pos, tok_slot, page_table, cos_t, sin_t, H, D, kv_base, rope_cs, rope_sn):
    """Step constants of decode_qkv_rope_append: kv_base [B] int64, rope_cs / rope_sn [B, D] fp16 (written)."""
    B = tok_seq.numel()
    _capi.call("ss_decode_rope_meta", _p(tok_seq), _p(tok_pos), _p(tok_slot), B, _p(page_table), page_table.shape[1],
               _p(cos_t), _p(sin_t), H, D, _p(kv_base), _p(rope_cs), _p(rope_sn), _stream())


def decode_qkv_rope_append(x, gamma, eps, wqkv_il, q_out, kcache, vcache, kv_base, rope_cs, rope_sn, H, D):
    B, K = x.shape
    _capi.call("ss_decode_qkv_rope_append_f16", _p(x), x.stride(0), _p(gamma), ctypes.c_float(eps), _p(wqkv_il),
               _p(q_out), _p(kcache), _p(vcache), _p(kv_base), _p(rope_cs), _p(rope_sn), B, H, D, K, _stream())


def attn_decode_workspace(B, H, D, splits, device):
    """1024 int32 arrival counters (zeroed once; the kernel re-zeroes them) + partial results [B, H, splits, D+2] fp32."""
    assert B * H <= 1024
    return torch.zeros(1024 + B * H * splits * (D + 2), dtype=torch.float32, device=device)


def rope_kv_append(qkv, q_out, kcache, vcache, tok_seq, tok_pos, tok_slot, page_table, cos_t, sin_t, H, D):
    ntok = qkv.shape[0]
    _capi.call("ss_rope_kv_append_f16", _p(qkv), qkv.stride(0), _p(q_out), _p(kcache), _p(vcache), _p(tok_seq),
               _p(tok_pos), _p(tok_slot), ntok, _p(page_table), page_table.shape[1], _p(cos_t), _p(sin_t), H, D,
               _stream())


def attn_decode_paged(q, kcache, vcache, seq_lens, page_table, out, workspace, H, D, splits, scale):
    B = q.shape[0]
    _capi.call("ss_attn_decode_paged_f16", _p(q), _p(kcache), _p(vcache), _p(seq_lens), _p(page_table),
               page_table.shape[1], _p(out), _p(workspace), B, H, D, splits, ctypes.c_float(scale), _stream())


def logits_process_argmax(logits, last_ids, img_ids, next_ids, suppress_ids=None):
    B, V = logits.shape
    _capi.call("ss_logits_process_argmax_f16", _p(logits), logits.stride(0), V, _p(last_ids), _p(img_ids),
               img_ids.numel() if img_ids is not None else 0, _p(suppress_ids),
               suppress_ids.numel() if suppress_ids is not None else 0, _p(next_ids), B, _stream())


def gather_rows(table, ids, out):
    _capi.call("ss_gather_rows_16b", _p(table), _p(ids), _p(out), out.stride(0), ids.numel(), table.shape[1],
               _stream())


def decode_advance(next_ids, cur_ids, tok_pos, tok_slot, seq_lens, out_ids, n_out, done, eos_id, schedule=None):
    B = next_ids.numel()
    _capi.call("ss_decode_advance", _p(next_ids), _p(cur_ids), _p(tok_pos), _p(tok_slot), _p(seq_lens), _p(out_ids),
               out_ids.shape[1], _p(n_out), _p(done), eos_id, B, _p(schedule),
               schedule.shape[1] if schedule is not None else 0, _stream())


def store_rows_indexed(src, dst, idx):
    """dst [B, cap, W]; dst[b, idx[b]] = src[b]."""
    B, cap, W = dst.shape
    _capi.call("ss_store_rows_indexed_16b", _p(src), src.stride(0), _p(dst), cap, _p(idx), B, W, _stream())


def kv_gather_tokens(k_pages, v_pages, src_pages, dst_pages, src_idx, H, D):
    """k_pages/v_pages [layers, pages, H, 64, D]; page lists and src_idx are int32 device tensors."""
    layers = k_pages.shape[0]
    _capi.call("ss_kv_gather_tokens_16b", _p(k_pages), _p(v_pages), layers, ctypes.c_longlong(k_pages.stride(0)),
               _p(src_pages), _p(dst_pages), _p(src_idx), src_idx.numel(), H, D, _stream())


def kv_scatter_tokens(k_pages_layer, v_pages_layer, k, v, dst_pages):
    """k, v [H, n, D] (last dim contiguous, any head / token pitch) -> token slots 0..n-1 of dst_pages (int32, device)."""
    _req_cuda(k, v)
    H, n, D = k.shape
    assert v.shape == k.shape and k.stride(2) == 1 and v.stride(2) == 1 and k.dtype == v.dtype == torch.float16
    LL = ctypes.c_longlong
    _capi.call("ss_kv_scatter_tokens_16b", _p(k_pages_layer), _p(v_pages_layer), _p(k), _p(v), LL(k.stride(0)),
               LL(k.stride(1)), LL(v.stride(0)), LL(v.stride(1)), _p(dst_pages), n, H, D, _stream())


def lora_merge(W, A, B, scaling):
    out = torch.empty_like(W)
    N, K = W.shape
    _capi.call("ss_lora_merge_f16", _p(W), _p(A.contiguous()), _p(B.contiguous()), _p(out), N, K, A.shape[0],
               ctypes.c_float(scaling), _stream())
    return out


ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
GLU_NONE, GLU_GEGLU, GLU_SWIGLU = 0, 1, 2

# Load-time weight matrices (packed once by an engine, never written afterwards).  Only these may be fetched by a
# GEMM before its programmatic-dependent-launch wait (SS_GEMM_B_CONST); everything else is treated as an
# activation that an earlier kernel on the stream may still be writing.  Keyed by data pointer, validated through a
# weak reference so a freed weight whose address is reused by an activation is not mistaken for a constant.
_CONST_W = {}


def register_const(t):
    if isinstance(t, torch.Tensor) and t.is_cuda:
        _CONST_W[t.data_ptr()] = weakref.ref(t)
    return t


def register_const_tree(obj, skip=("kv_ctx",), _depth=0):
    """Register every CUDA tensor reachable through dicts / lists / tuples / plain objects under `obj`, except under
    keys or attributes named in `skip` (per-image buffers that kernels write)."""
    if _depth > 12:
        return
    if isinstance(obj, torch.Tensor):
        register_const(obj)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            if k not in skip:
                register_const_tree(v, skip, _depth + 1)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            register_const_tree(v, skip, _depth + 1)
    elif hasattr(obj, "__dict__") and not isinstance(obj, (torch.nn.Module, type)):
        for k, v in vars(obj).items():
            if k not in skip:
                register_const_tree(v, skip, _depth + 1)


def is_const_weight(w):
    r = _CONST_W.get(w.data_ptr())
    if r is None:
        return False
    t = r()
    return t is not None and t.data_ptr() == w.data_ptr()


class FoldedLN:
    """A LayerNorm folded into the Linear that consumes it (ss_gemm_tn_ln): y = LN(x; gamma, beta, eps) W^T + b is
    evaluated as rstd (x W''^T) + shift on the RAW rows x, with the row-centred weights
    W''[n,k] = gamma[k] W[n,k] - mean_k(gamma[k] W[n,k]) — x W''^T equals (x - mean(x)) (gamma (.) W)^T, the mean
    subtraction rides on the contraction — and shift[n] = sum_k beta[k] W[n,k] + b[n] (load-time packing, fp64 math,
    one rounding to the compute type)."""

    def __init__(self, w, gamma, beta, eps, bias=None):
        wf = w.double()                                   # load-time packing: fp64 math, one rounding to the compute type
        wg = wf * gamma.double()[None, :]
        self.w = (wg - wg.mean(dim=1, keepdim=True)).to(w.dtype).contiguous()
        shift = wf @ beta.double()
        if bias is not None:
            shift = shift + bias.double()
        self.shift = shift.to(w.dtype).contiguous()
        self.eps = float(eps)


def gemm_row_stat_slots(M, N):
    """Slots of (sum, sum of squares) a GEMM with `stats_out=` writes per output row for an [M, N] output."""
    return int(_capi.lib().ss_gemm_row_stat_slots(M, N))


def row_stats_buffer(M, N, device):
    slots = gemm_row_stat_slots(M, N)
    return torch.zeros((slots, M, 2), dtype=torch.float32, device=device)


def gemm(a, w, bias=None, bias2=None, rows_per_group=0, residual=None, act=ACT_NONE, glu=GLU_NONE, alpha=1.0,
         out=None, force_bn=0, w_const=None, ln=None, ln_stats=None, stats_out=None):
    """out[M, N or N/2] = epilogue(a[M,K] @ w[N,K]^T) on tcgen05 tensor cores.  w_const (SS_GEMM_B_CONST: `w` is a
    weight matrix nothing queued on the stream writes, so its first tiles may be fetched before the PDL wait)
    defaults to "is `w` a registered load-time weight" (register_const); unregistered operands are activations.
    ln=FoldedLN + ln_stats=[slots, M, 2]: `a` holds raw rows whose LayerNorm is folded into this GEMM (w must be
    ln.w); stats_out=[slots, M, 2]: leave the row statistics of this GEMM's output for a later folded LayerNorm."""
    _req_cuda(a, w)
    if w_const is None:
        w_const = is_const_weight(w)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if glu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=a.dtype, device=a.device)
    assert out.stride(1) == 1
    if ln is not None or stats_out is not None:
        assert bias2 is None and alpha == 1.0 and force_bn == 0
        if ln is not None:
            assert bias is None and ln_stats is not None and ln_stats.shape[1] == M and w.data_ptr() == ln.w.data_ptr()
        if stats_out is not None:
            assert tuple(stats_out.shape) == (gemm_row_stat_slots(M, N), M, 2) and stats_out.dtype == torch.float32

        ln_bias = ln.shift if ln is not None else bias

        def launch():
            _capi.call("ss_gemm_tn_ln", _dt(a), _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K,
                       _p(ln_bias), _p(residual), residual.stride(0) if residual is not None else 0, act, glu,
                       1 if w_const else 0, _p(ln_stats) if ln is not None else None,
                       ln_stats.shape[0] if ln is not None else 0, ctypes.c_float(ln.eps if ln is not None else 0.0),
                       _p(stats_out), _stream())
    else:
        def launch():
            _capi.call("ss_gemm_tn", _dt(a), _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K,
                       _p(bias), _p(bias2), rows_per_group, _p(residual),
                       residual.stride(0) if residual is not None else 0, act, glu, ctypes.c_float(alpha), force_bn,
                       1 if w_const else 0, _stream())
    _e = _prof_begin()
    launch()
    name = f"gemm {M}x{N}x{K}" + (" glu" if glu else "") + (" +res" if residual is not None else "") + \
        (" ln" if ln is not None else "") + (" +stats" if stats_out is not None else "")
    _prof_end(_e, name, 2.0 * M * N * K)
    if RECORD is not None:
        RECORD.append((name, 2.0 * M * N * K, launch))
    return out


def unary(x, op, out=None):
    out = torch.empty_like(x) if out is None else out
    _capi.call("ss_unary", _dt(x), _p(x), _p(out), ctypes.c_longlong(x.numel()), op, _stream())
    return out


def conv3x3(x, w, bias=None, bias2=None, residual=None, act=ACT_NONE, out=None, force_bn=0):
    """x NHWC [N,H,W,Cin], w [Cout, 9*Cin] (tap-major), out NHWC [N,H,W,Cout]."""
    _req_cuda(x, w)
    Nimg, H, W_, Cin = x.shape
    Cout = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and w.shape[1] == 9 * Cin
    if out is None:
        out = torch.empty((Nimg, H, W_, Cout), dtype=x.dtype, device=x.device)
    def launch():
        _capi.call("ss_conv3x3_nhwc", _dt(x), _p(x), _p(w), _p(out), Nimg, H, W_, Cin, Cout, _p(bias), _p(bias2),
                   bias2.stride(0) if bias2 is not None else 0, _p(residual), act, force_bn, _stream())
    _e = _prof_begin()
    launch()
    name = f"conv3x3 {Nimg}x{H}x{W_} {Cin}->{Cout}" + (" +res" if residual is not None else "")
    _prof_end(_e, name, 2.0 * Nimg * H * W_ * Cout * 9 * Cin)
    if RECORD is not None:
        RECORD.append((name, 2.0 * Nimg * H * W_ * Cout * 9 * Cin, launch))
    return out


def fmha(q, k, v, out, B, H, Lq, Lk, D, q_strides, k_strides, v_strides, o_strides, scale, causal=False,
         kv_lens=None, page_table=None):
    """Strides are (batch, token, head) in elements; tensors may be views into fused buffers."""
    _req_cuda(q, k, v, out)
    LL = ctypes.c_longlong
    args = [_p(q), _p(k), _p(v), _p(out), B, H, Lq, Lk, D]
    for st in (q_strides, k_strides, v_strides, o_strides):
        args += [LL(int(st[0])), LL(int(st[1])), LL(int(st[2]))]
    args += [_p(kv_lens), _p(page_table), page_table.shape[1] if page_table is not None else 0,
             ctypes.c_float(scale), 1 if causal else 0, _stream()]
    _capi.call("ss_fmha_f16", *args)
    if RECORD is not None:
        RECORD.append((f"fmha B{B} H{H} Lq{Lq} Lk{Lk} D{D}", 4.0 * B * H * Lq * Lk * D * (0.5 if causal and Lq == Lk else 1.0),
                       lambda: _capi.call("ss_fmha_f16", *args)))
    return out


def fmha_path_counts(reset=False):
    """(calls served by the tcgen05 FMHA kernels, calls served by the mma.sync kernel) since the last reset."""
    a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
    _capi.call("ss_fmha_path_counts", ctypes.byref(a), ctypes.byref(b), 1 if reset else 0)
    return a.value, b.value


def mha_packed(q, k, v, heads, scale, causal=False, out=None):
    """q [B, Lq, heads*D], k/v [B, Lk, heads*D] (last dim contiguous; may be column slices of fused buffers)."""
    B, Lq, E = q.shape
    Lk = k.shape[1]
    D = E // heads
    if out is None:
        out = torch.empty((B, Lq, E), dtype=q.dtype, device=q.device)
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    return fmha(q, k, v, out, B, heads, Lq, Lk, D, (q.stride(0), q.stride(1), D), (k.stride(0), k.stride(1), D),
                (v.stride(0), v.stride(1), D), (out.stride(0), out.stride(1), D), scale, causal)


def im2col_patch(img, P, Kpad):
    B, C, S, _ = img.shape
    G = S // P
    out = torch.empty((B * G * G, Kpad), dtype=img.dtype, device=img.device)
    _capi.call("ss_im2col_patch_f16", _p(img.contiguous()), _p(out), B, C, S, P, Kpad, _stream())
    return out


def add_bcast(x, add, out=None):
    C = x.shape[-1]
    rows = x.numel() // C
    out = torch.empty_like(x) if out is None else out
    _capi.call("ss_add_bcast", _dt(x), _p(x), _p(add), _p(out), ctypes.c_longlong(rows), C, add.numel() // C,
               _stream())
    return out


def scatter_rows(src, dst_rows, dst):
    _capi.call("ss_scatter_rows_16b", _p(src), _p(dst_rows), _p(dst), dst.stride(0), src.shape[0], src.shape[1],
               _stream())


def groupnorm_ws(N, HW, C, groups, device):
    n = _capi.lib().ss_groupnorm_ws_floats(N, HW, C, groups)
    return torch.zeros(n, dtype=torch.float32, device=device)   # zero ONCE: the first 64 words are arrival counters


def groupnorm_nhwc(x, gamma, beta, groups, eps, silu, stats_ws, out=None):
    N, H, W, C = x.shape
    need = _capi.lib().ss_groupnorm_ws_floats(N, H * W, C, groups)
    assert stats_ws.numel() >= need, (stats_ws.numel(), need)
    out = torch.empty_like(x) if out is None else out
    _capi.call("ss_groupnorm_nhwc", _dt(x), _p(x), _p(out), _p(gamma), _p(beta), _p(stats_ws), N, H * W, C, groups,
               ctypes.c_float(eps), 1 if silu else 0, _stream())
    return out


def upsample2x(x, out=None):
    N, H, W, C = x.shape
    if out is None:
        out = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
    _capi.call("ss_upsample2x_nhwc_16b", _p(x), _p(out), N, H, W, C, _stream())
    return out


def concat_channels(a, b, out=None):
    Ca, Cb = a.shape[-1], b.shape[-1]
    rows = a.numel() // Ca
    if out is None:
        out = torch.empty(a.shape[:-1] + (Ca + Cb,), dtype=a.dtype, device=a.device)
    _capi.call("ss_concat_channels_16b", _p(a), _p(b), _p(out), ctypes.c_longlong(rows), Ca, Cb, _stream())
    return out


def im2col3x3_s2(x, out=None):
    N, H, W, C = x.shape
    if out is None:
        out = torch.empty((N * (H // 2) * (W // 2), 9 * C), dtype=x.dtype, device=x.device)
    _capi.call("ss_im2col3x3_s2_nhwc_16b", _p(x), _p(out), N, H, W, C, _stream())
    return out


def cfg_euler_step(eps, latents, next_in, C, guidance, sigma, sigma_next):
    """eps [2, HW, Cpad] fp16; latents [HW, C] fp16 updated in place; next_in [2, HW, Cin_pad] or None."""
    HW = latents.shape[0]
    _capi.call("ss_cfg_euler_step_f16", _p(eps), eps.shape[-1], _p(latents), _p(next_in),
               next_in.shape[-1] if next_in is not None else 0, HW, C, ctypes.c_float(guidance),
               ctypes.c_float(sigma), ctypes.c_float(sigma_next), _stream())


def cast_scale(x, out_dtype, scale=1.0, out=None):
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device) if out is None else out
    _capi.call("ss_cast_scale", _dt(x), _p(x), _dt(out), _p(out), ctypes.c_longlong(x.numel()),
               ctypes.c_float(scale), _stream())
    return out


def softmax_rows_(x, scale=1.0):
    rows, n = x.shape
    _capi.call("ss_softmax_rows", _dt(x), _p(x), x.stride(0), rows, n, ctypes.c_float(scale), _stream())
    return x


def transpose2d(x, out=None):
    R, C = x.shape
    out = torch.empty((C, R), dtype=x.dtype, device=x.device) if out is None else out
    _capi.call("ss_transpose_16b", _p(x), _p(out), R, C, _stream())
    return out


def mean_tokens(x):
    B, T, C = x.shape
    y = torch.empty((B, C), dtype=x.dtype, device=x.device)
    _capi.call("ss_mean_tokens_f16", _p(x.contiguous()), _p(y), B, T, C, _stream())
    return y


def image_to_uint8(x, C):
    """x [pixels, ld] (first C channels valid) -> uint8 [pixels, C]."""
    pixels = x.shape[0]
    out = torch.empty((pixels, C), dtype=torch.uint8, device=x.device)
    _capi.call("ss_image_to_uint8", _dt(x), _p(x), x.stride(0), _p(out), ctypes.c_longlong(pixels), C, _stream())
    return out
