"""Functional building blocks of the SDNet stack on top of the C ABI (fp32 activations).

Dense products go through the tcgen05 GEMM with split-bf16 operands (`parts` = 3 -> 6 partial
products, fp32-grade; 2 -> 3 products, ~2^-16; 1 -> plain bf16).  Everything else is one of the
fused kernels of csrc/sdnet_kernels.cu / csrc/lstm.cu.  No torch math here: torch allocates.
"""
import torch

from . import ops
from ._lib import current_stream, ptr
from .ops import call

_TERMS = {1: 1, 2: 3, 3: 6}
_consts = {}


_copy_streams = {}


def upload(arr, dev):
    """numpy array -> device tensor through pinned memory on a side stream: the copy neither waits
    for the kernels already queued on the compute stream nor blocks the host until they finish
    (a pageable cudaMemcpyAsync would do both)."""
    src = torch.from_numpy(arr)
    pinned = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    pinned.copy_(src)
    key = str(dev)
    cs = _copy_streams.get(key)
    if cs is None:
        cs = torch.cuda.Stream(device=dev)
        _copy_streams[key] = cs
    main = torch.cuda.current_stream(dev)
    with torch.cuda.stream(cs):
        d = pinned.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
    main.wait_event(ev)
    d.record_stream(main)
    return d


def rows2d(t):
    """View a [.., D] tensor whose leading dims are uniformly strided as (rows, cols, pitch)."""
    assert t.dtype == torch.float32 and t.stride(-1) == 1, "need fp32 rows with unit inner stride"
    if t.dim() == 1:
        return 1, t.shape[0], t.shape[0]
    pitch = t.stride(-2)
    rows = 1
    exp = pitch
    for d in range(t.dim() - 2, -1, -1):
        assert t.shape[d] == 1 or t.stride(d) == exp, "rows are not uniformly strided"
        exp = exp * t.shape[d]
        rows *= t.shape[d]
    return rows, t.shape[-1], pitch


def ones(dev):
    k = ("ones", str(dev))
    if k not in _consts:
        _consts[k] = torch.ones(1, dtype=torch.float32, device=dev)
    return _consts[k]


def ln_workspace(dev):
    k = ("lnws", str(dev), torch.cuda.current_stream().cuda_stream)
    if k not in _consts:
        _consts[k] = torch.empty(2048, dtype=torch.float64, device=dev)
    return _consts[k]


def split_act(x, parts, row_idx=None, n_rows=None):
    """fp32 rows -> bf16 split operand [rows, parts*Kp]."""
    rows, K, pitch = rows2d(x)
    if row_idx is not None:
        rows = n_rows if n_rows is not None else row_idx.numel()
    Kp = ops.round_up(K, 64)
    out = torch.empty((rows, parts * Kp), dtype=torch.bfloat16, device=x.device)
    call("ruart_split_bf16", ptr(x), pitch, ptr(row_idx), rows, K, Kp, parts, ptr(out), current_stream())
    return out, Kp


def split_concat(tensors, parts):
    """torch.cat(tensors, -1) -> bf16 split operand [rows, parts*Kp] without materialising the concatenation.
    tensors: fp32 [.., D_i] with equal leading dims (each may be a pitched view).  Returns (operand, Kp)."""
    import ctypes
    if len(tensors) == 1:
        return split_act(tensors[0], parts)
    assert len(tensors) <= 8
    infos = [rows2d(t) for t in tensors]
    rows = infos[0][0]
    assert all(i[0] == rows for i in infos), "concatenated tensors need equal row counts"
    K = sum(i[1] for i in infos)
    Kp = ops.round_up(K, 64)
    n = len(tensors)
    srcs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    pitches = (ctypes.c_longlong * n)(*[i[2] for i in infos])
    widths = (ctypes.c_int * n)(*[i[1] for i in infos])
    out = torch.empty((rows, parts * Kp), dtype=torch.bfloat16, device=tensors[0].device)
    call("ruart_split_concat_bf16", ctypes.addressof(srcs), ctypes.addressof(pitches), ctypes.addressof(widths), n,
         rows, Kp, parts, ptr(out), current_stream())
    return out, Kp


def _cache_of(owner):
    """Per-module cache (dies with the module: a global cache keyed by id() could be hit by a
    different module that re-uses the id and the freed CUDA addresses)."""
    c = owner.__dict__.get("_ruart_cache")
    if c is None:
        c = {}
        owner.__dict__["_ruart_cache"] = c
    return c


def prep_weight(owner, key, tensors, parts):
    """Cache the split-bf16 form of a weight matrix (concatenation of `tensors` along dim 0)."""
    _wcache = _cache_of(owner)
    ver = tuple((t.data_ptr(), t._version) for t in tensors) + (parts,)
    hit = _wcache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    w = torch.cat([t.detach().float() for t in tensors], 0).contiguous() if len(tensors) > 1 \
        else tensors[0].detach().float().contiguous()
    out, Kp = split_act(w, parts)
    _wcache[key] = (ver, out, Kp)
    return out, Kp


def prep_vector(owner, key, fn, tensors):
    _wcache = _cache_of(owner)
    ver = tuple((t.data_ptr(), t._version) for t in tensors)
    hit = _wcache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    v = fn().detach().float().contiguous()
    _wcache[key] = (ver, v)
    return v


def linear(a_split, Kp, w_split, rows, N, parts, out, epi=ops.EPI_NONE, bias=None, scale=None):
    """out[rows, N] (fp32, pitch from `out`) = epi(A W^T)."""
    _, _, pitch = rows2d(out)
    call("ruart_gemm_bf16", ptr(a_split), a_split.stride(0), parts, ptr(w_split), w_split.stride(0), parts,
         rows, N, Kp, _TERMS[parts], epi, ptr(bias), ptr(scale), 0 if scale is None else scale.numel(),
         ptr(out), pitch, None, 0, 1, 0, 0, None, 0, current_stream())
    return out


def gather_rows(src, src_idx, dst, dst_idx, n, D, dst2=None):
    _, _, sp = rows2d(src)
    _, _, dp = rows2d(dst)
    is64 = 1
    for ix in (src_idx, dst_idx):
        if ix is not None:
            is64 = 1 if ix.dtype == torch.int64 else 0
    d2p = rows2d(dst2)[2] if dst2 is not None else 0
    call("ruart_gather_rows", ptr(src), sp, ptr(src_idx), ptr(dst), dp, ptr(dst_idx), ptr(dst2), d2p, n, D,
         is64, current_stream())


def copy_cols(src, dst):
    """dst[..., :D] = src (row-wise copy between differently pitched fp32 buffers)."""
    rows, D, _ = rows2d(src)
    gather_rows(src, None, dst, None, rows, D)
    return dst


def concat_cols(tensors):
    """torch.cat(tensors, -1) for fp32 [.., D_i] tensors with equal leading dims, done by the row
    copy kernel (keeps the hot path free of library kernels)."""
    lead = tensors[0].shape[:-1]
    widths = [t.shape[-1] for t in tensors]
    out = torch.empty(tuple(lead) + (sum(widths),), dtype=torch.float32, device=tensors[0].device)
    col = 0
    for t, w in zip(tensors, widths):
        copy_cols(t, out[..., col:col + w])
        col += w
    return out


def whole_layernorm_(x, eps=1e-5):
    rows, cols, pitch = rows2d(x)
    call("ruart_whole_layernorm", ptr(x), rows, cols, pitch, eps, ptr(ln_workspace(x.device)), current_stream())
    return x


def attention_tail(p1, p2, mask_u8, x3, out, B, L1, L2, add=False, parts=3):
    _, hid, p1p = rows2d(p1)
    _, _, p2p = rows2d(p2)
    _, D3, x3p = rows2d(x3)
    _, _, op = rows2d(out)
    call("ruart_attention_tail", ptr(p1), p1p, ptr(p2), p2p, hid, ptr(mask_u8), ptr(x3), x3p, D3, ptr(out), op,
         B, L1, L2, 1 if add else 0, int(parts), current_stream())
    return out


def as_u8(mask):
    if mask.dtype == torch.uint8:
        return mask.contiguous()
    return mask.to(torch.uint8).contiguous()


def lstm_layer(owner, x, key, w_ih, w_hh, b_ih, b_hh, H, parts, out, whole_ln=False):
    """One (Bi)LSTM layer of StackedBRNN (Layers.py:156-170) for H <= 128 on [B, L, in] input.
    w_ih/w_hh/b_ih/b_hh are lists over directions.  `out` [B, L, ndir*H] may be a strided view.
    x may be a LIST of [B, L, D_i] tensors: the layer runs on their concatenation (never materialised)."""
    xs = list(x) if isinstance(x, (list, tuple)) else [x]
    B, L = xs[0].shape[0], xs[0].shape[1]
    ndir = len(w_ih)
    a, Kp = split_concat(xs, parts)
    w, _ = prep_weight(owner, (key, "w_ih"), w_ih, parts)
    bias = prep_vector(owner, (key, "bias"), lambda: torch.cat([bi + bh for bi, bh in zip(b_ih, b_hh)], 0), b_ih + b_hh)
    whh = prep_vector(owner, (key, "w_hh"), lambda: torch.stack(w_hh, 0), w_hh)
    xg = torch.empty((B * L, ndir * 4 * H), dtype=torch.float32, device=xs[0].device)
    linear(a, Kp, w, B * L, ndir * 4 * H, parts, xg, epi=ops.EPI_BIAS, bias=bias)
    _, _, op = rows2d(out)
    # fp32 mode (3-part operands): the fp32 FMA recurrence; otherwise the tensor-core one (bf16 hi|lo, ~2^-16)
    call("ruart_lstm_recurrence_f32" if parts == 3 else "ruart_lstm_recurrence", ptr(xg), xg.stride(0), ptr(whh),
         ptr(out), op, B, L, H, ndir, current_stream())
    if whole_ln:
        whole_layernorm_(out)
    return out
