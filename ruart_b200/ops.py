"""Thin torch-tensor wrappers over the C ABI (include/ruart_b200.h).

PyTorch is used for device memory and streams only; every computation below is a hand-written
sm_100a kernel inside libruart_b200.so.  All wrappers raise if the tensors are not on CUDA.
"""
import numpy as np
import torch

from . import _lib
from ._lib import current_stream, ptr


def call(name, *args):
    return _lib.call(name, *args)

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_RELU_SCALE, EPI_BIAS_RELU = 0, 1, 2, 3, 4
PHOC_DIM = 604


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ruart_b200 kernels need CUDA tensors; there is no CPU fallback")


def round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------ PHOC
def phoc_batch(chars, offsets, out=None, packed=False, check=True):
    """chars uint8 [total], offsets int32 [n+1] (device) -> float32 [n, 604] (or uint32 [n,19]).

    `out` may be a column view [n, 604] of a wider row-major fp32 buffer (row pitch a multiple of
    4 floats).  Raises RuntimeError("Error: unigram X is unknown") like Utils/cphoc.c:45-50; with
    check=False the 8-byte device error word is returned instead of being read back (no sync).
    """
    _need_cuda(chars, offsets)
    n = offsets.numel() - 1
    dev = offsets.device
    err = torch.empty(1, dtype=torch.int64, device=dev)
    if out is not None and not packed and not out.is_contiguous():
        if out.dim() != 2 or out.shape != (n, PHOC_DIM) or out.stride(1) != 1 or out.dtype != torch.float32:
            raise ValueError("phoc_batch: out must be an fp32 [n, 604] view with unit column stride")
        call("ruart_phoc_batch_pitched", ptr(chars), ptr(offsets), n, ptr(out), out.stride(0), ptr(err),
             current_stream())
    elif packed:
        if out is None:
            out = torch.empty((n, 19), dtype=torch.int32, device=dev)
        call("ruart_phoc_batch_packed", ptr(chars), ptr(offsets), n, ptr(out), ptr(err),
             current_stream())
    else:
        if out is None:
            out = torch.empty((n, PHOC_DIM), dtype=torch.float32, device=dev)
        call("ruart_phoc_batch", ptr(chars), ptr(offsets), n, ptr(out), ptr(err), current_stream())
    if not check:
        return out, err
    key = int(err.item())
    if key != -1:
        raise RuntimeError("Error: unigram %s is unknown" % chr(key & 0xFF))
    return out


def phoc_strings(strings, device="cuda"):
    """list[str] -> float32 [n, 604] on `device` (host flattening + one kernel)."""
    enc = [s.encode("latin-1") for s in strings]
    offsets = np.zeros(len(enc) + 1, dtype=np.int32)
    if enc:
        offsets[1:] = np.cumsum([len(e) for e in enc])
    chars = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
    d_chars = torch.from_numpy(chars).to(device)
    d_off = torch.from_numpy(offsets).to(device)
    return phoc_batch(d_chars, d_off)


# ------------------------------------------------------------------------------------ GEMM
def gemm(a, w, M, N, Kp, *, a_parts=1, w_parts=1, n_terms=1, epi=EPI_NONE, bias=None, scale=None,
         out_f32=None, out_bf16=None, out_parts=1, out_part_stride=0, fast_gelu=False, residual=None):
    """out[M,N] = epi(a[M, parts*Kp] @ w[N, parts*Kp]^T); a, w bf16 row-major (last dim contiguous)."""
    _need_cuda(a, w, bias, scale, out_f32, out_bf16)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    lda = a.stride(0) if a.dim() == 2 else a.stride(-2)
    ldw = w.stride(0)
    call("ruart_gemm_bf16", ptr(a), lda, a_parts, ptr(w), ldw, w_parts, M, N, Kp, n_terms, epi,
         ptr(bias), ptr(scale), 0 if scale is None else scale.numel(),
         ptr(out_f32), 0 if out_f32 is None else out_f32.stride(-2),
         ptr(out_bf16), 0 if out_bf16 is None else out_bf16.stride(-2), out_parts,
         out_part_stride, int(fast_gelu), ptr(residual), 0 if residual is None else residual.stride(-2),
         current_stream())


# ------------------------------------------------------------------------------------ answers
def select_answers(probs, num_cnt, label_no_answer=True):
    """Device form of SDNetTrainer.predict's index rule (SDNetTrainer.py:402-412):
    probs fp32 [B, M+1] (device), num_cnt list/tensor [B] -> int32 [B] answer indices (device)."""
    _need_cuda(probs)
    B, M1 = probs.shape
    if not torch.is_tensor(num_cnt):
        num_cnt = torch.tensor(list(num_cnt), dtype=torch.int32)
    num_cnt = num_cnt.to(device=probs.device, dtype=torch.int32)
    out = torch.empty(B, dtype=torch.int32, device=probs.device)
    call("ruart_select_answers", ptr(probs.contiguous()), ptr(num_cnt), B, M1, 1 if label_no_answer else 0,
         ptr(out), current_stream())
    return out
