"""ctypes binding of libruart_b200.so (the C ABI declared in include/ruart_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  `lib()` loads the in-tree .so (building it with nvcc if it is absent
and a compiler is available).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libruart_b200.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "ruart_b200.h")
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_i64 = ctypes.c_int64
c_float = ctypes.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "ruart_last_error": [],
    "ruart_version": [],
    "ruart_num_sms": [],
    "ruart_phoc_batch": [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p],
    "ruart_phoc_batch_pitched": [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p],
    "ruart_phoc_batch_packed": [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p],
    "ruart_phoc_batch_host": [ctypes.c_char_p, c_void_p, c_i64, c_void_p,
                              ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_int32)],
    "ruart_gemm_bf16": [c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int,
                        c_int, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_ll, c_int,
                        c_ll, c_int, c_void_p, c_ll, c_void_p],
    "ruart_gemm_bf16_fold": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_float, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p],
    "ruart_qkv_attention_fold": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_float, c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "ruart_seq_tiles": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "ruart_bert_embed_raw": [c_void_p] * 5 + [c_int, c_int, c_void_p, c_void_p, c_void_p],
    "ruart_subword_coef": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "ruart_subword_avg_layers_fold": [c_void_p, c_ll, c_void_p, c_ll, c_float, c_void_p, c_int, c_void_p, c_void_p,
                                      c_int, c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "ruart_bert_embed_ln": [c_void_p] * 7 + [c_float, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "ruart_add_layernorm": [c_void_p] * 6 + [c_float, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "ruart_bert_attention": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p,
                             c_void_p, c_int, c_void_p],
    "ruart_subword_avg_accum": [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                c_ll, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    "ruart_subword_avg_layers": [c_void_p, c_void_p, c_ll, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                 c_ll, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "ruart_seq_lengths": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "ruart_seq_scan": [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_int, c_void_p],
    "ruart_pack_tokens": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "ruart_split_bf16": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p],
    "ruart_split_concat_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p],
    "ruart_gather_rows": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_ll, c_int,
                          c_int, c_void_p],
    "ruart_whole_layernorm": [c_void_p, c_ll, c_int, c_ll, c_float, c_void_p, c_void_p],
    "ruart_attention_tail": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_ll, c_int,
                             c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "ruart_attention_tail_heads": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_ll, c_int,
                                   c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    "ruart_self_attn_pool": [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_ll, c_void_p],
    "ruart_final_scores": [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "ruart_lstm_cell": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                        c_int, c_void_p, c_void_p, c_void_p],
    "ruart_lstm_recurrence": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p],
    "ruart_lstm_recurrence_f32": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p],
    "ruart_grad_sqnorm": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p],
    "ruart_adamax_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_int,
                          c_void_p, c_float, c_void_p],
    "ruart_select_answers": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    # ---- differentiable path (training step)
    "ruart_bmm_f32": [c_void_p, c_ll, c_ll, c_int, c_void_p, c_ll, c_ll, c_int, c_void_p, c_ll, c_ll, c_int,
                      c_int, c_int, c_int, c_float, c_int, c_void_p],
    "ruart_masked_softmax": [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_void_p],
    "ruart_softmax_backward": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p],
    "ruart_eltwise": [c_int, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_void_p, c_ll, c_ll, c_int, c_void_p],
    "ruart_mask_fill": [c_void_p, c_void_p, c_ll, c_float, c_void_p, c_void_p],
    "ruart_colsum": [c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "ruart_split_bf16_t": [c_void_p, c_ll, c_ll, c_int, c_ll, c_int, c_void_p, c_void_p],
    "ruart_whole_layernorm_stats": [c_void_p, c_ll, c_int, c_ll, c_float, c_void_p, c_void_p, c_void_p],
    "ruart_whole_layernorm_backward": [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p, c_void_p,
                                       c_ll, c_void_p],
    "ruart_embedding_grad_workspace_bytes": [c_ll, c_int, c_int],
    "ruart_embedding_grad": [c_void_p, c_int, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_int,
                             c_void_p],
    "ruart_subword_layers_backward": [c_void_p, c_void_p, c_ll, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                      c_ll, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                      c_void_p],
    "ruart_lstm_recurrence_train": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p,
                                    c_ll, c_void_p],
    "ruart_lstm_recurrence_backward": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int,
                                       c_int, c_void_p],
    "ruart_lstm_cell_train": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p],
    "ruart_lstm_cell_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_int, c_int, c_void_p],
}
_RESTYPES = {"ruart_last_error": ctypes.c_char_p, "ruart_embedding_grad_workspace_bytes": ctypes.c_longlong}


def declared_symbols():
    """Every function name declared in include/ruart_b200.h."""
    with open(_HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"RUART_API\s+[\w\s\*]+?\b(ruart_\w+)\s*\(", text)))


def lib_path():
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        from . import build as _build
        _build.build()
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError("ruart_b200: %s is missing and could not be built; there is no "
                           "non-CUDA fallback" % _LIB_PATH)
    L = ctypes.CDLL(_LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().ruart_last_error()
        raise RuntimeError("ruart_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


# kernels launched per C-ABI call (for bench.py's gpu_launches claim)
_KERNELS_PER_CALL = {"ruart_seq_tiles": 2, "ruart_whole_layernorm": 2, "ruart_whole_layernorm_stats": 2, "ruart_colsum": 2,
                     "ruart_whole_layernorm_backward": 2, "ruart_subword_layers_backward": 2, "ruart_embedding_grad": 6}
launch_count = 0
_timing_hook = None  # set by bench.py: callable(name, args) -> context manager, or None


def set_timing_hook(hook):
    global _timing_hook
    _timing_hook = hook


def call(name, *args):
    """Call a C-ABI function by name and raise on a non-zero return code."""
    global launch_count
    launch_count += _KERNELS_PER_CALL.get(name, 1)
    if _timing_hook is not None:
        with _timing_hook(name, args):
            check(getattr(lib(), name)(*args))
    else:
        check(getattr(lib(), name)(*args))


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else t.data_ptr()


_raw_stream = None


def current_stream():
    """cudaStream_t of torch's current stream on the current device, as an int.  torch.cuda.current_stream() costs
    ~14 us of Python per call and the forward asks ~170 times: the raw accessor is two C calls."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        get_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        get_dev = getattr(torch._C, "_cuda_getDevice", None)
        if get_stream is not None and get_dev is not None:
            _raw_stream = lambda: get_stream(get_dev())
        else:
            _raw_stream = lambda: torch.cuda.current_stream().cuda_stream
    return _raw_stream()
