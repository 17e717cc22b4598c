"""Micro-benchmark of the fused query/key/value + attention kernel (ruart_qkv_attention_fold) at cfg-3's token
mix next to the unfused pair (folded QKV GEMM + ruart_bert_attention).  Usage (GPU box): python tools/bench_qkv_attn.py"""
import sys

import torch

sys.path.insert(0, ".")
from ruart_b200._lib import call, current_stream, ptr  # noqa: E402

torch.manual_seed(0)
H, heads = 768, 12
# cfg-3's mix: 256 questions of ~32 tokens, 13056 OCR items and 9472 OD labels of 3..8 tokens
lens = torch.cat([torch.randint(24, 42, (256,)), torch.randint(3, 9, (13056,)), torch.randint(3, 7, (9472,))])
cu = torch.zeros(lens.numel() + 1, dtype=torch.int32)
cu[1:] = torch.cumsum(lens, 0)
T, S = int(cu[-1]), lens.numel()
cu_d = cu.cuda()
raw = (torch.randn(T, H, device="cuda")).bfloat16()
stats = torch.zeros(T, 8, 2, device="cuda")
stats[:, 0, 1] = H
w = (torch.randn(3 * H, H, device="cuda") * 0.03).bfloat16()
v1, v2 = torch.randn(3 * H, device="cuda") * 0.1, torch.randn(3 * H, device="cuda") * 0.1
cap = 2 * T // 128 + 8
meta = torch.empty(cap, dtype=torch.int32, device="cuda")
bounds = torch.empty((T, 2), dtype=torch.int32, device="cuda")
st = current_stream()
call("ruart_seq_tiles", ptr(cu_d), S, ptr(meta), cap, ptr(bounds), st)
ctx = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
qkv = torch.empty(T, 3 * H, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
print("T=%d sequences=%d tiles=%d" % (T, S, int(meta[0])))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


fused = lambda: call("ruart_qkv_attention_fold", ptr(raw), H, ptr(w), H, T, H, heads, ptr(v1), ptr(v2), ptr(stats), 1e-12,
                     ptr(meta), ptr(bounds), ptr(ctx), H, st)
gemm = lambda: call("ruart_gemm_bf16_fold", ptr(raw), H, ptr(w), H, T, 3 * H, H, 1, 1, ptr(v1), ptr(v2), ptr(stats), 1e-12,
                    ptr(qkv), 3 * H, None, 0, None, st)
att = lambda: call("ruart_bert_attention", None, ptr(qkv), ptr(cu_d), S, heads, 0.125, 64, None, ptr(ctx), 1, st)
t_seq = timeit(lambda: call("ruart_seq_tiles", ptr(cu_d), S, ptr(meta), cap, ptr(bounds), st))
print("seq_tiles + token bounds: %.3f ms" % t_seq)
tf = timeit(fused)
tg, ta = timeit(gemm), timeit(att)
fl = 2.0 * T * 3 * H * H
print("fused qkv+attention: %.3f ms (%.0f TFLOP/s on the projection)  |  unfused: gemm %.3f + attention %.3f = %.3f ms"
      % (tf, fl / tf / 1e9, tg, ta, tg + ta))
