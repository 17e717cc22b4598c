"""Micro-benchmark of the tcgen05 GEMM at the BERT shapes of BASELINE cfg-3 (valid tokens only).
Usage (GPU box): python tools/bench_gemm.py [M]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from ruart_b200 import ops  # noqa: E402

GELU = int(os.environ.get("RUART_GELU_MODE", "1"))
M = int(sys.argv[1]) if len(sys.argv) > 1 else 113664
# (name, N, K, epilogue, fused residual) — as the encoder calls them
shapes = [("qkv", 2304, 768, 1, False), ("attn_out+res", 768, 768, 1, True), ("ffn_up+gelu", 3072, 768, 2, False),
          ("ffn_down+res", 768, 3072, 1, True)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
tot_t = tot_f = 0.0
for name, N, K, epi, has_res in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(M, N, device="cuda").bfloat16() if has_res else None
    for _ in range(3):
        ops.gemm(a, w, M, N, K, epi=epi, bias=b, out_bf16=o, fast_gelu=GELU, residual=res)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, M, N, K, epi=epi, bias=b, out_bf16=o, fast_gelu=GELU, residual=res)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    fl = 2.0 * M * N * K
    tot_t += t
    tot_f += fl
    # cuBLAS for context
    for _ in range(3):
        torch.matmul(a, w.t())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    tc = e0.elapsed_time(e1) / 5
    print("%-12s M=%d N=%d K=%d  ours %.3f ms  %.1f TFLOP/s | cublas %.3f ms %.1f TFLOP/s" %
          (name, M, N, K, t, fl / t / 1e9, tc, fl / tc / 1e9))
print("layer total: %.3f ms, %.1f TFLOP/s ; x12 layers = %.1f ms" % (tot_t, tot_f / tot_t / 1e9, 12 * tot_t))

# ---- the same four GEMMs with the folded BertLayerNorm epilogues (ruart_gemm_bf16_fold)
from ruart_b200._lib import call, current_stream, ptr  # noqa: E402
H = 768
stats = torch.zeros(M, 8, 2, device="cuda")
stats[:, 0, 1] = H
stats2 = torch.zeros(M, 8, 2, device="cuda")
tot_t = 0.0
for name, N, K, epi, fold in [("qkv", 2304, 768, 1, 1), ("attn_out+res", 768, 768, 1, 2), ("ffn_up+gelu", 3072, 768, 2, 1),
                              ("ffn_down+res", 768, 3072, 1, 2)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    v1, v2 = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(M, N, device="cuda").bfloat16() if fold == 2 else None
    run = lambda: call("ruart_gemm_bf16_fold", ptr(a), K, ptr(w), K, M, N, K, fold, epi, ptr(v1), ptr(v2), ptr(stats), 1e-12,
                       ptr(o), N, ptr(res), N if fold == 2 else 0, ptr(stats2) if fold == 2 else None, current_stream())
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    tot_t += t
    print("fold %-12s N=%d K=%d  %.3f ms  %.1f TFLOP/s" % (name, N, K, t, 2.0 * M * N * K / t / 1e9))
print("folded layer total: %.3f ms ; x12 layers = %.1f ms (no LayerNorm passes)" % (tot_t, 12 * tot_t))
