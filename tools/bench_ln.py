"""Time ruart_add_layernorm (bf16 in -> bf16 out, residual already fused) at the cfg-3 token count (GPU)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ruart_b200._lib import current_stream, ptr  # noqa: E402
from ruart_b200.ops import call  # noqa: E402

T, H = 113664, 768
x = torch.randn(T, H, device="cuda").bfloat16()
out = torch.empty_like(x)
g = torch.rand(H, device="cuda") + 0.5
b = torch.randn(H, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = current_stream()
run = lambda: call("ruart_add_layernorm", None, ptr(x), None, None, ptr(g), ptr(b), 1e-12, T, H, None, ptr(out), 1, st)
for _ in range(3):
    run()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
us = float(np.median(ts))
print(json.dumps({"T": T, "us": round(us, 1), "TB/s": round(2 * T * H * 2 / us / 1e6, 2)}))
