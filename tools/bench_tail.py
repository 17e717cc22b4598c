"""Time ruart_attention_tail on the shapes of the SDNet stack (GPU): fp32 CUDA-core form (parts 3) vs
the tensor-core form (parts 2)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ruart_b200 import sdnet_ops as K  # noqa: E402

B = 256
res = []
for name, L1, L2, Hd, D3 in (("self-attention", 100, 100, 250, 250), ("deep attention (question keys)", 100, 40, 250, 250),
                            ("pre-align", 64, 40, 300, 300), ("OD <-> OCR", 100, 37, 125, 250)):
    p1 = torch.relu(torch.randn(B * L1, Hd, device="cuda")) * 0.3
    p2 = torch.relu(torch.randn(B * L2, Hd, device="cuda")) * 0.3
    x3 = torch.randn(B, L2, D3, device="cuda")
    m8 = torch.ones(B, L2, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, L1, D3, device="cuda")
    row = {"shape": name, "L1": L1, "L2": L2, "Hd": Hd, "D3": D3}
    for parts in (3, 2):
        for _ in range(3):
            K.attention_tail(p1, p2, m8, x3, out, B, L1, L2, parts=parts)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            K.attention_tail(p1, p2, m8, x3, out, B, L1, L2, parts=parts)
        e1.record()
        torch.cuda.synchronize()
        row["us_parts%d" % parts] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    res.append(row)
print(json.dumps(res, indent=1))
