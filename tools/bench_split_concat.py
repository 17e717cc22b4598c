"""Time ruart_split_concat_bf16 on the operand shapes of the SDNet stack (GPU).  RUART_SPLIT_CONCAT_FLAT=1 selects the
first (flat-index) kernel for comparison."""
import json
import sys

import torch

sys.path.insert(0, ".")
from ruart_b200 import sdnet_ops as K  # noqa: E402

rows = 256 * 101
res = []
for name, widths in (("self-attention input", (300, 768, 12, 8, 300, 250, 250)), ("deep-attention RNN input", (250,) * 5),
                     ("encoder layer 2 input", (1388, 250)), ("OD rows", (250, 250))):
    pieces = [torch.randn(rows, w, device="cuda") for w in widths]
    for _ in range(3):
        out, Kp = K.split_concat(pieces, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        K.split_concat(pieces, 3)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    nbytes = rows * sum(widths) * 4 + rows * 3 * Kp * 2
    res.append({"shape": name, "rows": rows, "K": sum(widths), "us": round(us, 1), "GBps": round(nbytes / us / 1e3)})
print(json.dumps(res))
