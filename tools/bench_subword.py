"""Time ruart_subword_avg_layers (12 bf16 layer outputs -> weighted word means) on the OCR-item
shape of a cfg-3 step (GPU): 13 056 items, 1..2 words of 1..3 wordpieces.

    python tools/bench_subword.py
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ruart_b200._lib import current_stream, ptr  # noqa: E402
from ruart_b200.ops import call  # noqa: E402


def main():
    dev = "cuda"
    rng = np.random.default_rng(0)
    H, NL, N, W, XD = 768, 12, 13056, 20, 1388
    words, row_start, t = [], [], 0
    for item in range(N):
        row_start.append(t)
        pos = 1
        for j in range(int(rng.integers(1, 3))):
            n = int(rng.integers(1, 4))
            words.append((item, j, pos, pos + n))
            pos += n
        t += pos + 1
    T = t
    wt = torch.from_numpy(np.ascontiguousarray(np.array(words, dtype=np.int32).T)).to(dev)
    rs = torch.tensor(row_start, dtype=torch.int32, device=dev)
    hs = (torch.randn(NL, T, H, device=dev)).to(torch.bfloat16)
    dst = torch.zeros(N * W, XD, device=dev)
    alpha = torch.ones(NL, device=dev)
    gamma = torch.ones(1, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = current_stream()
    nw = wt.shape[1]
    run = lambda: call("ruart_subword_avg_layers", None, ptr(hs), T * H, ptr(wt), nw, ptr(rs), None, W,
                       dst.data_ptr() + 4 * 300, XD, ptr(alpha), NL, ptr(gamma), H, st)
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
    rows = sum(w[3] - w[2] for w in words)
    gb = (rows * NL * H * 2 + nw * H * 4) / 1e9
    print(json.dumps({"words": nw, "token_rows": rows, "us": round(us, 1), "TB/s": round(gb / us * 1e3, 2),
                      "checksum": float(dst.double().sum())}))


if __name__ == "__main__":
    main()
