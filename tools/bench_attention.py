"""Time ruart_bert_attention (bf16 in/out) on the three segment shapes of a cfg-3 step (GPU):
questions (256 x 22..42 tokens), OCR items (13 056 x 3..8), object labels (9 472 x 3..6).

    python tools/bench_attention.py
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
    H, heads = 768, 12
    res = []
    for name, n, lo, hi in (("question", 256, 22, 42), ("ocr", 13056, 3, 8), ("od", 9472, 3, 6)):
        lens = rng.integers(lo, hi + 1, size=n)
        cu = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).to(dev)
        T = int(lens.sum())
        qkv = (torch.randn(T, 3 * H, device=dev) * 0.5).to(torch.bfloat16)
        out = torch.empty(T, H, dtype=torch.bfloat16, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        st = current_stream()
        run = lambda: call("ruart_bert_attention", None, ptr(qkv), ptr(cu), n, heads, 0.125, int(lens.max()),
                           None, ptr(out), 1, st)
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.zero_()   # evict qkv from the 126 MB L2, as inside a step
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        gb = T * 4 * H * 2 / 1e9
        res.append({"segment": name, "seqs": n, "tokens": T, "us": round(us, 1), "TB/s": round(gb / us * 1e3, 2)})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
