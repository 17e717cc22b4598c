"""Time the persistent BiLSTM recurrence kernel on the shapes of the SDNet stack (GPU).

    python tools/bench_lstm.py            # B=256, H=125: L=100 (OCR), 37 (OD), 40 (question)
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from ruart_b200._lib import current_stream, ptr  # noqa: E402
from ruart_b200.ops import call  # noqa: E402


def main():
    dev = "cuda"
    H = 125
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    res = []
    for L in (100, 37, 40):
        xg = torch.randn(B * L, 8 * H, device=dev) * 0.5
        w = (torch.rand(2, 4 * H, H, device=dev) * 2 - 1) / H ** 0.5
        out = torch.empty(B * L, 2 * H, device=dev)
        st = current_stream()
        run = lambda: call("ruart_lstm_recurrence", ptr(xg), 8 * H, ptr(w), ptr(out), 2 * H, B, L, H, 2, st)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        gates = torch.empty(B * L, 10 * H, device=dev)
        ref = torch.empty_like(out)
        old = lambda: call("ruart_lstm_recurrence_train", ptr(xg), 8 * H, ptr(w), ptr(ref), 2 * H, B, L, H, 2,
                           ptr(gates), 10 * H, st)
        old()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            old()
        e1.record()
        torch.cuda.synchronize()
        us_old = e0.elapsed_time(e1) / 5 * 1e3
        res.append({"B": B, "L": L, "H": H, "us": round(us, 1), "us_per_step": round(us / L, 3),
                    "fma_kernel_with_saved_gates_us": round(us_old, 1),
                    "max_abs_diff": float((out - ref).abs().max())})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
