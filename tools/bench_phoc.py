"""PHOC featuriser (BASELINE config 2): 1M synthetic strings, kernel time vs the HBM roofline and the
CPU baselines (our C port, and the reference's own cphoc.c build when oracle/_ref exists)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import phoc_oracle  # noqa: E402
from ruart_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(2002)
lens = rng.integers(1, 21, size=n)
offsets = np.zeros(n + 1, np.int32)
offsets[1:] = np.cumsum(lens)
alpha = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", np.uint8)
chars = alpha[rng.integers(0, 36, size=int(offsets[-1]))]
d_c = torch.from_numpy(np.concatenate([chars, np.zeros(1, np.uint8)])).cuda()
d_o = torch.from_numpy(offsets).cuda()
out = torch.empty((n, 604), dtype=torch.float32, device="cuda")
packed = torch.empty((n, 19), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, kw, nbytes in (("dense_f32", dict(out=out), int(offsets[-1]) + 4 * (n + 1) + n * 604 * 4),
                         ("packed_bits", dict(out=packed, packed=True), int(offsets[-1]) + 4 * (n + 1) + n * 76)):
    for _ in range(3):
        ops.phoc_batch(d_c, d_o, **kw)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.phoc_batch(d_c, d_o, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    res[name] = {"ms": t, "strings_per_s": n / t * 1e3, "algorithmic_GBps": nbytes / t / 1e6}
peak = 6553.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
t0 = time.perf_counter()
want, _ = phoc_oracle.batch_flat(chars, offsets)
res["cpu_port_1core"] = {"s": time.perf_counter() - t0, "strings_per_s": n / (time.perf_counter() - t0)}
ref = phoc_oracle.ref_module()
if ref is not None:
    m = min(n, 200000)
    strs = [bytes(chars[offsets[i]:offsets[i + 1]]).decode() for i in range(m)]
    t0 = time.perf_counter()
    for s in strs:
        ref.build_phoc(s)
    dt = time.perf_counter() - t0
    res["cpu_reference_cphoc_1core"] = {"s": dt, "strings": m, "strings_per_s": m / dt}
assert np.array_equal(out.cpu().numpy(), want)
res["dense_f32"]["frac_of_hbm_peak"] = res["dense_f32"]["algorithmic_GBps"] / peak
res["hbm_peak_GBps"] = peak
res["bit_exact_vs_oracle"] = True
print(json.dumps(res, indent=1))
