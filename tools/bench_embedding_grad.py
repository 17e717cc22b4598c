"""ruart_embedding_grad: sorted form (counting sort + slab sums) against the first form (one warp per vocabulary
row scanning the id list; taken by passing the small workspace) on word-table shapes (GPU)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ruart_b200 import _lib  # noqa: E402
from ruart_b200._lib import current_stream, ptr  # noqa: E402
from ruart_b200.ops import call  # noqa: E402

D = 300
out = []
for V, n, live in ((5000, 66000, 0.3), (60000, 66000, 0.3), (60000, 400000, 0.3), (20, 66000, 0.3)):
    g = torch.Generator().manual_seed(V + n)
    ids = torch.randint(0, V, (n,), generator=g).cuda()
    dy = torch.randn(n, D, generator=g).cuda()
    dy[torch.rand(n, generator=g).cuda() > live] = 0
    dw = torch.empty(V, D, device="cuda")
    big = int(_lib.lib().ruart_embedding_grad_workspace_bytes(n, V, D))
    small = (n + 15) // 16 * 16 + (64 * V * D * 4 if V < 2048 else 0)
    rec = {"V": V, "n": n, "live_fraction": live}
    res = {}
    for name, nbytes in (("sorted_ms", big), ("scan_ms", small)):
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        run = lambda: call("ruart_embedding_grad", ptr(ids), 1, n, ptr(dy), D, D, V, ptr(ws), nbytes, ptr(dw), D, 0,
                           current_stream())
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        rec[name] = e0.elapsed_time(e1) / reps
        res[name] = dw.clone()
    rec["max_abs_diff"] = float((res["sorted_ms"] - res["scan_ms"]).abs().max())
    out.append(rec)
    print(json.dumps(rec), flush=True)
