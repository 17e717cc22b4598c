"""One timed-region step of bench.py (same net, same batch) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off ...   (launch list / --set full captures under profiles/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ruart_b200 import synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dev = torch.device("cuda", 0)
net, opt = bench.build_net(cfg, dev)
batch = synth.batch_to(synth.make_batch(cfg, seed=2003), dev)
with torch.no_grad():
    for _ in range(3):
        net(*tuple(dict(d) for d in batch))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    net(*tuple(dict(d) for d in batch))
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("step ms", e0.elapsed_time(e1))
