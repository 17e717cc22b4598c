"""Per-step device times of the bench loop (one CUDA event per step), with and without the nvidia-smi clock sampler:
   python tools/step_jitter.py [steps] — prints median / max step and the outliers (GPU)."""
import gc
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
net, opt = bench.build_net("cfg3", dev, check_nan="deferred")
from ruart_b200 import synth  # noqa: E402
from ruart_b200.Utils import collate  # noqa: E402
b = collate.attach_index_tensors(*synth.make_batch("cfg3", seed=2002, opt=opt))
dev_batch = synth.batch_to(b, dev)
fresh = lambda x: tuple(dict(d) for d in x)
gc.collect()
gc.freeze()
with torch.no_grad():
    for _ in range(5):
        net(*fresh(dev_batch))
    torch.cuda.synchronize()
    for label, period in (("no sampler", None), ("nvidia-smi -lms 50", 50), ("nvidia-smi -lms 500", 500), ("no sampler", None)):
        smp = None
        if period:
            import subprocess
            smp = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", str(period)],
                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            time.sleep(1.0)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        hs = []
        gc.disable()
        torch.cuda.synchronize()
        evs[0].record()
        for i in range(steps):
            t0 = time.perf_counter()
            net(*fresh(dev_batch))
            evs[i + 1].record()
            hs.append(1e3 * (time.perf_counter() - t0))
        net.check_pending()
        torch.cuda.synchronize()
        gc.enable()
        if smp:
            smp.terminate()
            smp.wait()
        d = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        sd = sorted(d)
        print("%-22s device ms/step: median %.2f mean %.2f max %.2f | steps > 1.15 x median: %s | host enqueue max %.1f ms" % (
            label, sd[len(sd) // 2], sum(d) / len(d), sd[-1],
            [(i, round(x, 1)) for i, x in enumerate(d) if x > 1.15 * sd[len(sd) // 2]], max(hs)))
