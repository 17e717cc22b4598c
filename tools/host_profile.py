"""cProfile of the HOST side of SDNet.forward at cfg-3 (what bounds `host_enqueue_ms_per_step` in bench.py):
   python tools/host_profile.py [steps]"""
import cProfile
import gc
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ruart_b200 import synth  # noqa: E402
from ruart_b200.Utils import collate  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
net, opt = bench.build_net("cfg3", dev, check_nan="deferred")
host = collate.attach_index_tensors(*synth.make_batch("cfg3", seed=2003, opt=opt))
batch = synth.batch_to(host, dev)
fresh = lambda b: tuple(dict(d) for d in b)
gc.collect()
gc.freeze()
with torch.no_grad():
    for _ in range(3):
        net(*fresh(batch))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        net(*fresh(batch))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host enqueue %.2f ms/step, total %.2f ms/step" % (1e3 * (t1 - t0) / steps, 1e3 * (t2 - t0) / steps))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        net(*fresh(batch))
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
