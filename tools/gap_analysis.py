"""GPU-timeline accounting of one forward: CUDA events around every C-ABI call on the main stream
(streams disabled): sum of kernel time per entry point and the idle time between calls."""
import contextlib
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ruart_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda", 0)
net, opt = bench.build_net("cfg3", dev)
net.use_streams = False
batch = synth.batch_to(synth.make_batch("cfg3", seed=2003), dev)
evs = []


class T(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *a):
        self.e1.record()
        evs.append((self.name, self.e0, self.e1))


with torch.no_grad():
    for _ in range(3):
        net(*tuple(dict(d) for d in batch))
    torch.cuda.synchronize()
    _lib.set_timing_hook(lambda name, a: T(name))
    s0 = torch.cuda.Event(enable_timing=True)
    s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    net(*tuple(dict(d) for d in batch))
    s1.record()
    torch.cuda.synchronize()
    _lib.set_timing_hook(None)
tot = s0.elapsed_time(s1)
busy = defaultdict(float)
cnt = defaultdict(int)
gaps = []
prev_end = s0
for name, e0, e1 in evs:
    busy[name] += e0.elapsed_time(e1)
    cnt[name] += 1
    gaps.append((prev_end.elapsed_time(e0), name))
    prev_end = e1
tail = prev_end.elapsed_time(s1)
print("total %.3f ms, in-call %.3f ms, between calls %.3f ms (tail %.3f)" % (tot, sum(busy.values()), sum(g for g, _ in gaps), tail))
for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print("  %-28s %4d calls %8.3f ms" % (k, cnt[k], v))
print("largest gaps (ms, before call):")
for g, n in sorted(gaps, reverse=True)[:15]:
    print("  %.3f  %s" % (g, n))
by = defaultdict(float)
for g, n in gaps:
    by[n] += g
print("gap time by following call:")
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-28s %8.3f ms" % (k, v))
