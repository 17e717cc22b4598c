"""Per-phase wall time of SDNet.forward (device-synchronised at phase boundaries) + host-only timing."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ruart_b200 import synth  # noqa: E402
from ruart_b200.bert_engine import flatten_offsets  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dev = torch.device("cuda", 0)
net, opt = bench.build_net(cfg, dev)
batch = synth.batch_to(synth.make_batch(cfg, seed=2003), dev)
with torch.no_grad():
    for _ in range(3):
        net(*tuple(dict(d) for d in batch))
    torch.cuda.synchronize()
    net.phase_log = []
    net(*tuple(dict(d) for d in batch))
    log = net.phase_log
    net.phase_log = None
    for (a, t0), (b, t1) in zip(log[:-1], log[1:]):
        print("%-16s %8.3f ms" % (b, 1e3 * (t1 - t0)))
    print("%-16s %8.3f ms (with phase syncs)" % ("total", 1e3 * (log[-1][1] - log[0][1])))
    # GPU-timeline phases (events on the main stream, streams enabled, no host syncs)
    for _ in range(2):
        net(*tuple(dict(d) for d in batch))
    torch.cuda.synchronize()
    net.phase_events = []
    e_start = torch.cuda.Event(enable_timing=True)
    e_start.record()
    net(*tuple(dict(d) for d in batch))
    e_end = torch.cuda.Event(enable_timing=True)
    e_end.record()
    torch.cuda.synchronize()
    evs = net.phase_events
    net.phase_events = None
    prev = e_start
    print("--- main-stream timeline (no syncs) ---")
    for lab, ev in evs:
        print("%-16s %8.3f ms" % (lab, prev.elapsed_time(ev)))
        prev = ev
    print("%-16s %8.3f ms" % ("tail", prev.elapsed_time(e_end)))
    print("%-16s %8.3f ms" % ("total", e_start.elapsed_time(e_end)))
    t0 = time.perf_counter()
    for d in batch:
        flatten_offsets(d["bert_offsets"], d["bert"].shape[0])
    print("flatten_offsets x3: %.3f ms" % (1e3 * (time.perf_counter() - t0)))
    t0 = time.perf_counter()
    net._item_index(batch[1]["num_cnt"], batch[1]["len_cnt"], 20, 100)
    net._item_index(batch[2]["num_cnt"], batch[2]["len_cnt"], 10, 37)
    print("_item_index x2: %.3f ms" % (1e3 * (time.perf_counter() - t0)))
    # async run: host time to enqueue vs device time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net.check_nan = False
    net(*tuple(dict(d) for d in batch))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue (host) %.3f ms, then wait %.3f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
