"""Time per C-ABI entry point of one training step (forward with autograd + backward) at a given config (GPU):
CUDA events around every call, streams as in the product."""
import contextlib
import io
import os
import sys
from collections import defaultdict

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ruart_b200 import _lib, synth  # noqa: E402
from ruart_b200.Models.SDNet import SDNet  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
dev = torch.device("cuda", 0)
opt = synth.make_opt(cfg, BERT_precision="bf16", DROPOUT=0.0, dropout_emb=0.0)
with contextlib.redirect_stdout(io.StringIO()):
    net = SDNet(opt, synth.make_embedding(1033))
synth.fill_state_dict(net, seed=1033)
net.to(dev).train()
net.drop_emb = True
b = synth.make_batch(cfg, seed=2100, opt=opt)
batch = synth.batch_to(b, dev)
targets = bench.bce_targets(b, opt["max_ocr_num"], 4242).to(dev)
params = [p for n, p in net.named_parameters() if p.requires_grad and not n.startswith("get_answer.rnn.")]


def step():
    scores, _ = net(*tuple(dict(d) for d in batch))
    loss = F.binary_cross_entropy_with_logits(scores, targets) * targets.size(1)
    return torch.autograd.grad(loss, params)


for _ in range(2):
    step()
torch.cuda.synchronize()
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


_lib.set_timing_hook(lambda name, a: T(name))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter()
torch.cuda.profiler.start()   # ncu --profile-from-start off: the launch list of exactly this step
s0.record()
step()
s1.record()
torch.cuda.profiler.stop()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
_lib.set_timing_hook(None)
busy, cnt = defaultdict(float), defaultdict(int)
for name, e0, e1 in evs:
    busy[name] += e0.elapsed_time(e1)
    cnt[name] += 1
print("step %.2f ms (GPU events), host enqueue %.2f ms, then wait %.2f ms; in-call %.2f ms over %d calls" % (
    s0.elapsed_time(s1), 1e3 * (t1 - t0), 1e3 * (t2 - t1), sum(busy.values()), len(evs)))
for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print("  %-32s %5d calls %9.3f ms" % (k, cnt[k], v))
