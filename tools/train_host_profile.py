"""cProfile of the host side of one training step (forward under autograd + backward) at cfg-5 (GPU)."""
import cProfile
import contextlib
import io
import os
import pstats
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ruart_b200 import synth  # noqa: E402
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


for _ in range(3):
    step()
torch.cuda.synchronize()
import gc
gc.collect(); gc.freeze(); gc.disable()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue %.2f ms per step (5 steps), wall incl. sync %.2f" % (1e3 * (t1 - t0) / 5, 1e3 * (time.perf_counter() - t0) / 5))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:6000])
