"""Target of the compute-sanitizer runs (profiles/r02_sanitizer_*.txt): a few forwards of the fused inference
path with the compute streams on (the first forward warms the weight caches and runs serially, the next ones use
the side streams), optionally one training step.   python tools/sanitize_forward.py <cfg> [train]"""
import contextlib
import copy
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ruart_b200 import synth  # noqa: E402
from ruart_b200.Models.SDNet import SDNet  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "small"
train = len(sys.argv) > 2 and sys.argv[2] == "train"
opt = synth.make_opt(cfg, BERT_precision="bf16", DROPOUT=0.0, dropout_emb=0.0, KEEP_LOGITS=True)
if os.environ.get("RUART_SANITIZE_BERT_LAYERS"):
    opt["BERT_num_layers"] = int(os.environ["RUART_SANITIZE_BERT_LAYERS"])
with contextlib.redirect_stdout(io.StringIO()):
    net = SDNet(opt, synth.make_embedding(1033))
synth.fill_state_dict(net, seed=1033)
net.cuda().eval()
net.drop_emb = False
batch = synth.make_batch(cfg, ragged=True)
outs = []
with torch.no_grad():
    for i in range(3):
        p, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
        outs.append(p.clone())
torch.cuda.synchronize()
assert net.use_streams and net._side is not None, "the side streams were not used"
assert torch.equal(outs[1], outs[2]), "forwards on the side streams are not deterministic"
print("forwards ok, max |p1 - p0| = %.3g" % float((outs[1] - outs[0]).abs().max()))
if train:
    import torch.nn.functional as F
    net.train()
    net.drop_emb = True
    scores, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    t = torch.zeros_like(scores)
    t[:, 0] = 1
    loss = F.binary_cross_entropy_with_logits(scores, t) * t.size(1)
    loss.backward()
    torch.cuda.synchronize()
    print("training step ok, loss %.4f" % float(loss))
