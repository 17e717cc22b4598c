"""Error budget of bf16 mode vs the CPU oracle (cfg-1 shape, reduced batch, pretrained-like BERT)."""
import copy
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import build_ours, rel_err  # noqa: E402
from oracle import sdnet_oracle  # noqa: E402
from ruart_b200 import synth  # noqa: E402

cfg = dict(B=6, n_ocr=50, n_od=10, max_ocr_num=100, max_od_num=30)
for init, seed in (("pretrained_like", 11), ("random", 11), ("pretrained_like", 12)):
    net, opt = build_ours(cfg, seed=seed, bert_init=init, device="cuda", KEEP_LOGITS=True)
    batch = synth.make_batch(cfg, seed=2001, ragged=True)
    cpu_sd = {k: v.cpu() for k, v in net.state_dict().items()}
    want_p, want_l, _ = sdnet_oracle.sdnet_forward(cpu_sd, opt, *copy.deepcopy(batch))
    for mode, res32 in (("bf16", False), ("bf16", True), ("fp32", True)):
        net.Bert.precision = mode
        net.Bert.residual_fp32 = res32
        b = synth.batch_to(copy.deepcopy(batch), "cuda")
        with torch.no_grad():
            probs, _ = net(*b)
        lg = net.get_answer.last_logits.cpu()
        print("%-16s seed %d mode %s res32 %s: logit rel err %.3e  max|dprob| %.3e argmax agree %s" % (
            init, seed, mode, res32, rel_err(lg, want_l), (probs.cpu() - want_p).abs().max().item(),
            bool((probs.cpu().argmax(1) == want_p.argmax(1)).all())))
