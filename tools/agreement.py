"""Answer agreement of the CUDA path (bf16 and fp32 mode) with the fp32 CPU oracle at the full
cfg-3 batch (256 questions; the whole-tensor LayerNorm makes the batch the unit of comparison)."""
import copy
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import build_ours, rel_err  # noqa: E402
from oracle import sdnet_oracle  # noqa: E402
from ruart_b200 import synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
torch.set_num_threads(os.cpu_count() or 1)
net, opt = build_ours(cfg, seed=1033, bert_init="random", device="cuda", KEEP_LOGITS=True)
batch = synth.make_batch(cfg, seed=2003)
cpu_sd = {k: v.cpu() for k, v in net.state_dict().items()}
t0 = time.perf_counter()
want_p, want_l, _ = sdnet_oracle.sdnet_forward(cpu_sd, opt, *copy.deepcopy(batch), bert_max_rows=2048)
t_cpu = time.perf_counter() - t0
want_pick = synth.select_answers(want_p, batch[1]["num_cnt"])
res = {"cfg": cfg, "questions": len(want_pick), "cpu_oracle_seconds": t_cpu, "weights": "reference-style random init, seed 1033"}
for mode in ("bf16", "fp32", "bf16_sdnet1"):
    net.Bert.precision = "fp32" if mode == "fp32" else "bf16"
    net.sdnet_parts = {"bf16": 2, "fp32": 3, "bf16_sdnet1": 1}[mode]   # bf16_sdnet1: plain bf16 SDNet GEMM operands
    with torch.no_grad():
        probs, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    lg = net.get_answer.last_logits.cpu()
    picks = synth.select_answers(probs, batch[1]["num_cnt"])
    res[mode] = {"logit_rel_err": rel_err(lg, want_l), "max_abs_dprob": float((probs.cpu() - want_p).abs().max()),
                 "answer_agreement": sum(int(a == b) for a, b in zip(picks, want_pick)) / len(picks),
                 "argmax_agreement": float((probs.cpu().argmax(1) == want_p.argmax(1)).float().mean())}
print(json.dumps(res, indent=1))
