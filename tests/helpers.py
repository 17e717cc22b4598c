"""Shared test helpers: build our SDNet (parameter container) with seeded weights, load goldens."""
import contextlib
import io
import os

import numpy as np
import torch

from ruart_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "tiny_uniform_random": ("tiny", False, "random", 1033),
    "tiny_ragged_pretrained": ("tiny", True, "pretrained_like", 1033),
    "small_ragged_random": ("small", True, "random", 77),
    "small_uniform_pretrained": ("small", False, "pretrained_like", 5),
}


def build_ours(cfg, seed=1033, bert_init="random", device=None, phoc_table=None, **opt_over):
    """Our SDNet with weights from synth.fill_state_dict (identical to what the reference got).
    phoc_table: [V, 604] array -> the PHOC option of the reference (SDNet.py:51-55) on the OCR/OD side."""
    from ruart_b200.Models.SDNet import SDNet
    embedding = synth.make_embedding(seed)
    if phoc_table is not None:
        opt_over = dict(synth.PHOC_OPT, **opt_over)
        embedding["phoc_embedding"] = torch.as_tensor(phoc_table).clone()
    opt = synth.make_opt(cfg, **opt_over)
    with contextlib.redirect_stdout(io.StringIO()):
        net = SDNet(opt, embedding)
    synth.fill_state_dict(net, seed=seed, bert_init=bert_init)
    net.eval()
    net.drop_emb = False
    if device is not None:
        net.to(device)
    return net, opt


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, "model_%s.npz" % name)))


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    fin = torch.isfinite(b)
    assert torch.equal(fin, torch.isfinite(a)), "inf/nan pattern differs"
    if fin.sum() == 0:
        return 0.0
    return float((a[fin] - b[fin]).abs().max() / b[fin].abs().max().clamp_min(1e-30))
