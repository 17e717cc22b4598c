"""CPU: oracle/train_oracle.py (one `SDNetTrainer.update`) against the golden vectors
of the UNMODIFIED reference (tests/golden/train_tiny.npz, made by oracle/gen_train_golden.py), and the
Adamax restatement against torch.optim.Adamax.  No CUDA kernel exists for this row yet; this pins the
oracle those kernels will be checked against."""
import os

import numpy as np
import torch

from oracle import train_oracle
from oracle.gen_train_golden import make_targets, train_opt
from ruart_b200 import synth

from helpers import GOLDEN, build_ours


def test_training_step_oracle_reproduces_reference_update():
    g = dict(np.load(os.path.join(GOLDEN, "train_tiny.npz")))
    opt = train_opt("tiny")
    net, _ = build_ours("tiny", seed=1033, bert_init="random", DROPOUT=0.0, dropout_emb=0.0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    before = {k: v.clone() for k, v in sd.items()}
    batch = synth.make_batch("tiny", ragged=True)
    targets = make_targets(batch, opt["max_ocr_num"])
    names = [str(n) for n in g["names"]]
    assert sum(sd[n].numel() for n in train_oracle.trainable_names(sd)) == 12246007
    loss, grads = train_oracle.loss_and_grads(sd, opt, batch, targets)
    assert names == sorted(grads) and len(names) == 89               # the unused GRUCell gets no gradient
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    coef, total = train_oracle.clip_coefficient(grads, float(opt["grad_clipping"]))
    clipped = {n: grads[n] * coef for n in names}
    got_norm = np.asarray([float(clipped[n].norm()) for n in names])
    assert np.allclose(got_norm, g["clipped_grad_norm"], rtol=2e-3, atol=1e-7)
    assert np.allclose(clipped["alphaBERT"].numpy(), g["grad_alphaBERT"], rtol=2e-3, atol=1e-7)
    assert np.allclose(clipped["gammaBERT"].numpy(), g["grad_gammaBERT"], rtol=2e-3, atol=1e-7)
    assert np.allclose(clipped["get_answer.attn.linear.weight"][:8, :16].numpy(), g["grad_attn_w"], rtol=2e-3, atol=1e-6)
    assert np.allclose(clipped["multi2one.rnns.0.weight_hh_l0"][:8, :16].numpy(), g["grad_multi2one_whh"], rtol=2e-3, atol=1e-6)
    rows = clipped["fast_embed.weight"].abs().sum(1).nonzero().flatten()[:32].numpy()
    assert np.array_equal(rows, g["grad_fast_rows"])                # the same word rows receive gradient
    # the whole update: Adamax step on the clipped gradients, TUNE_PARTIAL reset
    loss2, norm2, state = train_oracle.update(sd, opt, batch, targets)
    assert state["step"] == 1 and abs(float(loss2) - float(g["loss"])) < 1e-4 * float(g["loss"])
    delta = np.asarray([float((sd[n] - before[n]).norm()) for n in names])
    # (a parameter whose gradient is rounding noise — the bias of a softmax-ed score — moves by
    # lr * g / (|g| + eps), which is noise too: compared only where the gradient is real)
    real = g["clipped_grad_norm"] > 1e-6
    assert real.sum() >= 87 and np.allclose(delta[real], g["delta_norm"][real], rtol=5e-3, atol=1e-7)
    after_sum = np.asarray([float(sd[n].double().sum()) for n in names])
    assert np.allclose(after_sum, g["after_sum"], rtol=1e-4, atol=1e-2)
    k = opt["tune_partial"]
    assert bool(g["fast_tail_unchanged"]) and torch.equal(sd["fast_embed.weight"][k:], before["fast_embed.weight"][k:])
    assert not torch.equal(sd["fast_embed.weight"][:k], before["fast_embed.weight"][:k])
    for name in sd:
        if name.startswith("Bert."):
            assert torch.equal(sd[name], before[name])              # LOCK_BERT


def test_adamax_restatement_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(50, 7)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adamax([ref], lr=1e-3)
    m, u = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        train_oracle.adamax_step(p, g, m, u, step, 1e-3)
        assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-7)
