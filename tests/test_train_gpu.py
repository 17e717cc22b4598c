"""GPU: the training step on the drop-in.  `SDNetTrainer.update`
(Models/SDNetTrainer.py:330-376) restated call by call — network.train(), drop_emb, forward, the
BCE-with-logits-on-probabilities loss (:510-518), backward through the hand-written backward kernels,
clip_grad_norm_ (:366), Adamax step (:367), TUNE_PARTIAL reset (:369-373) — checked against
  (a) the CPU oracle's autograd gradients, parameter by parameter, and
  (b) the golden vectors of ONE UNMODIFIED reference `update` (tests/golden/train_tiny.npz)."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sdnet_oracle, train_oracle
from oracle.gen_train_golden import grad_projections, make_targets, train_opt
from ruart_b200 import synth

from helpers import GOLDEN, build_ours

pytestmark = pytest.mark.gpu


def _update(net, opt, batch, targets, optimizer):
    """SDNetTrainer.update, line by line."""
    net.train()
    net.drop_emb = True
    q, ocr, od = synth.batch_to(copy.deepcopy(batch), "cuda")
    scores, _ = net(q, ocr, od)
    assert not torch.isnan(scores).any()
    loss = F.binary_cross_entropy_with_logits(scores, targets.cuda())
    if opt["loss"] == "BCE_D1":
        loss = loss * targets.size(1)
    optimizer.zero_grad()
    loss.backward()
    pre = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.requires_grad and p.grad is not None}
    total = torch.nn.utils.clip_grad_norm_(net.parameters(), opt["grad_clipping"])
    clipped = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.requires_grad and p.grad is not None}
    optimizer.step()
    if "TUNE_PARTIAL" in opt:
        net.fast_embed.weight.data[opt["tune_partial"]:] = net.fixed_embedding_fast
        net.glove_embed.weight.data[opt["tune_partial"]:] = net.fixed_embedding_glove
    return float(loss), pre, clipped, float(total)


def test_training_step_matches_oracle_gradients_and_reference_update():
    g = dict(np.load(os.path.join(GOLDEN, "train_tiny.npz")))
    opt = train_opt("tiny")
    net, _ = build_ours("tiny", seed=1033, bert_init="random", DROPOUT=0.0, dropout_emb=0.0, BERT_precision="fp32")
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}   # (Bert() puts its model on the GPU)
    batch = synth.make_batch("tiny", ragged=True)
    targets = make_targets(batch, opt["max_ocr_num"])
    want_loss, want = train_oracle.loss_and_grads(sd, opt, batch, targets)
    net.cuda()
    # `network.cuda()` leaves fixed_embedding_* behind on the host (SDNet.py:78-81): the reset copies them back
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    optimizer = torch.optim.Adamax([p for p in net.parameters() if p.requires_grad], lr=opt["lr"])
    loss, pre, clipped, total = _update(net, opt, batch, targets, optimizer)
    names = [str(n) for n in g["names"]]
    assert sorted(pre) == names == sorted(want) and len(names) == 89     # the unused GRUCell gets no gradient
    assert abs(loss - float(want_loss)) < 1e-5 * abs(float(want_loss))
    assert abs(loss - float(g["loss"])) < 1e-4 * float(g["loss"])
    # (a) every gradient against the CPU oracle's autograd (fp32, different summation orders)
    worst = {}
    for n in names:
        a, b = pre[n].double().cpu(), want[n].double()
        worst[n] = float((a - b).norm() / b.norm().clamp_min(1e-12))
    real = {n: float(want[n].norm()) > 1e-6 for n in names}     # (the bias of a softmax-ed score: pure rounding noise)
    bad = {n: e for n, e in worst.items() if real[n] and e > 1e-4}
    assert not bad, bad
    # (b) the unmodified reference's update: clipped gradient norms / projections of every tensor
    got_norm = np.asarray([float(clipped[n].norm()) for n in names])
    rl = g["clipped_grad_norm"] > 1e-6
    assert np.allclose(got_norm[rl], g["clipped_grad_norm"][rl], rtol=2e-4)
    if "clipped_grad_proj" in g:
        # <e, r> of an error vector e has standard deviation |e|: 4 sigma of |e| = 1e-4 |g|
        proj = np.asarray([grad_projections(n, clipped[n]) for n in names])
        assert (np.abs(proj - g["clipped_grad_proj"])[rl] <= 4e-4 * g["clipped_grad_norm"][rl][:, None] + 1e-9).all()
    assert np.allclose(clipped["alphaBERT"].cpu().numpy(), g["grad_alphaBERT"], rtol=1e-3, atol=1e-7)
    assert np.allclose(clipped["gammaBERT"].cpu().numpy(), g["grad_gammaBERT"], rtol=1e-3, atol=1e-7)
    assert np.allclose(clipped["get_answer.attn.linear.weight"][:8, :16].cpu().numpy(), g["grad_attn_w"], rtol=1e-3, atol=1e-6)
    assert np.allclose(clipped["multi2one.rnns.0.weight_hh_l0"][:8, :16].cpu().numpy(), g["grad_multi2one_whh"],
                       rtol=1e-3, atol=1e-6)
    rows = clipped["fast_embed.weight"].abs().sum(1).nonzero().flatten()[:32].cpu().numpy()
    assert np.array_equal(rows, g["grad_fast_rows"])
    # post-step weights
    after = net.state_dict()
    delta = np.asarray([float((after[n] - before[n]).norm()) for n in names])
    assert np.allclose(delta[rl], g["delta_norm"][rl], rtol=5e-3, atol=1e-7)
    after_sum = np.asarray([float(after[n].double().sum()) for n in names])
    assert np.allclose(after_sum, g["after_sum"], rtol=1e-4, atol=1e-2)
    k = opt["tune_partial"]
    assert torch.equal(after["fast_embed.weight"][k:], before["fast_embed.weight"][k:])
    assert not torch.equal(after["fast_embed.weight"][:k], before["fast_embed.weight"][:k])
    for n in after:
        if n.startswith("Bert."):
            assert torch.equal(after[n], before[n])                          # LOCK_BERT
    # the inference path sees the UPDATED weights (prepared-weight caches are keyed on parameter versions)
    net.eval()
    net.drop_emb = False
    with torch.no_grad():
        probs, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    sd2 = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    want_p, _, _ = sdnet_oracle.sdnet_forward(sd2, opt, *copy.deepcopy(batch))
    assert (probs.cpu() - want_p).abs().max().item() < 1e-4


def test_two_part_training_precision_tracks_the_fp32_grade_gradients():
    """SDNET_precision 'bf16x2' (what a bf16 BERT selects; bench.py --train): forward, dgrad and wgrad GEMMs on 2-part
    splits (~2^-16).  Same net, same batch, fp32 BERT in both runs: every gradient within 1e-3 of the 3-part run."""
    opt = train_opt("tiny")
    batch = synth.make_batch("tiny", ragged=True)
    targets = make_targets(batch, opt["max_ocr_num"]).cuda()
    res = {}
    for prec in ("fp32", "bf16x2"):
        net, _ = build_ours("tiny", seed=1033, device="cuda", DROPOUT=0.0, dropout_emb=0.0, BERT_precision="fp32",
                            SDNET_precision=prec)
        net.train()
        net.drop_emb = True
        scores, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
        loss = F.binary_cross_entropy_with_logits(scores, targets) * targets.size(1)
        named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        grads = torch.autograd.grad(loss, [p for _, p in named], allow_unused=True)
        res[prec] = (float(loss), {n: g_ for (n, _), g_ in zip(named, grads) if g_ is not None})
    from ruart_b200.Models import Layers
    assert Layers.train_parts == 2
    assert abs(res["fp32"][0] - res["bf16x2"][0]) < 1e-4 * abs(res["fp32"][0])
    assert res["fp32"][1].keys() == res["bf16x2"][1].keys()
    differs = False
    for n, a in res["fp32"][1].items():
        b = res["bf16x2"][1][n]
        if float(a.norm()) > 1e-6:
            assert float((a - b).norm() / a.norm()) < 1e-3, n
        differs = differs or not torch.equal(a, b)
    assert differs                                                # the 2-part path really ran


def test_differentiable_forward_equals_fused_inference_forward():
    net, opt = build_ours("small", seed=77, device="cuda", DROPOUT=0.0, dropout_emb=0.0, BERT_precision="fp32",
                          KEEP_LOGITS=True)
    batch = synth.make_batch("small", ragged=True)
    with torch.no_grad():
        p_eval, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    net.train()
    net.drop_emb = True
    p_train, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    assert p_train.requires_grad and p_train.shape == p_eval.shape
    assert (p_train.detach() - p_eval).abs().max().item() < 2e-5
    assert (p_train.detach()[p_eval == 0] == 0).all()


def test_flat_adamax_step_is_seen_by_the_next_forward():
    # FlatAdamax updates through raw pointers; the prepared-weight caches of the inference
    # path must not serve stale splits afterwards
    from ruart_b200.train_utils import FlatAdamax
    opt = train_opt("tiny")
    net, _ = build_ours("tiny", seed=1033, device="cuda", DROPOUT=0.0, dropout_emb=0.0, BERT_precision="fp32",
                        KEEP_LOGITS=True)
    batch = synth.make_batch("tiny", ragged=True)
    targets = make_targets(batch, opt["max_ocr_num"]).cuda()
    with torch.no_grad():
        p0, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    net.train()
    net.drop_emb = True
    scores, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    loss = F.binary_cross_entropy_with_logits(scores, targets) * targets.size(1)
    params = [p for p in net.parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    live = [(p, gr) for p, gr in zip(params, grads) if gr is not None]
    fa = FlatAdamax([p for p, _ in live], lr=0.01, max_norm=float(opt["grad_clipping"]))
    fa.step([gr for _, gr in live])
    net.eval()
    net.drop_emb = False
    with torch.no_grad():
        p1, _ = net(*synth.batch_to(copy.deepcopy(batch), "cuda"))
    assert (p1 - p0).abs().max().item() > 1e-4          # the step moved the output ...
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    want_p, want_l, _ = sdnet_oracle.sdnet_forward(sd, opt, *copy.deepcopy(batch))
    from helpers import rel_err
    assert rel_err(net.get_answer.last_logits.cpu(), want_l) < 1e-4   # ... to what the updated weights give
    assert (p1.cpu() - want_p).abs().max().item() < 5e-4
