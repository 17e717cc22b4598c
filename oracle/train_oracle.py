"""TEST INFRASTRUCTURE — CPU restatement of one training step of the reference
(`SDNetTrainer.update`, Models/SDNetTrainer.py:330-376) for SURVEY.md §8 row a-19, on top of the
differentiable form of oracle/sdnet_oracle.py.  No CUDA kernel for this row exists yet; this oracle
(pinned against the unmodified reference by oracle/gen_train_golden.py -> tests/golden/train_*.npz)
is what those kernels will be checked against.

  loss      `instance_bce_with_logits` (SDNetTrainer.py:510-518): binary cross-entropy WITH LOGITS applied
            to the network's output — which is already a softmax probability row (the reference's
            quirk) — times the number of columns for loss 'BCE_D1'
  backward  autograd through sdnet_forward_grad; trainable = every tensor except `Bert.*`
            (LOCK_BERT, SDNet.py:91-93); both word tables are trainable under TUNE_PARTIAL
  clip      torch.nn.utils.clip_grad_norm_(parameters, grad_clipping) (:364)
  step      torch.optim.Adamax(lr=opt['lr']) (:313, optimizer '#'), restated in `adamax_step`
  reset     rows >= tune_partial of the word tables are restored after the step (:367-371)

Dropout must be off for a deterministic comparison (DROPOUT 0, dropout_emb 0).
"""
import copy

import torch
import torch.nn.functional as F

from . import sdnet_oracle


def trainable_names(sd):
    """`requires_grad` parameters of the reference: everything outside `Bert.*` except the constant
    similarity scales (`AttentionScore.diagonal` with do_similarity, Layers.py:196-198: one element,
    requires_grad=False)."""
    return [k for k, v in sd.items() if not k.startswith("Bert.") and torch.is_floating_point(v)
            and not (k.endswith(".diagonal") and v.numel() == 1)]


def loss_fn(opt, scores, targets):
    loss = F.binary_cross_entropy_with_logits(scores, targets)
    if opt["loss"] == "BCE_D1":
        loss = loss * targets.size(1)
    return loss


def loss_and_grads(sd, opt, batch, targets):
    """(loss, {name: grad}) for the trainable tensors of the state dict `sd`."""
    work = {k: v.detach().clone() for k, v in sd.items()}
    names = trainable_names(work)
    for k in names:
        work[k].requires_grad_(True)
    probs, _, _ = sdnet_oracle.sdnet_forward_grad(work, opt, *copy.deepcopy(batch))
    loss = loss_fn(opt, probs, targets)
    grads = torch.autograd.grad(loss, [work[k] for k in names], allow_unused=True)
    # parameters the forward never uses (the GRUCell of GetFinalScores, Layers.py:395-397) keep grad None in
    # the reference: clip_grad_norm_ and the optimizer skip them, and so does this dict
    return loss.detach(), {k: g for k, g in zip(names, grads) if g is not None}


def clip_coefficient(grads, max_norm):
    """clip_grad_norm_: total L2 norm over all gradients, scale = max_norm / (norm + 1e-6) clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    return torch.clamp(max_norm / (total + 1e-6), max=1.0), total


def adamax_step(p, g, exp_avg, exp_inf, step, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adamax (no weight decay): in place on p, exp_avg, exp_inf; `step` counts from 1."""
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    torch.maximum(exp_inf * betas[1], g.abs() + eps, out=exp_inf)
    p.addcdiv_(exp_avg, exp_inf, value=-lr / (1 - betas[0] ** step))


def update(sd, opt, batch, targets, state=None):
    """One `SDNetTrainer.update`.  `sd` is modified in place; returns (loss, grad_norm, state)."""
    if state is None:
        state = {"step": 0, "exp_avg": {}, "exp_inf": {}}
    fixed = {}
    if "TUNE_PARTIAL" in opt:
        k = opt["tune_partial"]
        fixed = {"fast_embed.weight": sd["fast_embed.weight"][k:].clone(),
                 "glove_embed.weight": sd["glove_embed.weight"][k:].clone()}
    loss, grads = loss_and_grads(sd, opt, batch, targets)
    coef, norm = clip_coefficient(grads, float(opt["grad_clipping"]))
    state["step"] += 1
    lr = float(opt["lr"]) if "lr" in opt else 2e-3
    for name, g in grads.items():
        if name not in state["exp_avg"]:
            state["exp_avg"][name] = torch.zeros_like(g)
            state["exp_inf"][name] = torch.zeros_like(g)
        adamax_step(sd[name], g * coef, state["exp_avg"][name], state["exp_inf"][name], state["step"], lr)
    for name, rows in fixed.items():
        sd[name][opt["tune_partial"]:] = rows
    return loss, norm, state
