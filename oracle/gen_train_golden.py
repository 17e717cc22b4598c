"""TEST INFRASTRUCTURE — golden vectors for the training step, produced by ONE
UNMODIFIED `SDNetTrainer.update` of the reference (oracle/ref_harness.run_reference_update) with all
dropout probabilities at 0.  Build container only:

    python -m oracle.gen_train_golden

Writes tests/golden/train_tiny.npz: loss, clipped-gradient statistics per parameter (pre-clip L2 norms),
slices of a few gradients, and per-parameter sums / norms of the weights after the Adamax step.
"""
import os

import numpy as np
import torch

from ruart_b200 import synth

from . import ref_harness


def make_targets(batch, M):
    """One-hot BCE targets over the valid OCR slots (seeded)."""
    g = torch.Generator().manual_seed(4242)
    num = batch[1]["num_cnt"]
    t = torch.zeros(len(num), M + 1)
    for b, n in enumerate(num):
        t[b, int(torch.randint(0, max(1, n - 1), (1,), generator=g))] = 1.0
    return t


N_PROJ = 8


def grad_projections(name, g):
    """N_PROJ seeded Rademacher projections <g, r_k> of one gradient tensor: a compact fingerprint of the WHOLE
    tensor (the full gradients are 49 MB), reproducible from the parameter name alone."""
    import zlib
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    flat = g.detach().double().reshape(-1).cpu()
    out = []
    for _ in range(N_PROJ):
        r = torch.randint(0, 2, (flat.numel(),), generator=gen, dtype=torch.int8).double() * 2 - 1
        out.append(float((flat * r).sum()))
    return out


def train_opt(cfg):
    return synth.make_opt(cfg, DROPOUT=0.0, dropout_emb=0.0)


def main():
    if not ref_harness.available():
        raise SystemExit("needs the reference tree (build container only)")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    cfg = "tiny"
    opt = train_opt(cfg)
    net = ref_harness.build_reference(opt, seed=1033, bert_init="random", bert_dropout=0.0)
    # shim (vi): `fixed_embedding_*` are VIEWS of the tensors that also became the embedding weights
    # (SDNet.py:54-67,78-81).  On a GPU `network.cuda()` moves the weights to new storage and leaves these
    # attributes behind, so the reset of SDNetTrainer.py:367-371 restores the original rows; with the CPU
    # harness's identity `.cuda()` they would alias the live weights and the reset would be a no-op.
    k = opt["tune_partial"]
    net.fixed_embedding_fast = net.fast_embed.weight.data[k:].clone()
    net.fixed_embedding_glove = net.glove_embed.weight.data[k:].clone()
    batch = synth.make_batch(cfg, ragged=True)
    targets = make_targets(batch, opt["max_ocr_num"])
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    loss, grads = ref_harness.run_reference_update(net, opt, batch, targets)
    after = net.state_dict()
    names = sorted(grads)
    # update() clipped the gradients in place: p.grad holds the CLIPPED values
    data = {"loss": np.float64(loss), "names": np.asarray(names),
            "clipped_grad_norm": np.asarray([float(grads[n].norm()) for n in names]),
            "clipped_grad_sum": np.asarray([float(grads[n].double().sum()) for n in names]),
            "clipped_grad_proj": np.asarray([grad_projections(n, grads[n]) for n in names]),
            "after_sum": np.asarray([float(after[n].double().sum()) for n in names]),
            "delta_norm": np.asarray([float((after[n] - before[n]).norm()) for n in names]),
            "grad_alphaBERT": grads["alphaBERT"].numpy(), "grad_gammaBERT": grads["gammaBERT"].numpy(),
            "grad_attn_w": grads["get_answer.attn.linear.weight"][:8, :16].numpy(),
            "grad_multi2one_whh": grads["multi2one.rnns.0.weight_hh_l0"][:8, :16].numpy(),
            "grad_fast_rows": grads["fast_embed.weight"].abs().sum(1).nonzero().flatten()[:32].numpy(),
            "fast_tail_unchanged": np.bool_(torch.equal(after["fast_embed.weight"][opt["tune_partial"]:],
                                                         before["fast_embed.weight"][opt["tune_partial"]:]))}
    path = os.path.join(out_dir, "train_%s.npz" % cfg)
    np.savez_compressed(path, **data)
    print("loss", loss, "params with grad", len(names), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
