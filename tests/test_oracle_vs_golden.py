"""CPU: oracle/sdnet_oracle.py against the golden vectors produced by the UNMODIFIED reference
(tests/golden/model_*.npz, made by oracle/gen_model_golden.py), plus the state_dict contract."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sdnet_oracle
from ruart_b200 import synth

from helpers import CASES, GOLDEN, build_ours, load_golden, rel_err


def test_state_dict_names_shapes_order_match_reference():
    with open(os.path.join(GOLDEN, "state_dict_manifest.json")) as f:
        want = json.load(f)
    net, _ = build_ours("tiny")
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(got) == list(want)
    assert got == want
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 12246007


@pytest.mark.parametrize("name", ["tiny_uniform_random", "tiny_ragged_pretrained", "small_ragged_random"])
def test_oracle_reproduces_reference(name):
    cfg, ragged, init, seed = CASES[name]
    g = load_golden(name)
    net, opt = build_ours(cfg, seed=seed, bert_init=init)
    batch = synth.make_batch(cfg, ragged=ragged)
    probs, logits, inter = sdnet_oracle.sdnet_forward(net.state_dict(), opt, *batch, keep=True)
    assert rel_err(logits, g["logits"]) < 2e-5
    assert np.abs(probs.numpy() - g["probs"]).max() < 2e-5
    assert synth.select_answers(probs, batch[1]["num_cnt"]) == g["picks"].tolist()
    assert rel_err(inter["ocr_layers"][-1][:, :12, :16], g["context_rnn_ocr_last"]) < 1e-4
    assert rel_err(inter["ocr_high"][:, :12, :16], g["high_lvl_context_ocr"]) < 1e-4
    assert rel_err(inter["q_final"][:, :8, :16], g["ques_self_attn"]) < 1e-4


def test_oracle_reproduces_reference_at_baseline_config1_size():
    # BASELINE.json configs[0]: 32 questions x 20 q-tokens x 50 OCR tokens (+10 OD labels), full size
    from helpers import build_ours as _build
    g = load_golden("cfg1_uniform_random")
    net, opt = _build("cfg1", seed=1033, bert_init="random")
    batch = synth.make_batch("cfg1")
    probs, logits, _ = sdnet_oracle.sdnet_forward(net.state_dict(), opt, *batch)
    assert probs.shape == (32, 101)
    assert rel_err(logits, g["logits"]) < 2e-5
    assert np.abs(probs.numpy() - g["probs"]).max() < 2e-5
    assert synth.select_answers(probs, batch[1]["num_cnt"]) == g["picks"].tolist()


def test_oracle_reproduces_reference_at_long_ocr_config5_shape():
    # BASELINE.json configs[4] shape: 200 OCR tokens per image (max_ocr_num 201), question rows padded to
    # the 512-token BERT window; 4 ragged images
    from helpers import build_ours as _build
    g = load_golden("cfg5s_ragged_random")
    net, opt = _build("cfg5s", seed=1033, bert_init="random")
    batch = synth.make_batch("cfg5s", ragged=True)
    assert batch[0]["bert"].shape == (4, 512) and batch[1]["position"].shape == (4, 201, 8)
    probs, logits, _ = sdnet_oracle.sdnet_forward(net.state_dict(), opt, *batch)
    assert probs.shape == (4, 202)
    assert rel_err(logits, g["logits"]) < 2e-5
    assert synth.select_answers(probs, batch[1]["num_cnt"]) == g["picks"].tolist()


def test_oracle_phoc_channel_reproduces_reference():
    # opt PHOC + 'phoc' in ocr_embedding (SDNet.py:51-55,441-446); the golden's table came from the
    # reference's own cphoc, here it comes from the C restatement
    from oracle import phoc_oracle
    g = load_golden("tiny_ragged_phoc")
    table = phoc_oracle.vocab_table(synth.make_vocab_words(1033), use_ref=False)
    assert table.shape == (synth.VOCAB_SIZE, 604) and table[0].sum() > 0   # <PAD> row = PHOC("pad")
    net, opt = build_ours("tiny", phoc_table=table)
    assert list(net.state_dict())[:3] == ["alphaBERT", "gammaBERT", "phoc_embed.weight"]
    assert net.multi2one.rnns[0].weight_ih_l0.shape == (1200, 1992)
    batch = synth.add_phoc(synth.make_batch("tiny", ragged=True))
    probs, logits, _ = sdnet_oracle.sdnet_forward(net.state_dict(), opt, *batch)
    assert rel_err(logits, g["logits"]) < 2e-5
    assert synth.select_answers(probs, batch[1]["num_cnt"]) == g["picks"].tolist()


def test_answer_selection_rule():
    # SDNetTrainer.py:402-412: skip the <OCR> end slot, stop at no-answer, accept idx < num_cnt
    p = torch.tensor([[0.1, 0.5, 0.3, 0.0, 0.1], [0.1, 0.2, 0.6, 0.0, 0.1], [0.0, 0.1, 0.2, 0.0, 0.7]])
    assert synth.select_answers(p, [3, 3, 3]) == [1, 1, 4]


def test_shard_batch_partitions_questions_and_items():
    batch = synth.make_batch("small", ragged=True)
    parts = [synth.shard_batch(batch, r, 2) for r in range(2)]
    assert sum(len(p[1]["num_cnt"]) for p in parts) == len(batch[1]["num_cnt"])
    assert torch.equal(torch.cat([p[1]["fasttext"] for p in parts]), batch[1]["fasttext"])
    assert torch.equal(torch.cat([p[0]["bert"] for p in parts]), batch[0]["bert"])
    assert parts[0][2]["bert_offsets"] + parts[1][2]["bert_offsets"] == batch[2]["bert_offsets"]
