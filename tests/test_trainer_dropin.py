"""Build container (needs /root/reference; skipped elsewhere): the UNMODIFIED trainer methods of the reference
— `SDNetTrainer.predict` (Models/SDNetTrainer.py:378-451), `load_model` (:453-466), `save_for_predict`
(:492-509), `save` (:468-490) — driven through a stand-in trainer object (oracle/ref_harness.stand_in_trainer)
with `ruart_b200.Models.SDNet` as `self.network`.  No GPU here, so the network's forward is replaced by a
table of crafted probability rows where a forward is needed: what is under test is the caller side of the
drop-in boundary — checkpoint round trips, attribute access, and the answer-index rule,
whose device kernel is then checked on the GPU against goldens made HERE by the reference's own loop."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness
from ruart_b200 import synth

from helpers import GOLDEN, build_ours

needs_ref = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (build container only)")


def crafted_probability_rows():
    """Rows that exercise every branch of SDNetTrainer.py:402-412 — no-answer column on top, the `<OCR>` end
    slot on top (must be skipped), a masked slot >= num_cnt on top (fall through), a clear winner, the winner
    being the last legal index, num_cnt == 1 (only the end item)."""
    M1 = 12
    num_cnt = [5, 5, 4, 6, 8, 1, 3]
    rows = np.zeros((len(num_cnt), M1), np.float32)
    rows[0, :5] = [0.1, 0.2, 0.05, 0.15, 0.1]; rows[0, -1] = 0.4            # no-answer wins
    rows[1, :5] = [0.1, 0.2, 0.05, 0.15, 0.45]; rows[1, -1] = 0.05          # end slot (4) on top -> skipped -> 1
    rows[2, :4] = [0.1, 0.15, 0.05, 0.1]; rows[2, 7] = 0.5; rows[2, -1] = 0.1   # index >= num_cnt on top -> next
    rows[3, :6] = [0.02, 0.6, 0.08, 0.1, 0.1, 0.05]; rows[3, -1] = 0.05     # clear winner
    rows[4, :8] = [0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.5, 0.2]; rows[4, -1] = 0.09   # last legal index wins
    rows[5, 0] = 0.7; rows[5, -1] = 0.3                                      # only the end item: no-answer
    rows[6, :3] = [0.3, 0.25, 0.35]; rows[6, -1] = 0.1                       # end slot on top, then index 0
    return rows, num_cnt


class _TableNet(torch.nn.Module):
    """A network whose forward returns given probabilities (the trainer only needs `(scores, _)`)."""

    def __init__(self, probs):
        super().__init__()
        self.probs = probs
        self.drop_emb = None

    def forward(self, q, ocr, od):
        return self.probs, None


@needs_ref
def test_reference_predict_rule_equals_restatement_and_writes_device_goldens():
    ref_harness.import_reference()
    rows, num_cnt = crafted_probability_rows()
    probs = torch.from_numpy(rows)
    opt = synth.make_opt("tiny")
    fake = ref_harness.stand_in_trainer(_TableNet(probs), opt)
    ocr = {"num_cnt": num_cnt, "position": torch.zeros(len(num_cnt), rows.shape[1] - 1, 8)}
    extra = [{"q_id": i, "answers": None, "ocr_list": ["t%d" % k for k in range(n - 1)] + ["<OCR>"]}
             for i, n in enumerate(num_cnt)]
    gt = torch.zeros_like(probs)
    gt[:, 0] = 1
    _loss, _anls, _acc, res, save_res = fake.predict(({}, ocr, {}, gt, extra))
    picks = [int(r["idx"]) for r in save_res]
    assert fake.network.drop_emb is False and not fake.network.training      # predict() put the module in eval mode
    assert picks == synth.select_answers(probs, num_cnt)
    assert picks == [11, 1, 1, 1, 6, 11, 0]
    assert [r["answer"] for r in res] == ["unanswerable", "t1", "t1", "t1", "t6", "unanswerable", "t0"]
    path = os.path.join(GOLDEN, "select_answers_cases.json")
    data = {"probs": rows.tolist(), "num_cnt": num_cnt, "picks": picks,
            "made_by": "SDNetTrainer.predict (unmodified) via tests/test_trainer_dropin.py"}
    if os.path.exists(path):
        assert json.load(open(path))["picks"] == picks
    else:
        with open(path, "w") as f:
            json.dump(data, f)


@needs_ref
def test_reference_checkpoint_methods_round_trip_on_the_dropin(tmp_path):
    """save_for_predict drops `Bert*` keys; load_model deletes unknown keys and fills missing ones from the
    live state; a checkpoint written from the REFERENCE network loads into ours and vice versa."""
    ref_harness.import_reference()
    opt = synth.make_opt("tiny")
    ours, _ = build_ours("tiny", seed=1033)
    fake = ref_harness.stand_in_trainer(ours, opt)
    ck = str(tmp_path / "ours_predict.pt")
    fake.save_for_predict(ck)
    saved = torch.load(ck)
    assert not any(k.startswith("Bert") for k in saved["state_dict"]["network"]) and len(saved["state_dict"]["network"]) == 96
    # load it into a differently-initialised drop-in: every non-BERT tensor is restored, BERT untouched
    other, _ = build_ours("tiny", seed=7)
    bert_before = {k: v.clone() for k, v in other.state_dict().items() if k.startswith("Bert")}
    fake2 = ref_harness.stand_in_trainer(other, opt)
    fake2.load_model(ck)
    for k, v in ours.state_dict().items():
        if k.startswith("Bert"):
            assert torch.equal(other.state_dict()[k], bert_before[k])
        else:
            assert torch.equal(other.state_dict()[k], v), k
    # reference network -> checkpoint -> ours, and back
    ref = ref_harness.build_reference(opt, seed=55)
    rk = str(tmp_path / "ref_full.pt")
    rt = ref_harness.stand_in_trainer(ref, opt)
    rt.optimizer = torch.optim.Adamax([p for p in ref.parameters() if p.requires_grad], lr=1e-3)
    rt.save(rk, 0, str(tmp_path / "none.pt"))
    fake2.load_model(rk)
    for k, v in ref.state_dict().items():
        assert torch.equal(other.state_dict()[k], v), k
    back = ref_harness.build_reference(opt, seed=99)
    bt = ref_harness.stand_in_trainer(back, opt)
    bt.load_model(ck)                                    # the drop-in's save_for_predict file into the reference
    for k, v in ours.state_dict().items():
        if not k.startswith("Bert"):
            assert torch.equal(back.state_dict()[k], v), k
    # an unknown key in the file is ignored (load_model deletes it, SDNetTrainer.py:458-460)
    saved["state_dict"]["network"]["not.a.parameter"] = torch.zeros(1)
    torch.save(saved, ck)
    fake2.load_model(ck)
    # attributes the trainer touches from outside
    assert hasattr(ours, "drop_emb") and hasattr(ours, "Bert") and ours.fixed_embedding_fast.shape == (4000, 300)
    assert ours.fast_embed.weight.shape == (5000, 300) and ours.glove_embed.weight.requires_grad


def test_device_answer_goldens_exist_and_match_the_restatement():
    d = json.load(open(os.path.join(GOLDEN, "select_answers_cases.json")))
    assert synth.select_answers(torch.tensor(d["probs"]), d["num_cnt"]) == d["picks"]
