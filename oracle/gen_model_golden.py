"""TEST INFRASTRUCTURE — golden vectors for the model path, made by running the UNMODIFIED
reference (oracle/ref_harness.py) on seeded synthetic inputs.  Build-container only:

    python -m oracle.gen_model_golden [case ...]

Writes tests/golden/model_<case>.npz with the reference's probabilities, pre-softmax logits,
selected answer indices (SDNetTrainer.py:402-412 rule) and a few slices of intermediate tensors.
Weights and inputs are NOT stored: both are regenerated from seeds by ruart_b200.synth
(fill_state_dict / make_batch), which is what the tests do.
"""
import os

import numpy as np
import torch

from ruart_b200 import synth

from . import phoc_oracle, ref_harness

CASES = {
    # name: (config, ragged, bert_init, weight seed)
    "tiny_uniform_random": ("tiny", False, "random", 1033),
    "tiny_ragged_pretrained": ("tiny", True, "pretrained_like", 1033),
    "small_ragged_random": ("small", True, "random", 77),
    "small_uniform_pretrained": ("small", False, "pretrained_like", 5),
    # PHOC channel on the OCR / OD side (opt PHOC, ocr_embedding phoc,...: SDNet.py:51-55,441-446)
    "tiny_ragged_phoc": ("tiny", True, "random", 1033),
    # BASELINE.json configs[0] at full size: 32 questions x 20 q-tokens x 50 OCR tokens (+10 OD labels)
    "cfg1_uniform_random": ("cfg1", False, "random", 1033),
    # BASELINE.json configs[4] shape (200 OCR tokens per image, question row padded to the 512-token
    # window) at 4 images
    "cfg5s_ragged_random": ("cfg5s", True, "random", 1033),
    # ---- the benchmarked configurations at FULL size; BERT_MAX_BatchSize set so the
    # reference fits in host memory (numerically neutral: BERT rows are independent, Bert.py:65-85);
    # no intermediate captures (the 12 x [N, W, 768] lists alone are > 9 GB)
    "cfg3_uniform_random": ("cfg3", False, "random", 1033),        # BASELINE configs[2]: what bench.py times
    "cfg3_ragged_pretrained": ("cfg3", True, "pretrained_like", 1033),  # B=256, realistic BERT activation scale
    "cfg4_shard3of8": ("cfg4", False, "random", 1033),             # BASELINE configs[3]: rank 3 of 8 (512 questions)
    "cfg5_uniform_random": ("cfg5", False, "random", 1033),        # BASELINE configs[4] forward at B=32
}
BIG_CASES = ("cfg3_uniform_random", "cfg3_ragged_pretrained", "cfg4_shard3of8", "cfg5_uniform_random")
SHARDS = {"cfg4_shard3of8": (3, 8)}
PHOC_CASES = ("tiny_ragged_phoc",)
CAPTURE = ("Bert", "multi2one", "context_rnn", "ques_rnn", "deep_attn", "high_lvl_context_rnn", "ques_self_attn")


def first(x):
    return x[0] if isinstance(x, (tuple, list)) else x


def main(only=()):
    if not ref_harness.available():
        raise SystemExit("needs the reference tree (build container only)")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    for name, (cfg, ragged, init, seed) in CASES.items():
        torch.manual_seed(0)
        if only and name not in only:
            continue
        embedding = None
        if name in PHOC_CASES:
            # the table is built by the reference's own cphoc (oracle/_ref) over the synthetic vocabulary
            opt = synth.make_opt(cfg, **synth.PHOC_OPT)
            embedding = synth.make_embedding(seed)
            embedding["phoc_embedding"] = torch.from_numpy(
                phoc_oracle.vocab_table(synth.make_vocab_words(seed), use_ref=True))
            batch = synth.add_phoc(synth.make_batch(cfg, ragged=ragged))
        else:
            opt = synth.make_opt(cfg)
            batch = synth.make_batch(cfg, ragged=ragged)
        if name in SHARDS:
            batch = synth.shard_batch(batch, *SHARDS[name])
        big = name in BIG_CASES
        if big:
            opt["BERT_MAX_BatchSize"] = 2048
        net = ref_harness.build_reference(opt, embedding=embedding, seed=seed, bert_init=init)
        # the forward runs INSIDE the unmodified SDNetTrainer.predict: `picks` are the reference's own
        # answer indices (SDNetTrainer.py:402-412), not a restatement
        probs, logits, picks, cap, _ = ref_harness.run_reference_predict(net, opt, batch,
                                                                          capture=() if big else CAPTURE)
        restated = synth.select_answers(probs, batch[1]["num_cnt"])
        assert restated == picks, "synth.select_answers differs from SDNetTrainer.predict: %s vs %s" % (restated, picks)
        meta = "cfg=%s ragged=%s bert_init=%s seed=%d torch=%s picks=SDNetTrainer.predict" % (
            cfg, ragged, init, seed, torch.__version__)
        if big:
            np.savez_compressed(os.path.join(out_dir, "model_%s.npz" % name), meta=np.asarray(meta),
                                probs=probs.numpy(), logits=logits.numpy(), picks=np.asarray(picks, np.int64))
            print(name, meta, "bytes", os.path.getsize(os.path.join(out_dir, "model_%s.npz" % name)), flush=True)
            del net
            continue
        bert_calls = cap["Bert"]  # q, ocr, od: each a list of 12 [N, W, 768]
        data = {
            "probs": probs.numpy(), "logits": logits.numpy(), "picks": np.asarray(picks, np.int64),
            "bert_q_l0": bert_calls[0][0][:2, :4].numpy(), "bert_q_l11": bert_calls[0][-1][:2, :4].numpy(),
            "bert_ocr_l0": bert_calls[1][0][:6, :2].numpy(), "bert_ocr_l11": bert_calls[1][-1][:6, :2].numpy(),
            "bert_od_l11": bert_calls[2][-1][:4, :2].numpy(),
            "multi2one_ocr": first(cap["multi2one"][0])[:6, :3].numpy(),
            "context_rnn_ocr_last": first(cap["context_rnn"][0])[:, :12, :16].numpy(),
            "ques_rnn_last": first(cap["ques_rnn"][0])[:, :8, :16].numpy(),
            "deep_attn_ocr_after": first(cap["deep_attn"][0])[:, :12, :16].numpy(),
            "high_lvl_context_ocr": first(cap["high_lvl_context_rnn"][0])[:, :12, :16].numpy(),
            "ques_self_attn": first(cap["ques_self_attn"][0])[:, :8, :16].numpy(),
        }
        np.savez_compressed(os.path.join(out_dir, "model_%s.npz" % name), meta=np.asarray(meta), **data)
        print(name, meta, "picks", picks, "bytes", os.path.getsize(os.path.join(out_dir, "model_%s.npz" % name)), flush=True)


if __name__ == "__main__":
    import sys
    main(only=tuple(sys.argv[1:]))
