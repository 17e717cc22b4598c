"""CPU: the PHOC oracle (oracle/phoc_oracle.c) against the golden vectors generated from the
reference's own cphoc.c (tests/golden/phoc_known.json, made by oracle/gen_phoc_golden.py)."""
import json
import os

import numpy as np

from oracle import phoc_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "phoc_known.json")


def _gold():
    with open(GOLD) as f:
        return json.load(f)


def test_oracle_matches_reference_known_answers():
    g = _gold()
    assert g["random_check"]["mismatches"] == 0 and g["random_check"]["strings"] >= 200000
    words = sorted(g["known"])
    out, bad = phoc_oracle.batch(words)
    assert bad == -1
    for i, w in enumerate(words):
        assert [int(k) for k in np.nonzero(out[i])[0]] == g["known"][w], w
        assert set(np.unique(out[i])) <= {0.0, 1.0}


def test_survey_known_answers():
    out, _ = phoc_oracle.batch(["hello", "the", ""])
    assert int(out[0].sum()) == 21
    assert list(np.nonzero(out[1])[0]) == [19, 40, 43, 91, 115, 148, 199, 259, 292, 343, 403, 472, 504, 555]
    assert out[2].sum() == 0


def test_oracle_rejects_unknown_unigrams():
    g = _gold()
    for w, msg in g["errors"].items():
        assert msg is not None and msg.startswith("Error: unigram")
    out, bad = phoc_oracle.batch(["abc", "A", "x-y"])
    assert bad == 1
    assert out[1].sum() == 0 and out[2].sum() == 0 and out[0].sum() > 0


def test_oracle_against_ref_build_when_present():
    ref = phoc_oracle.ref_module()
    if ref is None:
        import pytest
        pytest.skip("oracle/_ref not built (reference tree absent)")
    import random
    rng = random.Random(7)
    alpha = "abcdefghijklmnopqrstuvwxyz0123456789"
    words = ["".join(rng.choice(alpha) for _ in range(rng.randint(0, 40))) for _ in range(3000)]
    out, _ = phoc_oracle.batch(words)
    for i, w in enumerate(words):
        assert np.array_equal(np.asarray(ref.build_phoc(w), dtype=np.float32), out[i])
