"""TEST INFRASTRUCTURE — pins the PHOC oracle and writes tests/golden/phoc_known.json.

Run in the build container (needs oracle/_ref/, i.e. /root/reference at build time):
    python -m oracle.gen_phoc_golden
1. compares oracle/phoc_oracle.c with the reference's own cphoc.c (oracle/_ref) on 200 000 seeded
   random strings (len 1-20 over a-z0-9, plus len 21-70) — must be 0 mismatches;
2. records the reference's outputs (indices of the ones) for a fixed list of words.
"""
import json
import os
import random

import numpy as np

from . import phoc_oracle

ALPHABET = "abcdefghijklmnopqrstuvwxyz0123456789"
KNOWN = ["", "a", "z", "0", "9", "ab", "th", "he", "the", "hello", "world", "stvqa", "b200",
         "thethethe", "aaaaaaaaaaaaaaaaaaaa", "0123456789", "abcdefghijklmnopqrstuvwxyz",
         "international", "supercalifragilisticexpialidocious", "ti", "in", "el", "ll", "x" * 33,
         "erererererererer", "question", "answer", "coca", "cola", "stop", "exit", "1984", "a1b2c3"]


def main():
    ref = phoc_oracle.ref_module()
    if ref is None:
        raise SystemExit("oracle/_ref/cphoc*.so missing: run `make -C oracle` with /root/reference present")
    rng = random.Random(2002)
    strings = ["".join(rng.choice(ALPHABET) for _ in range(rng.randint(1, 20))) for _ in range(190000)]
    strings += ["".join(rng.choice(ALPHABET) for _ in range(rng.randint(21, 70))) for _ in range(10000)]
    got, bad = phoc_oracle.batch(strings)
    assert bad == -1
    mism = 0
    for i, s in enumerate(strings):
        want = np.asarray(ref.build_phoc(s), dtype=np.float32)
        if not np.array_equal(want, got[i]):
            mism += 1
    print("oracle vs reference cphoc.c: %d mismatches / %d strings" % (mism, len(strings)))
    assert mism == 0
    known = {}
    for w in KNOWN:
        v = np.asarray(ref.build_phoc(w), dtype=np.float32)
        assert set(np.unique(v)) <= {0.0, 1.0}
        known[w] = [int(i) for i in np.nonzero(v)[0]]
    errors = {}
    for w in ["A", "a b", "x-y", "Hello", "ok!"]:
        try:
            ref.build_phoc(w)
            errors[w] = None
        except RuntimeError as e:
            errors[w] = str(e)
    out = {"source": "reference Utils/cphoc.c compiled by oracle/Makefile (gcc, -O2)",
           "random_check": {"strings": len(strings), "mismatches": mism, "seed": 2002},
           "known": known, "errors": errors}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "phoc_known.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
