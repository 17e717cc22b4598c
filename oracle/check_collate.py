"""TEST INFRASTRUCTURE — ruart_b200.Utils.collate.VQA_collate against the UNMODIFIED reference collate
(Utils/VQA_Dataset.py:438-542) on per-sample dicts recovered from seeded synth batches.  Build container
only (imports the reference tree):

    python -m oracle.check_collate

The committed CPU test (tests/test_host_logic.py) checks the same drop-in against the collated synth
batches, whose layout this script shows to be what the reference's collate emits.
"""
import sys
import time

import torch

from ruart_b200 import synth
from ruart_b200.Utils.collate import VQA_collate

from . import ref_harness


def main():
    if not ref_harness.available():
        raise SystemExit("needs the reference tree (build container only)")
    ref_harness._install_stubs()
    if ref_harness.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REF_ROOT)
    from Utils.VQA_Dataset import VQA_collate as RefCollate
    for cfg, ragged in (("tiny", True), ("small", False), ("cfg1", True)):
        opt = synth.make_opt(cfg)
        samples = synth.uncollate(synth.make_batch(cfg, ragged=ragged))
        t0 = time.perf_counter()
        want = RefCollate(opt).VQA_collate_fun(samples)
        t1 = time.perf_counter()
        got = VQA_collate(opt, index=False).VQA_collate_fun(samples)
        t2 = time.perf_counter()
        for g, w in zip(got[:3], want[:3]):
            assert set(g) == set(w), set(g) ^ set(w)
            for k, v in w.items():
                if torch.is_tensor(v):
                    assert g[k].dtype == v.dtype and g[k].shape == v.shape and torch.equal(g[k], v), (cfg, k)
                else:
                    assert g[k] == v, (cfg, k)
        assert torch.equal(got[3], want[3]) and got[4] == want[4]
        print("%s: identical to the reference collate (reference %.1f ms, ours %.1f ms)" %
              (cfg, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))


if __name__ == "__main__":
    main()
