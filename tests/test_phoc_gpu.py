"""GPU: the sm_100a PHOC kernel must be bit-exact with the CPU oracle (which is pinned to the
reference's cphoc.c) — BASELINE config 2."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import phoc_oracle

pytestmark = pytest.mark.gpu
ALPHA = "abcdefghijklmnopqrstuvwxyz0123456789"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "phoc_known.json")


def _run(strings, packed=False):
    from ruart_b200 import ops
    chars, offsets = phoc_oracle.flatten(strings)
    chars = np.concatenate([chars, np.zeros(1, np.uint8)])
    d_c = torch.from_numpy(chars).cuda()
    d_o = torch.from_numpy(offsets).cuda()
    return ops.phoc_batch(d_c, d_o, packed=packed)


def test_known_answers_from_reference():
    with open(GOLD) as f:
        g = json.load(f)
    words = sorted(g["known"])
    out = _run(words).cpu().numpy()
    for i, w in enumerate(words):
        assert [int(k) for k in np.nonzero(out[i])[0]] == g["known"][w], w


@pytest.mark.parametrize("n,lo,hi,seed", [(200000, 1, 20, 2002), (20000, 0, 3, 1), (5000, 21, 200, 2),
                                          (33, 1, 20, 3), (1, 5, 5, 4)])
def test_bit_exact_vs_oracle(n, lo, hi, seed):
    rng = random.Random(seed)
    words = ["".join(rng.choice(ALPHA) for _ in range(rng.randint(lo, hi))) for _ in range(n)]
    want, bad = phoc_oracle.batch(words)
    assert bad == -1
    got = _run(words).cpu().numpy()
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_packed_output_matches_dense():
    rng = random.Random(11)
    words = ["".join(rng.choice(ALPHA) for _ in range(rng.randint(0, 25))) for _ in range(4097)]
    dense = _run(words).cpu().numpy()
    packed = _run(words, packed=True).cpu().numpy().view(np.uint32)
    bits = ((packed[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(len(words), -1)
    assert np.array_equal(bits[:, :604].astype(np.float32), dense)
    assert bits[:, 604:].sum() == 0


def test_empty_batch_and_empty_strings():
    out = _run([])
    assert out.shape == (0, 604)
    out = _run(["", "", "a"]).cpu().numpy()
    want, _ = phoc_oracle.batch(["", "", "a"])
    assert out[0].sum() == 0 and out[1].sum() == 0 and np.array_equal(out, want)


def test_unknown_unigram_raises_like_reference():
    with pytest.raises(RuntimeError, match="Error: unigram A is unknown"):
        _run(["abc", "A", "def"])
    with pytest.raises(RuntimeError, match="Error: unigram - is unknown"):
        _run(["x-y"])


def test_host_entry_point():
    import ctypes
    from ruart_b200 import _lib
    words = ["hello", "the", "", "b200"]
    chars, offsets = phoc_oracle.flatten(words)
    out = np.empty((len(words), 604), np.float32)
    bad_i = ctypes.c_int64(0)
    bad_c = ctypes.c_int32(0)
    rc = _lib.lib().ruart_phoc_batch_host(chars.tobytes() + b"\0", offsets.ctypes.data, len(words),
                                          out.ctypes.data, ctypes.byref(bad_i), ctypes.byref(bad_c))
    assert rc == 0 and bad_i.value == -1
    want, _ = phoc_oracle.batch(words)
    assert np.array_equal(out, want)


def test_unit_drop_in_build_phoc():
    from ruart_b200.Utils.phoc import build_phoc
    v = build_phoc(" Hello! ")
    want, _ = phoc_oracle.batch(["hello"])
    assert isinstance(v, list) and len(v) == 604 and isinstance(v[0], float)
    assert np.array_equal(np.asarray(v, np.float32), want[0])


def test_baseline_config2_one_million_strings_bit_exact():
    """BASELINE.json configs[1]: 1M synthetic OCR strings (len 1-20) vs Utils/cphoc.c (through the
    C oracle pinned to it), compared through a checksum of checksums and a dense sample."""
    from ruart_b200 import ops
    rng = np.random.default_rng(2002)
    n = 1_000_000
    lens = rng.integers(1, 21, size=n)
    offsets = np.zeros(n + 1, np.int32)
    offsets[1:] = np.cumsum(lens)
    alpha = np.frombuffer(ALPHA.encode(), np.uint8)
    chars = alpha[rng.integers(0, 36, size=int(offsets[-1]))]
    want, bad = phoc_oracle.batch_flat(chars, offsets)
    assert bad == -1
    got = ops.phoc_batch(torch.from_numpy(np.concatenate([chars, np.zeros(1, np.uint8)])).cuda(),
                         torch.from_numpy(offsets).cuda())
    assert got.shape == (n, 604)
    w = torch.arange(1, 605, device="cuda", dtype=torch.float64)
    row_sig = (got.double() * w).sum(1).cpu().numpy()          # per-string checksum
    want_sig = (want.astype(np.float64) * np.arange(1, 605)).sum(1)
    assert np.array_equal(row_sig, want_sig)
    idx = rng.integers(0, n, size=20000)
    assert np.array_equal(got[torch.from_numpy(idx).cuda()].cpu().numpy(), want[idx])
