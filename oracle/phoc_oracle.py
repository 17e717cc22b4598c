"""TEST INFRASTRUCTURE — Python access to the PHOC CPU oracles.

`batch(strings)`  : our plain-C restatement (oracle/phoc_oracle.c, follows Utils/cphoc.c:12-113)
`ref_build_phoc`  : the reference's own cphoc.c compiled into oracle/_ref/ (None when absent)
"""
import ctypes
import glob
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libphoc_oracle.so")
PHOC_DIM = 604
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True, capture_output=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.phoc_oracle_batch.restype = ctypes.c_longlong
        _lib.phoc_oracle_batch.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong,
                                           ctypes.c_void_p]
    return _lib


def flatten(strings):
    """list[str] -> (uint8 char buffer, int32 offsets[n+1]) — the C ABI's batch input layout."""
    enc = [s.encode("latin-1") for s in strings]
    offsets = np.zeros(len(enc) + 1, dtype=np.int32)
    if enc:
        offsets[1:] = np.cumsum([len(e) for e in enc])
    chars = np.frombuffer(b"".join(enc), dtype=np.uint8).copy()
    return chars, offsets


def batch_flat(chars, offsets):
    n = len(offsets) - 1
    out = np.empty((n, PHOC_DIM), dtype=np.float32)
    buf = chars.tobytes() + b"\0"
    bad = _load().phoc_oracle_batch(buf, offsets.ctypes.data, n, out.ctypes.data)
    return out, int(bad)


def batch(strings):
    chars, offsets = flatten(strings)
    return batch_flat(chars, offsets)


def ref_module():
    """The reference's cphoc extension built by oracle/Makefile into oracle/_ref/, or None."""
    cands = glob.glob(os.path.join(_HERE, "_ref", "cphoc*.so"))
    if not cands:
        return None
    spec = importlib.util.spec_from_file_location("cphoc", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_ALPHABET = set("abcdefghijklmnopqrstuvwxyz0123456789")


def vocab_table(words, use_ref=True):
    """[len(words), 604] float32: what CoQAUtils.build_phoc_embedding (CoQAUtils.py:75-87) intends —
    row i = build_phoc(word i) with the wrapper's normalisation (Utils/phoc.py:8-13).  Uses the
    reference's own cphoc (oracle/_ref) when it was built, else the C restatement (bit-identical)."""
    norm = ["".join(c for c in w.lower().strip() if c in _ALPHABET) for w in words]
    ref = ref_module() if use_ref else None
    if ref is not None:
        return np.asarray([ref.build_phoc(w) for w in norm], dtype=np.float32)
    out, bad = batch(norm)
    assert bad < 0, "unknown unigram after normalisation"
    return out
