"""Drop-in for the reference's Utils/phoc.py:8-13 (`build_phoc(token) -> list[float] * 604`),
backed by the sm_100a PHOC kernel instead of the cphoc CPython extension.

`build_phoc_batch` is the form the GPU is meant for (one launch for many tokens); the single
token form exists for API compatibility with SDNetTrainer.load_fixed_answers
(SDNetTrainer.py:273-274).  `build_phoc_embedding` replaces CoQAUtils.build_phoc_embedding
(CoQAUtils.py:75-87), which builds the `[vocab, 604]` table behind `SDNet.phoc_embed`
(SDNet.py:51-55); `encode_tokens` prepares the strings of a batch for the table-free PHOC channel
(`item_list['phoc_chars']` / `['phoc_offsets']`, see ruart_b200/Models/SDNet.py).
"""
import numpy as np

from .. import ops

_alphabet = set("abcdefghijklmnopqrstuvwxyz0123456789")


def _normalise(token):
    token = token.lower().strip()
    return "".join(c for c in token if c in _alphabet)


def encode_tokens(tokens):
    """list[str] -> (chars uint8 [total+1], offsets int32 [n+1]) after the wrapper's normalisation
    (lower, strip, drop characters outside [a-z0-9]: Utils/phoc.py:9-10)."""
    enc = [_normalise(t).encode("ascii") for t in tokens]
    offsets = np.zeros(len(enc) + 1, dtype=np.int32)
    if enc:
        offsets[1:] = np.cumsum([len(e) for e in enc])
    chars = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
    return chars, offsets


def build_phoc_batch(tokens, device="cuda"):
    """list[str] -> float32 tensor [n, 604] on `device`."""
    return ops.phoc_strings([_normalise(t) for t in tokens], device=device)


def build_phoc(token):
    return build_phoc_batch([token])[0].cpu().tolist()


def build_phoc_embedding(targ_vocab, wv_dim=604, device="cuda"):
    """[len(vocab), 604] float32 numpy table, row i = PHOC of vocabulary word i
    (CoQAUtils.py:75-87: every row, including <PAD>, is overwritten by the word's PHOC)."""
    if wv_dim != ops.PHOC_DIM:
        raise ValueError("PHOC vectors have %d entries" % ops.PHOC_DIM)
    return build_phoc_batch(list(targ_vocab), device=device).cpu().numpy()
