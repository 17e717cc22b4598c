"""Drop-in for the reference's Utils/phoc.py:8-13 (`build_phoc(token) -> list[float] * 604`),
backed by the sm_100a PHOC kernel instead of the cphoc CPython extension.

`build_phoc_batch` is the form the GPU is meant for (one launch for many tokens); the single
token form exists for API compatibility with CoQAUtils.build_phoc_embedding (CoQAUtils.py:75-87)
and SDNetTrainer.load_fixed_answers (SDNetTrainer.py:273-274).
"""
from .. import ops

_alphabet = set("abcdefghijklmnopqrstuvwxyz0123456789")


def _normalise(token):
    token = token.lower().strip()
    return "".join(c for c in token if c in _alphabet)


def build_phoc_batch(tokens, device="cuda"):
    """list[str] -> float32 tensor [n, 604] on `device`."""
    return ops.phoc_strings([_normalise(t) for t in tokens], device=device)


def build_phoc(token):
    return build_phoc_batch([token])[0].cpu().tolist()
