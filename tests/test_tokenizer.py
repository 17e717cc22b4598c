"""CPU: the WordPiece tokenizer / bertify drop-in (ruart_b200/Utils/tokenization.py) against
golden vectors produced by the UNMODIFIED reference (oracle/gen_tokenizer_golden.py)."""
import json
import os

import numpy as np

from ruart_b200.bert_engine import flatten_offsets
from ruart_b200.Utils import tokenization as T

from helpers import GOLDEN


def _load():
    with open(os.path.join(GOLDEN, "tokenizer_golden.json"), encoding="utf8") as f:
        g = json.load(f)
    return g, T.BertTokenizer(os.path.join(GOLDEN, "tokenizer_vocab.txt"))


def test_tokenize_matches_reference_on_every_golden_string():
    g, tok = _load()
    assert len(g["strings"]) >= 400
    for s, want in zip(g["strings"], g["tokens"]):
        assert tok.tokenize(s) == want, repr(s)
    # known behaviours: greedy longest match, lower-casing + accent stripping, > 100 characters -> [UNK]
    assert tok.tokenize("unaffable") == ["un", "##aff", "##able"]
    assert tok.tokenize("Café") == ["cafe"]
    assert tok.tokenize("a" * 101) == ["[UNK]"]
    ids = tok.convert_tokens_to_ids(["[CLS]", "stop", "[SEP]"])
    assert tok.convert_ids_to_tokens(ids) == ["[CLS]", "stop", "[SEP]"]
    assert T.BertTokenizer.from_pretrained("/nonexistent/vocab.txt") is None


def test_bertify_matches_reference_and_feeds_the_collate_layout():
    g, tok = _load()
    for words, (ids, offs) in zip(g["items"], g["bertify"]):
        got_ids, got_offs = tok.bertify(list(words))
        assert got_ids == ids and got_offs == offs
    for s, (ids, offs) in zip(g["strings"][2:12], g["bertify_str"]):
        assert tok.bertify(s) == (ids, offs)
    batch = tok.bertify_batch(g["items"])
    N = len(g["items"])
    assert batch["bert"].shape[0] == N and batch["bert"].dtype == np.int64
    for r, (ids, offs) in enumerate(g["bertify"]):
        assert batch["bert"][r, :len(ids)].tolist() == ids and not batch["bert"][r, len(ids):].any()
        assert batch["bert_offsets"][r] == offs
    assert (batch["bert_mask"] == (batch["bert"] != 0)).all()
    # the CSR array is what the engine derives from the nested lists (incl. the flat [1, 1] of an empty item)
    assert np.array_equal(batch["bert_offsets_csr"], flatten_offsets(batch["bert_offsets"], N))
    padded = tok.bertify_batch(g["items"], max_len=batch["bert"].shape[1] + 5)
    assert padded["bert"].shape[1] == batch["bert"].shape[1] + 5
