"""CPU: property tests (hypothesis) of the host-side index logic and the tokenizer drop-in."""
import os

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from ruart_b200 import host_index
from ruart_b200.bert_engine import flatten_offsets
from ruart_b200.Utils import tokenization as T

from helpers import GOLDEN

images = st.lists(st.lists(st.integers(1, 6), min_size=1, max_size=9), min_size=1, max_size=6)


@settings(max_examples=60, deadline=None)
@given(ocr=images, od_seed=st.integers(0, 1000))
def test_forward_plan_invariants(ocr, od_seed):
    rng = np.random.default_rng(od_seed)
    od = [[int(x) for x in rng.integers(1, 4, size=int(rng.integers(1, 5)))] for _ in ocr]
    Wo, Wd, M, M_od = 6, 3, 9, 4
    plan = host_index.forward_plan([len(i) for i in ocr], ocr, [len(i) for i in od], od, Wo, Wd, M, M_od)
    lens = np.array([l for img in ocr for l in img] + [l for img in od for l in img])
    B, N_ocr = len(ocr), sum(len(i) for i in ocr)
    c = plan["cuts"]
    ocr_src, ocr_dst, od_src, od_dst, steps, last = [plan["i32"][c[i]:c[i + 1]] for i in range(6)]
    # pre-align pack: a bijection between real word slots and packed positions, image by image
    assert len(set(ocr_dst.tolist())) == len(ocr_dst) == int(lens[:N_ocr].sum())
    assert (ocr_dst // plan["T_max"][0] == np.repeat(np.arange(B), [sum(i) for i in ocr])).all()
    assert (ocr_src % Wo < Wo).all() and (od_src % Wd < Wd).all()
    # multi2one schedule: every real word step exactly once, grouped by step, longest items first
    assert len(steps) == int(lens.sum()) and len(set(steps.tolist())) == len(steps)
    assert plan["n_t"] == [int((lens > t).sum()) for t in range(int(lens.max()))]
    assert np.array_equal(np.sort(last), np.sort(lens - 1)) and (np.diff(last) <= 0).all()
    # every item lands in its own slot; masks count the items
    assert len(set(plan["slots"].tolist())) == len(lens)
    masks = plan["masks"]
    assert masks[:B * M].reshape(B, M).sum(1).tolist() == [len(i) for i in ocr]
    assert masks[B * M:].reshape(B, M_od).sum(1).tolist() == [len(i) for i in od]


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.tuples(st.integers(0, 20), st.integers(0, 20)), max_size=5), min_size=1, max_size=8))
def test_flatten_offsets_round_trip(items):
    offs = [[list(p) for p in it] for it in items]
    w = flatten_offsets(offs, len(offs))
    assert w.shape == (4, sum(len(i) for i in offs)) and w.dtype == np.int32
    back = [[] for _ in offs]
    for r, j, s, e in w.T.tolist():
        assert j == len(back[r])
        back[r].append([s, e])
    assert back == offs
    assert np.array_equal(flatten_offsets(w, len(offs)), w)


_TOK = T.BertTokenizer(os.path.join(GOLDEN, "tokenizer_vocab.txt"))


@settings(max_examples=200, deadline=None)
@given(st.text(alphabet="abcxyz019 .,!-ABC", max_size=30))
def test_tokenizer_properties_on_ascii(text):
    toks = _TOK.tokenize(text)
    vocab = _TOK.vocab
    assert all(t in vocab for t in toks)                      # only vocabulary entries come out
    # the pieces spell the lower-cased text without its whitespace (this vocabulary covers every
    # ASCII letter, digit and the punctuation used here, so nothing maps to [UNK])
    spelled = "".join(t[2:] if t.startswith("##") else t for t in toks)
    assert spelled == "".join(text.lower().split())
    # tokenizing word by word gives the same pieces (what bertify relies on)
    assert [t for w in text.split() for t in _TOK.tokenize(w)] == toks
    ids, offs = _TOK.bertify(text.split())
    assert ids[0] == vocab["[CLS]"] and ids[-1] == vocab["[SEP]"]
    if text.split():
        assert offs[0][0] == 1 and offs[-1][1] == len(ids) - 1 and all(a <= b for a, b in offs)
