"""Seeded synthetic ST-VQA-shaped workloads: options, weights and batches.

Shared by the product's bench/smoke, the tests and the oracle so that all of them see the same
inputs.  Nothing here reads the reference tree; `DEFAULT_OPT` restates the hyper-parameters of
the reference's shipped `conf` file (the result of Utils/Arguments.py:41-66 on it) as data.

Batches have exactly the layout `VQA_collate.VQA_collate_fun` emits (Utils/VQA_Dataset.py:448-542):
three dicts of padded int64/bool/fp32 tensors plus the Python lists `bert_offsets`, `num_cnt`,
`len_cnt`.  Pad id is 0 everywhere; every image ends with the `<OCR>` / `<OD>` end item
(Utils/VQA_Dataset.py:336-349).
"""
import zlib

import numpy as np
import torch

# |POS| and |ENT| come from spaCy 2.0.18 en_core_web_sm in the reference (Utils/CoQAUtils.py:31-32)
# and cannot be known here; the harness fixes them.
POS_SIZE = 50
ENT_SIZE = 74
VOCAB_SIZE = 5000
BERT_VOCAB = 30522

DEFAULT_OPT = {
    "RESUME": True, "MODEL_PATH": "conf~/model/ANLS_best_model.pt", "Task": "test",
    "score_name": "ANLS", "mask_score": True, "label_no_answer": True,
    "max_ocr_num": 100, "max_od_num": 30, "max_ocr_len": 20, "max_od_len": 10,
    "max_ocr_bert_len": 30, "max_od_bert_len": 10, "max_q_len": 40, "max_q_bert_len": 50,
    "GLOVE": True, "glove_dim": 300, "FastText": True, "fast_dim": 300,
    "q_embedding": "glove,pos,ent,bert", "ocr_embedding": "fasttext,pos,ent,bert",
    "q_emb_initial": "glove", "ocr_emb_initial": "fasttext",
    "loss": "BCE_D1", "optimizer": "#", "batch_size": 16, "lr": 0.001, "num_worker": 0,
    "LN": True, "DROPOUT": 0.3, "VARIATIONAL_DROPOUT": True,
    "BERT": True, "dropout_emb": 0.4, "LOCK_BERT": True, "BERT_LINEAR_COMBINE": True,
    "SEED": 1033, "QUES_SELF_ATTN": True, "concat_rnn": False, "grad_clipping": 10,
    "do_seq_dropout": True, "TUNE_PARTIAL": True, "tune_partial": 1000, "embedding_dim": 300,
    "prealign_hidden": 300, "PRE_ALIGN": True, "PRE_ALIGN_befor_rnn": True, "pos_dim": 12,
    "ent_dim": 8, "flow_hidden_size": 300, "query_self_attn_hidden_size": 300, "hidden_size": 125,
    "deep_att_hidden_size_per_abstr": 250, "deep_inter_att_use_CoVe": 1, "in_rnn_layers": 2,
    "highlvl_hidden_size": 125, "question_high_lvl_rnn_layers": 1,
    "multi2one_hidden_size": 300, "multi2one_bidir": False,
    "position_dim": 8, "position_mod": "qk+", "pos_att_merge_mod": "cat",
    "useES": True, "ES_ocr": "ES_ocr", "ES_ocr_len": 10, "ES_sort_way": "frequency",
    "ES_using_way": "as_ocr",
}

# name -> (B, n_ocr (without the end item), n_od, max_ocr_num, max_od_num, q bert len)
CONFIGS = {
    "tiny": dict(B=3, n_ocr=12, n_od=4, max_ocr_num=100, max_od_num=30),
    "small": dict(B=8, n_ocr=20, n_od=8, max_ocr_num=100, max_od_num=30),
    "cfg1": dict(B=32, n_ocr=50, n_od=10, max_ocr_num=100, max_od_num=30),
    "cfg3": dict(B=256, n_ocr=50, n_od=36, max_ocr_num=100, max_od_num=37),
    "cfg4": dict(B=4096, n_ocr=50, n_od=36, max_ocr_num=100, max_od_num=37),
    "cfg5": dict(B=32, n_ocr=200, n_od=36, max_ocr_num=201, max_od_num=37, max_q_bert_len=512),
    "cfg5s": dict(B=4, n_ocr=200, n_od=36, max_ocr_num=201, max_od_num=37, max_q_bert_len=512),  # cfg5 shape, 4 images
}


def make_opt(cfg="cfg1", **overrides):
    """`opt` dict for a named config: the shipped conf + the harness keys of SURVEY.md §8c."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else dict(cfg)
    opt = dict(DEFAULT_OPT)
    opt["vocab_size"] = VOCAB_SIZE
    opt["cuda"] = torch.cuda.is_available()
    opt["datadir"] = ""
    opt["BERT_model_file"] = ""
    opt["pos_size"] = POS_SIZE
    opt["ent_size"] = ENT_SIZE
    for k in ("max_ocr_num", "max_od_num", "max_q_bert_len"):
        if k in c:
            opt[k] = c[k]
    opt.update(overrides)
    return opt


def make_embedding(seed=1033, vocab=VOCAB_SIZE, dim=300):
    """{'glove_embedding','fast_embedding'}: N(0,1) with the <PAD> row zeroed (CoQAUtils.py:37,55)."""
    out = {}
    for i, name in enumerate(("glove_embedding", "fast_embedding")):
        g = torch.Generator().manual_seed(seed * 7 + i)
        e = torch.randn(vocab, dim, generator=g)
        e[0] = 0
        out[name] = e
    return out


def make_vocab_words(seed=1033, vocab=VOCAB_SIZE):
    """Synthetic vocabulary strings: the five specials (CoQAPreprocess.py:531-535) then random
    tokens over [a-z0-9] with some upper case / punctuation for the PHOC wrapper to normalise."""
    rng = np.random.default_rng(seed * 13 + 5)
    alphabet = np.array(list("abcdefghijklmnopqrstuvwxyz0123456789"))
    words = ["<PAD>", "<UNK>", "<Q>", "<OCR>", "<OD>"]
    while len(words) < vocab:
        w = "".join(rng.choice(alphabet, size=int(rng.integers(1, 15))))
        r = rng.random()
        if r < 0.1:
            w = w.upper()
        elif r < 0.2:
            w = w[:1] + "-" + w[1:] + "."
        words.append(w)
    return words


PHOC_OPT = {"PHOC": True, "phoc_dim": 604, "ocr_embedding": "phoc,fasttext,pos,ent,bert"}


def add_phoc(batch, vocab_words=None, q_side=False):
    """Add the PHOC channel inputs to a batch: `phoc` = the word ids (VQA_Dataset.py:359-360).
    With `vocab_words` also the table-free form: `phoc_chars` / `phoc_offsets` holding the string
    of every word slot (see ruart_b200.Utils.phoc.encode_tokens)."""
    from .Utils.phoc import encode_tokens
    q, ocr, od = batch
    for d, key in ((q, "glove"), (ocr, "fasttext"), (od, "fasttext")):
        if d is q and not q_side:
            continue
        d["phoc"] = d[key].clone()
        if vocab_words is not None:
            chars, offsets = encode_tokens([vocab_words[i] for i in d[key].reshape(-1).tolist()])
            d["phoc_chars"] = torch.from_numpy(chars)
            d["phoc_offsets"] = torch.from_numpy(offsets)
    return batch


def _seed_for(name, seed):
    return (zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF


@torch.no_grad()
def fill_state_dict(module, seed=1033, bert_init="random"):
    """Deterministically overwrite every parameter/buffer of `module` from (name, seed).

    Works on the reference's SDNet and on ours alike (same state_dict names -> same values).
    bert_init: "random" keeps init_bert_weights' LayerNorm gamma,beta ~ N(0,0.02)
    (modeling.py:439-441); "pretrained_like" uses gamma=1, beta=0 so that BERT activations have a
    realistic scale (SURVEY.md hard parts).
    """
    sd = module.state_dict()
    for name, t in sd.items():
        if not torch.is_floating_point(t):
            continue
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        shape = tuple(t.shape)
        leaf = name.split(".")[-1]
        if name.startswith("Bert.") or name.startswith("bert_model."):
            if leaf == "gamma":
                v = torch.randn(shape, generator=g) * 0.02 if bert_init == "random" else torch.ones(shape)
            elif leaf == "beta":
                v = torch.randn(shape, generator=g) * 0.02 if bert_init == "random" else torch.zeros(shape)
            elif leaf == "bias":
                v = torch.randn(shape, generator=g) * 0.02
            else:
                v = torch.randn(shape, generator=g) * (0.02 if bert_init == "random" else 0.04)
        elif name == "phoc_embed.weight":
            continue  # the PHOC table is data (CoQAUtils.py:75-87), not a random weight
        elif name in ("glove_embed.weight", "fast_embed.weight"):
            v = torch.randn(shape, generator=g)
            v[0] = 0
        elif name == "alphaBERT":
            v = 1.0 + 0.5 * torch.randn(shape, generator=g)
        elif name == "gammaBERT":
            v = torch.full(shape, 0.9)
        elif leaf == "diagonal":
            if t.numel() == 1:
                v = t.clone().float()  # constant 1/sqrt(hidden) (Layers.py:198), not trainable
            else:
                v = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif "embedding" in name:
            v = torch.randn(shape, generator=g)
        elif leaf.startswith("weight_ih") or leaf.startswith("weight_hh") or leaf.startswith("bias_ih") \
                or leaf.startswith("bias_hh"):
            hidden = shape[0] // (3 if name.startswith("get_answer.rnn") else 4)
            k = 1.0 / (hidden ** 0.5)
            v = (torch.rand(shape, generator=g) * 2 - 1) * k
        elif leaf == "weight":
            k = 1.0 / (shape[-1] ** 0.5)
            v = (torch.rand(shape, generator=g) * 2 - 1) * k
        elif leaf == "bias":
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        else:
            v = torch.randn(shape, generator=g) * 0.1
        t.copy_(v.to(t.dtype))
    return module


def _bertify(piece_counts, rng, first_piece=None):
    """[CLS] pieces... [SEP] with per-word [st, ed) offsets (VQA_Dataset.py:415-436)."""
    ids = [101]
    offs = []
    for w, n in enumerate(piece_counts):
        offs.append([len(ids), len(ids) + n])
        if first_piece is not None and w == 0:
            ids.extend([first_piece] * n)
        else:
            ids.extend(int(x) for x in rng.integers(1000, 30000, size=n))
    ids.append(102)
    return ids, offs


def _items(B, n_items_per_image, end_word_id, W, L, M, max_words, max_pieces, rng):
    """One of ocr_list / od_list."""
    num_cnt, len_cnt, offsets = [], [], []
    rows_word, rows_pos, rows_ent, rows_bert = [], [], [], []
    position = torch.zeros(B, M, 8)
    for b in range(B):
        n = int(n_items_per_image[b])
        assert n + 1 <= M
        num_cnt.append(n + 1)
        lens = []
        for it in range(n + 1):
            end = it == n
            nw = 1 if end else int(rng.integers(1, max_words + 1))
            pieces = [1] if end else [int(x) for x in rng.integers(1, max_pieces + 1, size=nw)]
            ids, offs = _bertify(pieces, rng, first_piece=1000 + end_word_id if end else None)
            assert len(ids) <= L and nw <= W
            word = np.zeros(W, np.int64)
            pos = np.zeros(W, np.int64)
            ent = np.zeros(W, np.int64)
            if end:
                word[0] = end_word_id
            else:
                word[:nw] = rng.integers(5, VOCAB_SIZE, size=nw)
                pos[:nw] = rng.integers(1, POS_SIZE, size=nw)
                ent[:nw] = rng.integers(1, ENT_SIZE, size=nw)
                position[b, it] = torch.from_numpy(rng.random(8).astype(np.float32))
            bert = np.zeros(L, np.int64)
            bert[:len(ids)] = ids
            rows_word.append(word)
            rows_pos.append(pos)
            rows_ent.append(ent)
            rows_bert.append(bert)
            offsets.append(offs)
            lens.append(nw)
        len_cnt.append(lens)
    res = {
        "fasttext": torch.from_numpy(np.stack(rows_word)),
        "pos": torch.from_numpy(np.stack(rows_pos)),
        "ent": torch.from_numpy(np.stack(rows_ent)),
        "bert": torch.from_numpy(np.stack(rows_bert)),
        "bert_offsets": offsets,
        "position": position,
        "num_cnt": num_cnt,
        "len_cnt": len_cnt,
    }
    res["fasttext_mask"] = ~res["fasttext"].eq(0)
    res["bert_mask"] = ~res["bert"].eq(0)
    return res


def make_batch(cfg="cfg1", seed=None, ragged=False, opt=None, n_q_words=20):
    """(q_list, ocr_list, od_list) on the CPU, exactly as VQA_collate_fun would hand them over."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else dict(cfg)
    if opt is None:
        opt = make_opt(c)
    if seed is None:
        seed = 2000 + (list(CONFIGS).index(cfg) if isinstance(cfg, str) else 99)
    rng = np.random.default_rng(seed)
    B = c["B"]
    Lq, Wq = opt["max_q_bert_len"], opt["max_q_len"]
    glove = np.zeros((B, Wq), np.int64)
    pos = np.zeros((B, Wq), np.int64)
    ent = np.zeros((B, Wq), np.int64)
    bert = np.zeros((B, Lq), np.int64)
    q_offsets = []
    for b in range(B):
        nw = n_q_words if not ragged else int(rng.integers(3, n_q_words + 1))
        glove[b, :nw] = rng.integers(5, VOCAB_SIZE, size=nw)
        pos[b, :nw] = rng.integers(1, POS_SIZE, size=nw)
        ent[b, :nw] = rng.integers(1, ENT_SIZE, size=nw)
        ids, offs = _bertify([int(x) for x in rng.integers(1, 3, size=nw)], rng)
        assert len(ids) <= Lq
        bert[b, :len(ids)] = ids
        q_offsets.append(offs)
    q_list = {"glove": torch.from_numpy(glove), "pos": torch.from_numpy(pos), "ent": torch.from_numpy(ent),
              "bert": torch.from_numpy(bert), "bert_offsets": q_offsets}
    q_list["glove_mask"] = ~q_list["glove"].eq(0)
    q_list["bert_mask"] = ~q_list["bert"].eq(0)
    if ragged:
        n_ocr = rng.integers(min(5, c["n_ocr"]), c["n_ocr"] + 1, size=B)
        n_od = rng.integers(min(1, c["n_od"]), c["n_od"] + 1, size=B)
    else:
        n_ocr = np.full(B, c["n_ocr"])
        n_od = np.full(B, c["n_od"])
    ocr_list = _items(B, n_ocr, 3, opt["max_ocr_len"], opt["max_ocr_bert_len"], opt["max_ocr_num"], 2, 3, rng)
    od_list = _items(B, n_od, 4, opt["max_od_len"], opt["max_od_bert_len"], opt["max_od_num"], 2, 2, rng)
    return q_list, ocr_list, od_list


def uncollate(batch):
    """Per-sample dicts as VQA_Dataset.__getitem__ hands them to the collate function
    (Utils/VQA_Dataset.py:448-462: keys 'q', 'ocr', 'od', 'gt', 'extra_info'), recovered from a
    collated synth batch — the input side of Utils/collate.py's tests."""
    q, ocr, od = batch
    B = len(ocr["num_cnt"])
    samples = []
    row = {"ocr": 0, "od": 0}
    for b in range(B):
        nq = int(q["glove_mask"][b].sum())
        nb = int(q["bert_mask"][b].sum())
        qs = {"glove": q["glove"][b, :nq].tolist(), "pos": q["pos"][b, :nq].tolist(), "ent": q["ent"][b, :nq].tolist(),
              "bert": q["bert"][b, :nb].tolist(), "bert_offsets": q["bert_offsets"][b]}
        lists = {}
        for name, d in (("ocr", ocr), ("od", od)):
            items = []
            for k in range(d["num_cnt"][b]):
                r = row[name]
                nw = d["len_cnt"][b][k]
                nt = int(d["bert_mask"][r].sum())
                items.append({"fasttext": d["fasttext"][r, :nw].tolist(), "pos": d["pos"][r, :nw].tolist(),
                              "ent": d["ent"][r, :nw].tolist(), "bert": d["bert"][r, :nt].tolist(),
                              "bert_offsets": d["bert_offsets"][r], "position": d["position"][b, k].tolist()})
                row[name] += 1
            lists[name] = items
        gt = torch.zeros(1, ocr["position"].size(1) + 1)
        gt[0, b % max(1, ocr["num_cnt"][b] - 1)] = 1.0
        samples.append({"q": qs, "ocr": lists["ocr"], "od": lists["od"], "gt": gt, "extra_info": {"q_id": b}})
    return samples


def batch_to(batch, device):
    """ToCUDA (SDNetTrainer.py:208-230): tensors to `device`, lists stay on the host."""
    out = []
    for d in batch:
        out.append({k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()})
    return tuple(out)


def shard_batch(batch, rank, world):
    """Contiguous split of the B questions (and their item rows) across `world` ranks."""
    q, ocr, od = batch
    B = len(ocr["num_cnt"])
    per = (B + world - 1) // world
    lo, hi = min(rank * per, B), min((rank + 1) * per, B)
    if hi <= lo:
        raise ValueError("rank %d of %d gets an empty shard of a %d-question batch" % (rank, world, B))
    # keys derived from the WHOLE batch by Utils.collate.attach_index_tensors (forward plan, CSR word
    # offsets, token counts) cannot be sliced: drop them — run attach_index_tensors again per shard
    derived = ("ruart_plan", "bert_offsets_csr", "bert_totals")

    def cut_items(d):
        if "phoc_chars" in d:
            raise ValueError("shard first, then add_phoc(..., vocab_words): the string buffers are per shard")
        start = sum(d["num_cnt"][:lo])
        stop = start + sum(d["num_cnt"][lo:hi])
        r = {}
        for k, v in d.items():
            if k in derived:
                continue
            if k in ("num_cnt", "len_cnt"):
                r[k] = v[lo:hi]
            elif k == "position":
                r[k] = v[lo:hi]
            elif k == "bert_offsets":
                r[k] = v[start:stop]
            else:
                r[k] = v[start:stop]
        return r

    qs = {k: v[lo:hi] for k, v in q.items() if k not in derived}
    return qs, cut_items(ocr), cut_items(od)


def select_answers(prob, num_cnt, label_no_answer=True):
    """Index rule of SDNetTrainer.predict (SDNetTrainer.py:402-412) on CPU probabilities.

    Walk the slots by descending probability; stop at the no-answer column (last), skip the
    `<OCR>` end slot (index num_cnt-1), accept the first index < num_cnt.
    """
    prob = prob.detach().float().cpu()
    res = []
    for i in range(prob.size(0)):
        _, ids = torch.sort(prob[i, :], descending=True)
        pick = int(ids[-1])
        for idx in ids.tolist():
            pick = idx
            if label_no_answer and idx == ids.numel() - 1:
                break
            if idx == num_cnt[i] - 1:
                continue
            if idx < num_cnt[i]:
                break
        res.append(pick)
    return res
