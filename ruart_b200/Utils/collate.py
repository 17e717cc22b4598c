"""Collate-side helpers: turn the Python lists `VQA_collate_fun` emits
(Utils/VQA_Dataset.py:476-542: `bert_offsets`, `num_cnt`, `len_cnt`) into flat index arrays once,
in the DataLoader worker, and stage the batch in pinned memory so `ToCUDA`
(Models/SDNetTrainer.py:208-230) becomes asynchronous copies.

    batch = VQA_collate_fun(samples)                     # the reference's collate, unchanged
    q, ocr, od = attach_index_tensors(*batch)            # adds keys, keeps every original key
    q, ocr, od = to_cuda(pin((q, ocr, od)))              # pinned staging + non_blocking H2D
    scores, _ = network(q, ocr, od)

The added keys are optional: `SDNet.forward` rebuilds them from the lists when they are absent, so a
batch straight from the reference's collate works unchanged.
"""
import numpy as np
import torch

from ..bert_engine import flatten_offsets
from ..host_index import forward_plan

CSR_KEY = "bert_offsets_csr"   # int32 [4, n_words]: (row, word, st, ed) of every word of the list
PLAN_KEY = "ruart_plan"        # host_index.forward_plan(...) of the batch, stored in ocr_list
TOTALS_KEY = "bert_totals"     # (real wordpieces of the list, longest 512-token window): lets the token
                               # packing of the BERT pass skip its one device -> host read-back
WINDOW = 512                   # Bert.BERT_MAX_LEN (Bert.py:18)


def check_offsets(csr, mask):
    """The packed encoder keeps only the real wordpieces of a row, so (a) `bert_mask` must be a contiguous
    prefix of each row (what `~ids.eq(0)` gives for the reference's right-padded ids, VQA_Dataset.py:476-542)
    and (b) a word's [st, ed) must lie inside its row's real wordpieces.  The reference would read a pad
    position's hidden state for an offset past them (Bert.py:153-165); here that position does not exist, so it
    is an error instead of a silent read of the neighbouring row.  Degenerate spans (st >= ed) read nothing."""
    mask = mask.numpy() if torch.is_tensor(mask) else np.asarray(mask)
    if mask.size and (mask[:, 1:] & ~mask[:, :-1]).any():
        raise ValueError("bert_mask must be a contiguous prefix of every row (right-padded wordpiece ids)")
    if csr.shape[1] == 0:
        return
    lens = mask.sum(1)
    live = csr[2] < csr[3]
    bad = live & ((csr[2] < 0) | (csr[3] > lens[csr[0]]))
    if bad.any():
        k = int(np.flatnonzero(bad)[0])
        raise ValueError("bert_offsets: word %d of row %d spans wordpieces [%d, %d) but the row has %d"
                         % (csr[1, k], csr[0, k], csr[2, k], csr[3, k], lens[csr[0, k]]))


def attach_index_tensors(q_list, ocr_list, od_list):
    for d in (q_list, ocr_list, od_list):
        d[CSR_KEY] = flatten_offsets(d["bert_offsets"], len(d["bert_offsets"]))
        m = d["bert_mask"].cpu() != 0
        check_offsets(d[CSR_KEY], m)
        longest = max((int(m[:, p:p + WINDOW].sum(1).max()) for p in range(0, m.shape[1], WINDOW)), default=0) \
            if m.shape[0] else 0
        d[TOTALS_KEY] = (int(m.sum()), longest)
    ocr_list[PLAN_KEY] = forward_plan(
        ocr_list["num_cnt"], ocr_list["len_cnt"], od_list["num_cnt"], od_list["len_cnt"],
        ocr_list["fasttext"].size(1), od_list["fasttext"].size(1),
        ocr_list["position"].size(1), od_list["position"].size(1))
    return q_list, ocr_list, od_list


class VQA_collate(object):
    """Drop-in for the reference's collate class (Utils/VQA_Dataset.py:438-542): same constructor,
    same `VQA_collate_fun(batch) -> (q_list, ocr_list, od_list, gt, extra_info)` and the same dict
    layout (int64 id tensors padded with 0, `<key>_mask` = ~eq(0), fp32 `position` [B, max_num, 8],
    `bert_offsets` / `num_cnt` / `len_cnt` lists) — filled through numpy in one pass per key instead of
    a `torch.cat` per image, and with the index tensors of `attach_index_tensors` already attached
    (`index=False` gives the reference's keys only; `pinned=True` stages the tensors in pinned memory)."""

    ID_KEYS = ("glove", "fasttext", "phoc", "bert", "bert_only")

    def __init__(self, opt, index=True, pinned=False):
        self.opt, self.index, self.pinned = opt, index, pinned

    def VQA_collate_fun(self, batch):
        opt = self.opt
        q_list = self.que_collate([t["q"] for t in batch], opt["max_q_len"], opt["max_q_bert_len"])
        ocr_list = self.item_collate([t["ocr"] for t in batch], opt["max_ocr_len"], opt["max_ocr_bert_len"],
                                     opt["max_ocr_num"])
        od_list = self.item_collate([t["od"] for t in batch], opt["max_od_len"], opt["max_od_bert_len"],
                                    opt["max_od_num"])
        gt = torch.cat([t["gt"] for t in batch], dim=0)
        if self.index:
            attach_index_tensors(q_list, ocr_list, od_list)
        if self.pinned:
            q_list, ocr_list, od_list = pin((q_list, ocr_list, od_list))
        return q_list, ocr_list, od_list, gt, [t["extra_info"] for t in batch]

    @staticmethod
    def _pad_ids(rows, width):
        out = np.zeros((len(rows), width), dtype=np.int64)
        for i, r in enumerate(rows):
            out[i, :len(r)] = r          # a row longer than `width` raises, like the reference's slice assignment
        return torch.from_numpy(out)

    def _finish(self, res):
        for k in list(res):
            if k in self.ID_KEYS:
                res[k + "_mask"] = ~res[k].eq(0)
        return res

    def que_collate(self, q_list, max_len, max_bert_len):
        res = {}
        for k in q_list[0].keys():
            if k in ("img_features", "img_spatials"):
                res[k] = torch.cat([t[k] for t in q_list], dim=0)
            elif "offset" in k:
                res[k] = [t[k] for t in q_list]
            else:
                res[k] = self._pad_ids([t[k] for t in q_list], max_bert_len if k in ("bert", "bert_only") else max_len)
        return self._finish(res)

    def item_collate(self, item_list, max_len, max_bert_len, max_num):
        res = {}
        flat = [it for items in item_list for it in items]
        for k in item_list[0][0].keys():
            if "offset" in k:
                res[k] = [it[k] for it in flat]
            elif k == "position":
                pos = np.zeros((len(item_list), max_num, 8), dtype=np.float32)
                for b, items in enumerate(item_list):
                    if items:
                        pos[b, :len(items)] = np.asarray([it[k] for it in items], dtype=np.float32)
                res[k] = torch.from_numpy(pos)
            else:
                res[k] = self._pad_ids([it[k] for it in flat], max_bert_len if k in ("bert", "bert_only") else max_len)
        self._finish(res)
        res["num_cnt"] = [len(items) for items in item_list]
        word_key = "fasttext" if "FastText" in self.opt else "glove"
        res["len_cnt"] = [[len(it[word_key]) for it in items] for items in item_list]
        return res


def pin(batch):
    """Tensors of the three dicts into pinned host memory (lists / plans untouched)."""
    return tuple({k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in d.items()} for d in batch)


def to_cuda(batch, device="cuda"):
    """ToCUDA (SDNetTrainer.py:208-230) with non_blocking copies; host lists / plans stay on the host."""
    return tuple({k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()}
                 for d in batch)


class CudaPrefetcher(object):
    """`ToCUDA` (SDNetTrainer.py:208-230) one batch AHEAD: the host -> device copies of batch i + 1 run on a copy
    stream while batch i is on the compute stream ("pinned-memory async H2D in ToCUDA").

        for q, ocr, od in CudaPrefetcher(pinned_batches):      # pinned_batches: iterable of pin(...)-ed batches
            scores, _ = network(q, ocr, od)

    The device tensors live in TWO reusable sets of staging buffers (re-allocated only when a shape changes): set
    i % 2 is overwritten by batch i + 2 only after the compute stream has passed the point where batch i + 1 was
    requested, i.e. after everything that was queued for batch i.  (Fresh allocations on the copy stream +
    record_stream were measured first: the caching allocator cannot recycle those blocks while the host runs
    steps ahead of the GPU and the loop degenerated into cudaMalloc calls — 29 vs 24 ms per step.)
    The yielded dicts are new objects each time (SDNet.forward adds keys to them); the tensors inside are only
    valid until the batch after next is requested."""

    def __init__(self, batches=None, device="cuda"):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._sets = [{}, {}]
        self._free = [None, None]     # event on the compute stream after which set k may be overwritten
        self._n = 0
        self._next = None
        self.it = iter(())
        if batches is not None:
            self.feed(batches)

    def feed(self, batches):
        """Start over on another iterable of pinned batches, KEEPING the copy stream and the staging buffers (a
        data loader's next epoch; bench.py's warm-up and timed loops).  Returns self, so `for b in pf.feed(x)`."""
        if self._next is not None:
            raise RuntimeError("CudaPrefetcher.feed(): the previous iterable is not exhausted")
        self.it = iter(batches)
        self._preload()
        return self

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        k = self._n % 2
        self._n += 1
        bufs = self._sets[k]
        if self._free[k] is not None:
            self.stream.wait_event(self._free[k])
        out = []
        with torch.cuda.stream(self.stream):
            for di, d in enumerate(host):
                o = {}
                for key, v in d.items():
                    if torch.is_tensor(v):
                        b = bufs.get((di, key))
                        if b is None or b.shape != v.shape or b.dtype != v.dtype:
                            b = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                            bufs[(di, key)] = b
                        b.copy_(v, non_blocking=True)
                        o[key] = b
                    else:
                        o[key] = v
                out.append(o)
        ready = torch.cuda.Event()
        ready.record(self.stream)
        self._next = (tuple(out), ready, k)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        batch, ready, k = self._next
        cur = torch.cuda.current_stream(self.device)
        # everything queued so far used (at most) the OTHER set: it may be overwritten once the stream gets here
        done = torch.cuda.Event()
        done.record(cur)
        self._free[1 - k] = done
        cur.wait_event(ready)
        self._preload()
        return batch
