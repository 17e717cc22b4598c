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
import torch

from ..bert_engine import flatten_offsets
from ..host_index import forward_plan

CSR_KEY = "bert_offsets_csr"   # int32 [4, n_words]: (row, word, st, ed) of every word of the list
PLAN_KEY = "ruart_plan"        # host_index.forward_plan(...) of the batch, stored in ocr_list
TOTALS_KEY = "bert_totals"     # (real wordpieces of the list, longest 512-token window): lets the token
                               # packing of the BERT pass skip its one device -> host read-back
WINDOW = 512                   # Bert.BERT_MAX_LEN (Bert.py:18)


def attach_index_tensors(q_list, ocr_list, od_list):
    for d in (q_list, ocr_list, od_list):
        d[CSR_KEY] = flatten_offsets(d["bert_offsets"], len(d["bert_offsets"]))
        m = d["bert_mask"].cpu() != 0
        longest = max((int(m[:, p:p + WINDOW].sum(1).max()) for p in range(0, m.shape[1], WINDOW)), default=0) \
            if m.shape[0] else 0
        d[TOTALS_KEY] = (int(m.sum()), longest)
    ocr_list[PLAN_KEY] = forward_plan(
        ocr_list["num_cnt"], ocr_list["len_cnt"], od_list["num_cnt"], od_list["len_cnt"],
        ocr_list["fasttext"].size(1), od_list["fasttext"].size(1),
        ocr_list["position"].size(1), od_list["position"].size(1))
    return q_list, ocr_list, od_list


def pin(batch):
    """Tensors of the three dicts into pinned host memory (lists / plans untouched)."""
    return tuple({k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in d.items()} for d in batch)


def to_cuda(batch, device="cuda"):
    """ToCUDA (SDNetTrainer.py:208-230) with non_blocking copies; host lists / plans stay on the host."""
    return tuple({k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()}
                 for d in batch)
