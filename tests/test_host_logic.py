"""CPU: the host-side index logic that replaces the reference's Python loops — checked against
straightforward loop restatements of Models/SDNet.py:300-318,498-550 and Models/Bert/Bert.py:153-165."""
import numpy as np
import torch

from ruart_b200 import synth
from ruart_b200.bert_engine import flatten_offsets
from ruart_b200.Models.SDNet import SDNet


def test_flatten_offsets_matches_nested_loops():
    _, ocr, _ = synth.make_batch("small", ragged=True)
    offs = ocr["bert_offsets"]
    w = flatten_offsets(offs, len(offs))
    k = 0
    for i, item in enumerate(offs):
        for j, (st, ed) in enumerate(item):
            assert tuple(w[:, k]) == (i, j, st, ed)
            k += 1
    assert k == w.shape[1]
    # bertify's flat [1, 1] for an item without words (VQA_Dataset.py:426-427)
    w2 = flatten_offsets([[[1, 2], [2, 4]], [1, 1], [[1, 3]]], 3)
    assert w2.T.tolist() == [[0, 0, 1, 2], [0, 1, 2, 4], [1, 0, 1, 1], [2, 0, 1, 3]]


def test_item_index_matches_reference_loops():
    W, M = 20, 100
    _, ocr, _ = synth.make_batch("small", ragged=True)
    num_cnt, len_cnt = ocr["num_cnt"], ocr["len_cnt"]
    idx = SDNet._item_index(num_cnt, len_cnt, W, M)
    B = len(num_cnt)
    # slot scatter + mask (SDNet.py:300-318)
    it = 0
    for b in range(B):
        for k, n in enumerate(len_cnt[b]):
            assert idx["item_img"][it] == b and idx["item_slot"][it] == k and idx["lens"][it] == n
            it += 1
        assert idx["mask"][b].tolist() == [1] * num_cnt[b] + [0] * (M - num_cnt[b])
    assert it == idx["n_items"]
    # pre-align pack / unpack (SDNet.py:498-520,540-550)
    T_max = max(sum(l) for l in len_cnt)
    assert idx["T_max"] == T_max
    src, dst = [], []
    it = 0
    for b in range(B):
        t = 0
        for n in len_cnt[b]:
            for w in range(n):
                src.append(it * W + w)
                dst.append(b * T_max + t + w)
            t += n
            it += 1
    assert idx["word_src"].tolist() == src and idx["word_dst"].tolist() == dst
    assert idx["total_words"] == len(src)


def test_make_batch_layout_matches_collate_contract():
    q, ocr, od = synth.make_batch("tiny", ragged=True)
    B = q["glove"].shape[0]
    assert q["glove"].shape == (B, 40) and q["bert"].shape == (B, 50) and q["glove"].dtype == torch.int64
    assert ocr["fasttext"].shape[1] == 20 and ocr["bert"].shape[1] == 30 and ocr["position"].shape == (B, 100, 8)
    assert od["fasttext"].shape[1] == 10 and od["bert"].shape[1] == 10 and od["position"].shape == (B, 30, 8)
    assert sum(ocr["num_cnt"]) == ocr["fasttext"].shape[0] == len(ocr["bert_offsets"])
    # every image ends with the <OCR> / <OD> end item (VQA_Dataset.py:336-349)
    last = np.cumsum(ocr["num_cnt"]) - 1
    assert (ocr["fasttext"][last, 0] == 3).all() and (od["fasttext"][np.cumsum(od["num_cnt"]) - 1, 0] == 4).all()
    assert torch.equal(ocr["fasttext_mask"], ~ocr["fasttext"].eq(0)) and torch.equal(q["bert_mask"], ~q["bert"].eq(0))


def test_unsupported_options_are_rejected_loudly():
    import pytest
    opt = synth.make_opt("tiny")
    opt["img_feature"] = True
    with pytest.raises(NotImplementedError):
        SDNet(opt, synth.make_embedding())
    opt = synth.make_opt("tiny", ocr_embedding="phoc,fasttext,pos,ent,bert")   # 'phoc' without the PHOC option
    with pytest.raises(KeyError):
        SDNet(opt, synth.make_embedding())
    opt = synth.make_opt("tiny")
    opt["position_mod"] = "cat"
    with pytest.raises(NotImplementedError):
        SDNet(opt, synth.make_embedding())


def test_collate_index_tensors_equal_the_per_forward_ones():
    # the collate-side CSR / plan are what the forward would build itself
    from ruart_b200 import host_index
    from ruart_b200.Utils import collate
    q, ocr, od = synth.make_batch("small", ragged=True)
    keys = [set(d) for d in (q, ocr, od)]
    q2, ocr2, od2 = collate.attach_index_tensors(q, ocr, od)
    for d, k in zip((q2, ocr2, od2), keys):
        assert k <= set(d)                                   # every original key is kept
        csr = d[collate.CSR_KEY]
        assert csr.dtype == np.int32 and np.array_equal(csr, flatten_offsets(d["bert_offsets"], len(d["bert_offsets"])))
        assert flatten_offsets(csr, len(d["bert_offsets"])) is not None   # the array form passes through
        assert d[collate.TOTALS_KEY] == (int(d["bert_mask"].sum()), int(d["bert_mask"].sum(1).max()))
    plan = ocr2[collate.PLAN_KEY]
    want = host_index.forward_plan(ocr["num_cnt"], ocr["len_cnt"], od["num_cnt"], od["len_cnt"], 20, 10, 100, 30)
    assert plan["key"] == want["key"] == (8, sum(ocr["num_cnt"]), sum(od["num_cnt"]), 20, 10, 100, 30)
    for k in ("i32", "slots", "masks"):
        assert np.array_equal(plan[k], want[k])
    # multi2one schedule: step t holds the items with more than t words, longest first
    lens = np.array([l for img in ocr["len_cnt"] for l in img] + [l for img in od["len_cnt"] for l in img])
    assert plan["n_t"] == [int((lens > t).sum()) for t in range(lens.max())]
    assert plan["n_step_rows"] == int(lens.sum()) and plan["n_items"] == lens.size
    last = plan["i32"][plan["cuts"][5]:plan["cuts"][6]]
    assert np.array_equal(np.sort(last)[::-1], last) and np.array_equal(np.sort(last), np.sort(lens - 1))
    import pytest
    with pytest.raises(ValueError):
        flatten_offsets(np.zeros((3, 5), np.int32), 4)


def test_collate_drop_in_rebuilds_the_reference_layout():
    # Utils.collate.VQA_collate vs the collated synth batches (oracle/check_collate.py shows, in the build
    # container, that the unmodified reference's collate emits exactly this layout from the same samples)
    from ruart_b200.Utils.collate import CSR_KEY, PLAN_KEY, TOTALS_KEY, VQA_collate
    for cfg, ragged in (("tiny", True), ("small", False)):
        opt = synth.make_opt(cfg)
        batch = synth.make_batch(cfg, ragged=ragged)
        samples = synth.uncollate(batch)
        assert len(samples) == len(batch[1]["num_cnt"]) and set(samples[0]) == {"q", "ocr", "od", "gt", "extra_info"}
        q, ocr, od, gt, extra = VQA_collate(opt).VQA_collate_fun(samples)
        for got, want in ((q, batch[0]), (ocr, batch[1]), (od, batch[2])):
            assert set(want) <= set(got) and set(got) - set(want) <= {CSR_KEY, PLAN_KEY, TOTALS_KEY}
            for k, v in want.items():
                if torch.is_tensor(v):
                    assert got[k].dtype == v.dtype and torch.equal(got[k], v), k
                else:
                    assert got[k] == v, k
        assert CSR_KEY in q and PLAN_KEY in ocr and ocr[TOTALS_KEY][0] == int(batch[1]["bert_mask"].sum())
        assert gt.shape == (len(samples), opt["max_ocr_num"] + 1) and [e["q_id"] for e in extra] == list(range(len(samples)))
    import pytest
    opt = synth.make_opt("tiny", max_ocr_len=1)          # an item with 2 words no longer fits: error, like the reference
    with pytest.raises(ValueError):
        VQA_collate(opt).VQA_collate_fun(synth.uncollate(synth.make_batch("tiny")))


def test_item_index_handles_images_without_items():
    # num_cnt[i] == 0 (the reference's loops just skip such an image, SDNet.py:300-318,498-550)
    W, M = 20, 100
    num_cnt = [2, 0, 3, 0]
    len_cnt = [[1, 2], [], [2, 1, 1], []]
    idx = SDNet._item_index(num_cnt, len_cnt, W, M)
    assert idx["T_max"] == 4 and idx["total_words"] == 7 and idx["n_items"] == 5
    src, dst, it = [], [], 0
    for b in range(4):
        t = 0
        for n in len_cnt[b]:
            for w in range(n):
                src.append(it * W + w)
                dst.append(b * 4 + t + w)
            t += n
            it += 1
    assert idx["word_src"].tolist() == src and idx["word_dst"].tolist() == dst
    assert idx["mask"][1].sum() == 0 and idx["mask"][3].sum() == 0 and idx["mask"][2].sum() == 3
    # the empty image last, and a wholly empty batch
    idx2 = SDNet._item_index([1, 0], [[1], []], W, M)
    assert idx2["T_max"] == 1 and idx2["word_dst"].tolist() == [0]
    idx3 = SDNet._item_index([], [], W, M)
    assert idx3["T_max"] == 0 and idx3["n_items"] == 0


def test_shard_batch_drops_whole_batch_index_keys_and_rejects_empty_shards():
    import pytest
    from ruart_b200.Utils import collate
    batch = collate.attach_index_tensors(*synth.make_batch("small", ragged=True))
    for r in range(2):
        q, ocr, od = synth.shard_batch(batch, r, 2)
        for d in (q, ocr, od):
            assert "ruart_plan" not in d and "bert_offsets_csr" not in d and "bert_totals" not in d
        assert len(ocr["num_cnt"]) == 4 and q["glove"].shape[0] == 4
        assert ocr["fasttext"].shape[0] == sum(ocr["num_cnt"]) == len(ocr["bert_offsets"])
        # re-attached per shard
        q2, ocr2, od2 = collate.attach_index_tensors(q, ocr, od)
        assert ocr2["ruart_plan"]["key"][0] == 4
    with pytest.raises(ValueError, match="empty shard"):
        synth.shard_batch(synth.make_batch("tiny"), 3, 4)


def test_collate_rejects_offsets_past_the_real_wordpieces_and_non_prefix_masks():
    # the packed encoder has no pad positions: what the reference would read from one (Bert.py:153-165) is an
    # error here, raised on the host before anything is uploaded; degenerate spans (st >= ed) stay legal
    import copy
    import pytest
    from ruart_b200.Utils import collate
    q, ocr, od = synth.make_batch("tiny", ragged=True)
    collate.attach_index_tensors(*copy.deepcopy((q, ocr, od)))            # the synthetic batch itself is clean
    row_len = int(ocr["bert_mask"][0].sum())
    bad = copy.deepcopy((q, ocr, od))
    bad[1]["bert_offsets"][0][0] = [1, row_len + 1]
    with pytest.raises(ValueError, match="spans wordpieces"):
        collate.attach_index_tensors(*bad)
    ok = copy.deepcopy((q, ocr, od))
    ok[1]["bert_offsets"][0][0] = [row_len + 3, row_len + 3]               # st == ed: reads nothing
    collate.attach_index_tensors(*ok)
    holes = copy.deepcopy((q, ocr, od))
    holes[0]["bert_mask"][0, 1] = False                                    # a hole inside the row
    with pytest.raises(ValueError, match="contiguous prefix"):
        collate.attach_index_tensors(*holes)
