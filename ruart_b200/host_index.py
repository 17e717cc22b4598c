"""Host-side (numpy) index building for one batch — what replaces the reference's Python loops over
`num_cnt` / `len_cnt` (Models/SDNet.py:300-318 slot scatter + mask, :498-550 pre-align pack/unpack)
and schedules the step-synchronous multi2one LSTM (real word steps only, items sorted by length).

`forward_plan` depends on the batch only (not on weights), so a collate function can build it in a
DataLoader worker (ruart_b200/Utils/collate.py); SDNet.forward builds it itself
when the batch does not carry one.
"""
import numpy as np


def item_index(num_cnt, len_cnt, W, M):
    """numpy indices replacing the loops of SDNet.py:300-318 and :498-550 for one item list."""
    B = len(num_cnt)
    num = np.asarray(num_cnt, dtype=np.int64)
    lens = np.fromiter((l for img in len_cnt for l in img), dtype=np.int64, count=int(num.sum()))
    if lens.size and lens.min() < 1:
        raise ValueError("every item must have at least one word (len_cnt >= 1)")
    if lens.size and lens.max() > W:
        raise ValueError("len_cnt exceeds the word slots of an item row")
    if B and int(num.max()) > M:
        raise ValueError("num_cnt exceeds the slot count of `position`")
    if B and int(num.min()) < 0:
        raise ValueError("num_cnt must be >= 0")
    n_items = lens.size
    item_img = np.repeat(np.arange(B, dtype=np.int64), num)
    first_item = np.cumsum(num) - num
    item_slot = np.arange(n_items, dtype=np.int64) - first_item[item_img]
    word_off = np.cumsum(lens) - lens                      # first word of each item, global
    # images WITHOUT items (num_cnt 0: the reference's loops simply skip them, SDNet.py:300-318,498-550)
    # have 0 words; reduceat / fancy indexing by `first_item` would read the next image's values instead
    img_words = np.bincount(item_img, weights=lens, minlength=B).astype(np.int64)
    img_first_word = np.cumsum(img_words) - img_words
    t0_item = word_off - img_first_word[item_img]          # word offset of the item inside its image
    T_max = int(img_words.max()) if B else 0
    total = int(lens.sum())
    item_of_word = np.repeat(np.arange(n_items, dtype=np.int64), lens)
    w = np.arange(total, dtype=np.int64) - np.repeat(word_off, lens)
    word_src = item_of_word * W + w                        # row in the [items*W] word layout
    word_dst = item_img[item_of_word] * T_max + t0_item[item_of_word] + w   # row in [B*T_max]
    mask = (np.arange(M)[None, :] < num[:, None]).astype(np.uint8)
    return dict(B=B, n_items=n_items, lens=lens, item_img=item_img, item_slot=item_slot, T_max=T_max,
                word_src=word_src, word_dst=word_dst, mask=mask, total_words=total)


def forward_plan(ocr_num_cnt, ocr_len_cnt, od_num_cnt, od_len_cnt, Wo, Wd, M, M_od):
    """Everything SDNet.forward needs from the four count lists, as three flat arrays + sizes.

    i32   : [ocr word_src | ocr word_dst | od word_src | od word_dst | multi2one step rows | last step]
    slots : int64 slot row (OCR slots first, then OD) of every item in length-sorted order
    masks : uint8 [B*M | B*M_od] slot masks (SDNet.py:306,312)
    """
    io = item_index(ocr_num_cnt, ocr_len_cnt, Wo, M)
    id_ = item_index(od_num_cnt, od_len_cnt, Wd, M_od)
    B, N_ocr, N_od = io['B'], io['n_items'], id_['n_items']
    if id_['B'] != B:
        raise ValueError("ocr_list and od_list describe different numbers of images")
    # multi2one over OCR and OD items together (shared weights, SDNet.py:270-271)
    lens_all = np.concatenate([io['lens'], id_['lens']])
    base_all = np.concatenate([np.arange(N_ocr, dtype=np.int64) * Wo,
                               N_ocr * Wo + np.arange(N_od, dtype=np.int64) * Wd])
    slot_all = np.concatenate([io['item_img'] * M + io['item_slot'],
                               B * M + id_['item_img'] * M_od + id_['item_slot']])
    perm = np.argsort(-lens_all, kind='stable')
    max_len = int(lens_all.max()) if lens_all.size else 0
    n_t = [int((lens_all > t).sum()) for t in range(max_len)]
    a_rows = np.concatenate([base_all[perm[:n]] + t for t, n in enumerate(n_t)]) if n_t \
        else np.zeros(0, np.int64)
    i32 = np.concatenate([io['word_src'], io['word_dst'], id_['word_src'], id_['word_dst'], a_rows,
                          lens_all[perm] - 1]).astype(np.int32)
    cuts = np.cumsum([0, io['total_words'], io['total_words'], id_['total_words'], id_['total_words'],
                      a_rows.size, lens_all.size])
    return dict(key=(B, N_ocr, N_od, Wo, Wd, M, M_od), i32=i32, cuts=[int(c) for c in cuts],
                slots=slot_all[perm].astype(np.int64),
                masks=np.concatenate([io['mask'].reshape(-1), id_['mask'].reshape(-1)]),
                n_t=n_t, n_step_rows=int(a_rows.size), n_items=int(lens_all.size),
                T_max=(io['T_max'], id_['T_max']), total_words=(io['total_words'], id_['total_words']))
