"""GPU: every differentiable op of ruart_b200/autograd_ops.py (forward AND backward kernels, through the
C ABI) against torch autograd on the same fp32 inputs; then the `Layers.py`-level forwards that the
fused inference path never calls (VERDICT r1 missing #6) against the reference's formulas."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-5   # fp32 kernels vs torch fp32 (different summation orders); GEMMs use 3-part splits (~2^-24)


def rel(a, b):
    a, b = a.double(), b.double()
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def leaf(*shape, gen=None, scale=1.0):
    return (torch.randn(*shape, generator=gen, device="cpu") * scale).cuda().requires_grad_(True)


def grads(out, inputs, gen):
    w = torch.randn(out.shape, generator=gen).cuda()
    return torch.autograd.grad((out * w).sum(), inputs, allow_unused=True)


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.mark.parametrize("rows,Kin,N,bias,relu", [(70, 300, 250, False, True), (513, 1388, 1200, True, False),
                                                  (33, 250, 1, True, False), (5, 8, 125, False, True),
                                                  (128, 500, 96, False, False)])
def test_linear_forward_dgrad_wgrad(rows, Kin, N, bias, relu):
    from ruart_b200 import autograd_ops as A
    g = torch.Generator().manual_seed(rows + N)
    x = leaf(rows, Kin, gen=g)
    w = leaf(N, Kin, gen=g, scale=Kin ** -0.5)
    b = leaf(N, gen=g) if bias else None
    y = A.linear(x, w, b, relu=relu, parts=3)
    ref = F.linear(x, w, b)
    ref = torch.relu(ref) if relu else ref
    assert rel(y, ref) < TOL
    ins = [x, w] + ([b] if bias else [])
    g1 = torch.Generator().manual_seed(7)
    got = grads(y, ins, g1)
    g1 = torch.Generator().manual_seed(7)
    want = grads(ref, ins, g1)
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < TOL


@pytest.mark.parametrize("trans_b", [False, True])
def test_bmm_forward_backward(trans_b):
    from ruart_b200 import autograd_ops as A
    g = torch.Generator().manual_seed(3)
    a = leaf(5, 100, 70, gen=g)
    b = leaf(5, 37, 70, gen=g) if trans_b else leaf(5, 70, 37, gen=g)
    y = A.bmm(a, b, trans_b=trans_b)
    ref = a.bmm(b.transpose(1, 2) if trans_b else b)
    assert rel(y, ref) < TOL
    got = grads(y, [a, b], torch.Generator().manual_seed(1))
    want = grads(ref, [a, b], torch.Generator().manual_seed(1))
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < TOL


def test_masked_softmax_and_mask_fill():
    from ruart_b200 import autograd_ops as A
    g = torch.Generator().manual_seed(5)
    x = leaf(4, 9, 41, gen=g, scale=3.0)
    mask = (torch.rand(4, 41, generator=g) > 0.3)
    mask[:, 0] = True
    mask = mask.cuda()
    y = A.masked_softmax(x, mask)
    ref = torch.softmax(x.masked_fill(~mask.unsqueeze(1), float("-inf")), -1)
    assert rel(y, ref) < TOL and (y[~mask.unsqueeze(1).expand_as(y)] == 0).all()
    got = grads(y, [x], torch.Generator().manual_seed(2))[0]
    want = grads(ref, [x], torch.Generator().manual_seed(2))[0]
    assert rel(got, want) < TOL
    # 2-D form, no mask, -inf inputs (the final softmax over [ES | OCR | no-answer], Layers.py:413-418)
    z = leaf(6, 30, gen=g)
    zm = A.mask_fill_neg_inf(z, mask[:1, :30].expand(6, 30))
    assert torch.equal(torch.isinf(zm), ~mask[:1, :30].expand(6, 30))
    p = A.masked_softmax(zm, None)
    pref = torch.softmax(z.masked_fill(~mask[:1, :30].expand(6, 30), float("-inf")), -1)
    assert rel(p, pref) < TOL
    got = grads(p, [z], torch.Generator().manual_seed(4))[0]
    want = grads(pref, [z], torch.Generator().manual_seed(4))[0]
    assert rel(got, want) < TOL
    # a row without a live key is NaN, like torch
    dead = A.masked_softmax(x.detach()[:1], torch.zeros(1, 41, dtype=torch.bool, device="cuda"))
    assert torch.isnan(dead).all()


def test_scale_cols_whole_layernorm_embedding_permute():
    from ruart_b200 import autograd_ops as A
    g = torch.Generator().manual_seed(6)
    x = leaf(3, 17, 250, gen=g)
    for d in (leaf(1, 1, 250, gen=g), leaf(1, 1, 1, gen=g)):
        y = A.scale_cols(x, d)
        ref = x * d.expand_as(x)
        assert rel(y, ref) < TOL
        for a_, b_ in zip(grads(y, [x, d], torch.Generator().manual_seed(1)),
                          grads(ref, [x, d], torch.Generator().manual_seed(1))):
            assert rel(a_, b_) < TOL
    # F.layer_norm(x, x.size()) (Layers.py:167-168)
    h = leaf(4, 23, 250, gen=g, scale=2.0)
    y = A.whole_layernorm(h + 0.5)
    ref = F.layer_norm(h + 0.5, (h + 0.5).size())
    assert rel(y, ref) < TOL
    assert rel(grads(y, [h], torch.Generator().manual_seed(8))[0], grads(ref, [h], torch.Generator().manual_seed(8))[0]) < 5e-5
    # nn.Embedding with repeated ids
    w = leaf(50, 12, gen=g)
    ids = torch.randint(0, 50, (7, 20), generator=g).cuda()
    y = A.embedding(ids, w)
    ref = F.embedding(ids, w)
    assert torch.equal(y, ref)
    assert rel(grads(y, [w], torch.Generator().manual_seed(9))[0], grads(ref, [w], torch.Generator().manual_seed(9))[0]) < TOL
    # pre-align style pack: unique source / destination rows
    src = leaf(30, 300, gen=g)
    si = torch.randperm(30, generator=g)[:18].int().cuda()
    di = torch.randperm(40, generator=g)[:18].int().cuda()
    y = A.permute_rows(src, si, di, 40)
    ref = torch.zeros(40, 300, device="cuda").index_put((di.long(),), src[si.long()])
    assert torch.equal(y, ref)
    assert torch.equal(grads(y, [src], torch.Generator().manual_seed(3))[0], grads(ref, [src], torch.Generator().manual_seed(3))[0])


@pytest.mark.parametrize("B,L,Kin,bidir", [(5, 40, 300, True), (3, 100, 1250, True), (9, 7, 64, False)])
def test_lstm_layer_matches_torch_lstm_forward_and_bptt(B, L, Kin, bidir):
    from ruart_b200 import autograd_ops as A
    H = 125
    torch.manual_seed(B * L)
    rnn = torch.nn.LSTM(Kin, H, num_layers=1, bidirectional=bidir, batch_first=True).cuda()
    g = torch.Generator().manual_seed(11)
    x = leaf(B, L, Kin, gen=g)
    y = A.lstm_layer(x, rnn, parts=3)
    ref = rnn(x)[0]
    assert rel(y, ref) < 5e-5
    params = list(rnn.parameters())
    got = grads(y, [x] + params, torch.Generator().manual_seed(12))
    want = grads(ref, [x] + params, torch.Generator().manual_seed(12))
    for (n, _), a_, b_ in zip([("x", None)] + list(rnn.named_parameters()), got, want):
        assert rel(a_, b_) < 1e-4, n


def test_multi2one_real_steps_matches_padded_lstm_and_its_gradients():
    # SDNet.py:270-271,300-318: uni-LSTM over ALL padded word steps, then out[item][len-1] -> slot; the
    # step-synchronous form runs the real steps only — same values, same gradients
    from ruart_b200 import autograd_ops as A, host_index
    H, XD, W = 300, 1388, 20
    torch.manual_seed(3)
    rnn = torch.nn.LSTM(XD, H, num_layers=1, batch_first=True).cuda()
    num_cnt = [3, 2]
    len_cnt = [[2, 1, 4], [1, 3]]
    od_num, od_len = [1, 2], [[1], [2, 1]]
    M, M_od, Wd = 5, 3, 10
    plan = host_index.forward_plan(num_cnt, len_cnt, od_num, od_len, W, Wd, M, M_od)
    N_ocr, N_od = 5, 3
    g = torch.Generator().manual_seed(2)
    items = leaf(N_ocr * W + N_od * Wd, XD, gen=g, scale=0.3)
    cuts = plan["cuts"]
    i32 = torch.from_numpy(plan["i32"]).cuda()
    a_rows, last = i32[cuts[4]:cuts[5]], i32[cuts[5]:cuts[6]]
    slot_off = torch.from_numpy(plan["slots"] * H).cuda()
    n_slots = 2 * M + 2 * M_od
    slots = A.multi2one(items, rnn, (a_rows, last, slot_off, plan["n_t"], n_slots), parts=3)
    # reference: padded LSTM per list, pick step len-1, scatter
    ocr = items[:N_ocr * W].view(N_ocr, W, XD)
    od = items[N_ocr * W:].view(N_od, Wd, XD)
    o1, o2 = rnn(ocr)[0], rnn(od)[0]
    ref = torch.zeros(n_slots, H, device="cuda")
    rows = []
    it = 0
    for b, lens in enumerate(len_cnt):
        for k, n in enumerate(lens):
            rows.append((b * M + k, o1[it, n - 1]))
            it += 1
    it = 0
    for b, lens in enumerate(od_len):
        for k, n in enumerate(lens):
            rows.append((2 * M + b * M_od + k, o2[it, n - 1]))
            it += 1
    ref = ref.index_put((torch.tensor([r for r, _ in rows], device="cuda"),), torch.stack([v for _, v in rows]))
    assert rel(slots, ref) < 5e-5
    params = list(rnn.parameters())
    got = grads(slots, [items] + params, torch.Generator().manual_seed(5))
    want = grads(ref, [items] + params, torch.Generator().manual_seed(5))
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < 1e-4


@pytest.mark.parametrize("bf16", [False, True])
def test_subword_mix_gradients_for_alpha_and_gamma(bf16):
    # Bert.py:149-165 + SDNet.py:573-583 on synthetic hidden states: values and d/d(alpha, gamma); bf16 hidden states
    # (what the bf16 encoder keeps) take the 16-byte-load backward kernel
    from ruart_b200 import autograd_ops as A
    g = torch.Generator().manual_seed(8)
    NL, H, N, W = 12, 768, 4, 5
    lens = [6, 9, 4, 7]
    T = sum(lens)
    hs = torch.randn(NL, T, H, generator=g).cuda()
    hs_b = hs.bfloat16() if bf16 else None
    if bf16:
        hs = hs_b.float()
    row_start = torch.tensor(np.cumsum([0] + lens[:-1]), dtype=torch.int32).cuda()
    words = []  # (item, j, st, ed)
    for i, n in enumerate(lens):
        p = 1
        for j in range(3):
            k = 1 + (i + j) % 2
            words.append((i, j, p, p + k))
            p += k
    words.append((0, 3, 2, 2))            # st == ed under a live mask -> zeros (Appendix-A quirk 5)
    wt = torch.tensor(words, dtype=torch.int32).t().contiguous().cuda()
    wmask = torch.zeros(N, W, dtype=torch.uint8)
    wmask[:, :4] = 1
    wmask[1, 1] = 0                       # a masked word is skipped
    wmask = wmask.cuda()
    alpha = leaf(NL, gen=g)
    gamma = leaf(1, 1, gen=g)
    pack = (None if bf16 else hs, hs_b, T * H, wt, wt.shape[1], row_start, wmask, N, W, NL, H)
    out = A.subword_mix(alpha, gamma, pack)
    ref_layers = []
    for l in range(NL):
        o = torch.zeros(N, W, H, device="cuda")
        for (i, j, st, ed) in words:
            if wmask[i, j] and st < ed:
                r0 = int(row_start[i])
                o[i, j] = hs[l, r0 + st:r0 + ed].sum(0) / float(ed - st) if ed - st > 1 else hs[l, r0 + st]
        ref_layers.append(o)
    a = torch.softmax(alpha, 0)
    ref = sum(ref_layers[l] * a[l] * gamma for l in range(NL))
    assert rel(out, ref) < TOL
    assert (out[0, 3] == 0).all() and (out[1, 1] == 0).all()
    got = grads(out, [alpha, gamma], torch.Generator().manual_seed(4))
    want = grads(ref, [alpha, gamma], torch.Generator().manual_seed(4))
    assert rel(got[0], want[0]) < 1e-4 and rel(got[1], want[1]) < 1e-4
    # SDNet.linear_sum on materialised layers gives the same numbers and gradients
    lay = [t.clone().requires_grad_(True) for t in ref_layers]
    out2 = A.layer_mix(lay, alpha, gamma)
    assert rel(out2, ref) < TOL
    got2 = grads(out2, [alpha, gamma, lay[3]], torch.Generator().manual_seed(4))
    a2 = torch.softmax(alpha, 0)
    want2 = grads(sum(lay[l] * a2[l] * gamma for l in range(NL)), [alpha, gamma, lay[3]], torch.Generator().manual_seed(4))
    for a_, b_ in zip(got2, want2):
        assert rel(a_, b_) < 1e-4


# ------------------------------------------------------------------------------------------------
# `Layers.py`-level forwards (reference formulas: Layers.py:208-245,328-341,421-432,446-468,529-534)
def test_layers_level_forwards_match_reference_formulas():
    from ruart_b200.Models import Layers
    Layers.set_dropout_prob(0.0)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    B, L1, L2, D, Hd = 3, 11, 7, 64, 48
    x1 = leaf(B, L1, D, gen=g)
    x2 = leaf(B, L2, D, gen=g)
    mask = torch.ones(B, L2, dtype=torch.bool)
    mask[0, 5:] = False
    mask[2, 3:] = False
    mask = mask.cuda()
    sc = Layers.AttentionScore(D, Hd, correlation_func=3).cuda()
    with torch.no_grad():
        sc.diagonal.copy_(1 + 0.2 * torch.randn(1, 1, Hd, generator=g))
    s = sc(x1, x2)
    a = torch.relu(F.linear(x1, sc.linear.weight)) * sc.diagonal
    b = torch.relu(F.linear(x2, sc.linear.weight))
    ref = a.bmm(b.transpose(1, 2))
    assert s.shape == (B, L1, L2) and rel(s, ref) < TOL
    got = grads(s, [x1, sc.linear.weight, sc.diagonal], torch.Generator().manual_seed(2))
    want = grads(ref, [x1, sc.linear.weight, sc.diagonal], torch.Generator().manual_seed(2))
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < 1e-4
    # Attention.forward in train mode (differentiable form) == eval mode (fused kernels)
    att = Layers.Attention(D, Hd, correlation_func=3).cuda()
    att.train()
    y_train = att(x1, x2, mask)
    att.eval()
    with torch.no_grad():
        y_eval = att(x1.detach(), x2.detach(), mask)
    ref_att = torch.softmax((torch.relu(F.linear(x1, att.scoring.linear.weight)) * att.scoring.diagonal).bmm(
        torch.relu(F.linear(x2, att.scoring.linear.weight)).transpose(1, 2)).masked_fill(~mask.unsqueeze(1), float("-inf")), -1).bmm(x2)
    assert rel(y_train, ref_att) < TOL and rel(y_eval, ref_att) < 1e-4
    # LinearSelfAttn + weighted_avg
    lsa = Layers.LinearSelfAttn(D).cuda()
    alpha = lsa(x2, mask)
    ref_alpha = torch.softmax(lsa.linear(x2).squeeze(2).masked_fill(~mask, float("-inf")), 1)
    assert alpha.shape == (B, L2) and rel(alpha, ref_alpha) < TOL
    avg = Layers.weighted_avg(x2, alpha)
    assert rel(avg, ref_alpha.unsqueeze(1).bmm(x2).squeeze(1)) < TOL
    with torch.no_grad():
        assert rel(lsa.pooled(x2.detach(), mask), avg.detach()) < 1e-5      # the fused kernel agrees
    # BilinearSeqAttn and GetFinalScores.get_single_score
    y = leaf(B, 20, gen=g)
    bil = Layers.BilinearSeqAttn(D, 20).cuda()
    o = bil(x2, y, mask)
    ref_o = x2.bmm(bil.linear(y).unsqueeze(2)).squeeze(2).masked_fill(~mask, float("-inf"))
    assert torch.equal(torch.isinf(o), torch.isinf(ref_o))
    fin = ~torch.isinf(ref_o)
    assert rel(o[fin], ref_o[fin]) < TOL
    o2 = bil(x2, y, mask, mask_flag=False)
    assert not torch.isinf(o2).any()
    gfs = Layers.GetFinalScores(D, 20, yesno=False, no_answer=True, useES=True).cuda()
    single = gfs.get_single_score(x2, y, mask, gfs.noanswer_linear, gfs.noanswer_w)
    xWh = x2.bmm(gfs.noanswer_linear(y).unsqueeze(2)).squeeze(2).masked_fill(~mask, float("-inf"))
    ref_single = gfs.noanswer_w(torch.softmax(xWh, 1).unsqueeze(1).bmm(x2)).squeeze(2)
    assert single.shape == (B, 1) and rel(single, ref_single) < TOL
    got = grads(single, [x2, y, gfs.noanswer_linear.weight, gfs.noanswer_w.weight], torch.Generator().manual_seed(6))
    want = grads(ref_single, [x2, y, gfs.noanswer_linear.weight, gfs.noanswer_w.weight], torch.Generator().manual_seed(6))
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < 1e-4


def test_stacked_brnn_train_mode_equals_eval_mode_and_torch():
    from ruart_b200.Models import Layers
    Layers.set_dropout_prob(0.0)
    torch.manual_seed(4)
    rnn = Layers.StackedBRNN(300, 125, num_layers=2).cuda()
    g = torch.Generator().manual_seed(3)
    x = leaf(4, 30, 300, gen=g)
    rnn.train()
    out, layers = rnn(x, None, return_list=True, LN=True)
    rnn.eval()
    with torch.no_grad():
        out_eval = rnn(x.detach(), None, LN=True)
    rnn.train()            # (cuDNN's RNN backward needs train mode)
    cur = x
    for i in range(2):
        cur = rnn.rnns[i](cur)[0]
        cur = F.layer_norm(cur, cur.size())
    assert rel(out, cur) < 5e-5 and rel(out_eval, cur) < 1e-4
    params = list(rnn.parameters())
    got = grads(out, [x] + params, torch.Generator().manual_seed(5))
    want = grads(cur, [x] + params, torch.Generator().manual_seed(5))
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < 2e-4


def test_train_mode_dropout_masks_scale_and_gradient():
    from ruart_b200.Models import Layers
    torch.manual_seed(0)
    x = torch.ones(8, 50, 64, device="cuda", requires_grad=True)
    Layers.set_seq_dropout(True)
    y = Layers.dropout(x, p=0.3, training=True)
    kept = (y != 0).float()
    assert torch.equal(kept[:, 0], kept[:, 17])                 # one mask per (batch, feature), shared along the sequence
    assert abs(float(kept.mean()) - 0.7) < 0.08 and rel(y[y != 0], torch.full_like(y[y != 0], 1 / 0.7)) < 1e-6
    (gx,) = torch.autograd.grad(y.sum(), [x])
    assert torch.equal(gx, y.detach())
    assert Layers.dropout(x, p=0.3, training=False) is x
    Layers.set_seq_dropout(False)
    y2 = Layers.dropout(x, p=0.5, training=True)
    assert not torch.equal((y2 != 0)[:, 0], (y2 != 0)[:, 17])
    Layers.set_seq_dropout(True)


@pytest.mark.parametrize("V,D,n,ids_kind,dtype", [
    (5000, 300, 40000, "uniform", torch.int64),      # the word tables of a cfg-5 step
    (5000, 300, 20000, "zipf", torch.int64),         # a few rows own most positions (buckets across many slabs)
    (60000, 300, 3000, "uniform", torch.int32),      # real-vocabulary size: most rows empty
    (50, 12, 30000, "uniform", torch.int64),         # pos table: every bucket spans slabs
    (20, 8, 777, "single", torch.int64),             # one row owns everything
    (7, 33, 64, "uniform", torch.int64), (7, 33, 65, "uniform", torch.int64), (3, 5, 1, "uniform", torch.int64),
    (700, 321, 5000, "out_of_range", torch.int64),   # D > one column pass; ids outside the table are skipped
])
def test_embedding_grad_sorted_form(V, D, n, ids_kind, dtype):
    """ruart_embedding_grad, sorted form (counting sort of the live positions + slab sums) against an fp64
    index_add, against the first form (one warp per row, taken with a small workspace) and against itself
    (deterministic: bit-identical repeats); `accumulate` adds to what dW holds."""
    from ruart_b200 import _lib
    from ruart_b200._lib import current_stream, ptr
    from ruart_b200.ops import call
    g = torch.Generator().manual_seed(V * 31 + n)
    if ids_kind == "uniform":
        ids = torch.randint(0, V, (n,), generator=g)
    elif ids_kind == "zipf":
        ids = torch.from_numpy(np.minimum(np.random.default_rng(n).zipf(1.2, n) - 1, V - 1))
    elif ids_kind == "single":
        ids = torch.full((n,), V - 1)
    else:
        ids = torch.randint(-3, V + 3, (n,), generator=g)
    ids = ids.to(dtype).cuda()
    dy = torch.randn(n, D, generator=g).cuda()
    dy[torch.rand(n, generator=g).cuda() < 0.4] = 0          # pad-word slots: exactly zero rows
    valid = (ids >= 0) & (ids < V)
    ref = torch.zeros(V, D, dtype=torch.float64, device="cuda").index_add_(0, ids[valid].long(), dy[valid].double())
    big = int(_lib.lib().ruart_embedding_grad_workspace_bytes(n, V, D))
    small = (n + 15) // 16 * 16 + (64 * V * D * 4 if V < 2048 else 0)
    assert big > small or V < 2048
    is64 = 1 if dtype == torch.int64 else 0

    def run(ws_bytes, accumulate=0, init=None):
        ws = torch.empty(ws_bytes + 64, dtype=torch.uint8, device="cuda")
        ws[ws_bytes:] = 0x5A                                   # guard band behind the workspace
        dw = torch.full((V + 2, D), 7.25, device="cuda")       # guard rows around the table
        if init is not None:
            dw[1:V + 1] = init
        call("ruart_embedding_grad", ptr(ids), is64, n, ptr(dy), D, D, V, ptr(ws), ws_bytes, dw.data_ptr() + 4 * D, D,
             accumulate, current_stream())
        torch.cuda.synchronize()
        assert (ws[ws_bytes:] == 0x5A).all() and (dw[0] == 7.25).all() and (dw[V + 1] == 7.25).all()
        return dw[1:V + 1].clone()

    got = run(big)
    scale = float(ref.abs().max().clamp_min(1.0))
    assert float((got.double() - ref).abs().max()) < 2e-5 * scale
    assert torch.equal(got, run(big))                          # deterministic
    first = run(small) if small < big else None                # the scan form
    if first is not None:
        assert float((first.double() - ref).abs().max()) < 2e-5 * scale
        # rows whose live positions fit one 64-entry slab are summed in the same order by both forms
        cnt = torch.bincount(ids[valid & (dy != 0).any(1)].long(), minlength=V)
        light = cnt <= 1
        assert torch.equal(got[light], first[light])
    init = torch.randn(V, D, generator=g).cuda()
    acc = run(big, accumulate=1, init=init)
    assert float((acc.double() - (ref + init.double())).abs().max()) < 2e-5 * scale
