"""GPU: the packed BERT kernels against plain torch fp32 references of the same ops."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attention(qkv, cu, heads):
    T, H3 = qkv.shape
    H = H3 // 3
    out = torch.zeros(T, H, device=qkv.device)
    for s in range(len(cu) - 1):
        a, b = cu[s], cu[s + 1]
        if b == a:
            continue
        q, k, v = [x.view(b - a, heads, 64).transpose(0, 1) for x in qkv[a:b].float().split(H, 1)]
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0, -1)
        out[a:b] = (p @ v).transpose(0, 1).reshape(b - a, H)
    return out


@pytest.mark.parametrize("lens", [[3, 5, 8, 1, 2, 7, 6, 4] * 40, [33, 20, 42, 5, 64, 3, 16, 17], [100, 7, 513 - 400, 65, 2],
                                  [16] * 9 + [17] * 3, [1]])
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_varlen_attention(lens, dtype):
    from ruart_b200._lib import call, current_stream, ptr
    heads, H = 12, 768
    cu = [0]
    for l in lens:
        cu.append(cu[-1] + l)
    T = cu[-1]
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn(T, 3 * H, device="cuda", generator=g)
    cu_d = torch.tensor(cu, dtype=torch.int32, device="cuda")
    if dtype == "bf16":
        x = qkv.bfloat16()
        out = torch.full((T, H), float("nan"), device="cuda", dtype=torch.bfloat16)
        call("ruart_bert_attention", None, ptr(x), ptr(cu_d), len(lens), heads, 0.125, max(lens), None, ptr(out), 1,
             current_stream())
        want = _ref_attention(x, cu, heads)
        tol = 2e-2
    else:
        x = qkv
        out = torch.full((T, H), float("nan"), device="cuda")
        call("ruart_bert_attention", ptr(x), None, ptr(cu_d), len(lens), heads, 0.125, max(lens), ptr(out), None, 1,
             current_stream())
        want = _ref_attention(x, cu, heads)
        tol = 2e-5
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert (out.float() - want).abs().max().item() < tol


def test_add_layernorm_and_embed():
    from ruart_b200._lib import call, current_stream, ptr
    T, H = 1000, 768
    x = torch.randn(T, H, device="cuda")
    r = torch.randn(T, H, device="cuda")
    g = torch.randn(H, device="cuda")
    b = torch.randn(H, device="cuda")
    out = torch.empty(T, H, device="cuda")
    out16 = torch.empty(T, 3 * H, device="cuda", dtype=torch.bfloat16)
    call("ruart_add_layernorm", ptr(x), None, ptr(r), None, ptr(g), ptr(b), 1e-12, T, H, ptr(out), ptr(out16), 3,
         current_stream())
    y = x + r
    u = y.mean(-1, keepdim=True)
    s = (y - u).pow(2).mean(-1, keepdim=True)
    want = g * ((y - u) / torch.sqrt(s + 1e-12)) + b
    assert (out - want).abs().max().item() < 1e-4
    recon = out16[:, :H].float() + out16[:, H:2 * H].float() + out16[:, 2 * H:].float()
    assert (recon - want).abs().max().item() < 1e-4
    ids = torch.randint(0, 3000, (T,), device="cuda", dtype=torch.int32)
    pos = torch.randint(0, 512, (T,), device="cuda", dtype=torch.int32)
    we, pe, te = torch.randn(3000, H, device="cuda"), torch.randn(512, H, device="cuda"), torch.randn(2, H, device="cuda")
    call("ruart_bert_embed_ln", ptr(ids), ptr(pos), ptr(we), ptr(pe), ptr(te), ptr(g), ptr(b), 1e-12, T, H, ptr(out),
         None, 1, current_stream())
    y = we[ids.long()] + pe[pos.long()] + te[0]
    u = y.mean(-1, keepdim=True)
    s = (y - u).pow(2).mean(-1, keepdim=True)
    want = g * ((y - u) / torch.sqrt(s + 1e-12)) + b
    assert (out - want).abs().max().item() < 1e-4


def test_embed_raw_and_folded_subword_mix():
    """ruart_bert_embed_raw / ruart_subword_coef / ruart_subword_avg_layers_fold (the ends of the folded-LayerNorm
    encoder) against torch: LayerNorm of the stored rows, subword mean (Bert.py:149-165), learned layer sum
    (SDNet.py:573-583)."""
    from ruart_b200._lib import call, current_stream, ptr
    st = current_stream()
    T, H, NL, eps = 3300, 768, 12, 1e-12
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    ids = torch.randint(0, 3000, (T,), device="cuda", dtype=torch.int32)
    pos = torch.randint(0, 512, (T,), device="cuda", dtype=torch.int32)
    we, pe, te = rnd(3000, H), rnd(512, H), rnd(2, H)
    raw = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
    stats = torch.full((T, 8, 2), 9.0, device="cuda")
    call("ruart_bert_embed_raw", ptr(ids), ptr(pos), ptr(we), ptr(pe), ptr(te), T, H, ptr(raw), ptr(stats), st)
    y = we[ids.long()] + pe[pos.long()] + te[0]
    assert (raw.float() - y).abs().max().item() < 2e-2 * y.abs().max().item()
    assert (stats[:, 0, 0] - y.sum(1)).abs().max().item() < 1e-2
    assert ((stats[:, 0, 1] - (y * y).sum(1)).abs() / (y * y).sum(1)).max().item() < 1e-5
    assert (stats[:, 1:] == 0).all()
    # ---- layer mix over NL stored (pre-LayerNorm) layers
    hs = (rnd(NL, T, H) * 1.7 + 0.3).bfloat16()
    x = hs.float()
    stt = torch.zeros(NL, T, 8, 2, device="cuda")
    for k, chunk in enumerate(x.split(128, dim=2)):
        stt[:, :, k, 0] = chunk.sum(2)
        stt[:, :, k, 1] = (chunk * chunk).sum(2)
    ln_g, ln_b = rnd(NL, H) * 0.4 + 1.0, rnd(NL, H) * 0.2
    alpha, gam = rnd(NL), torch.tensor([[1.3]], device="cuda")
    G = torch.empty(NL, H, device="cuda")
    C = torch.empty(H, device="cuda")
    call("ruart_subword_coef", ptr(alpha), ptr(gam), NL, ptr(ln_g), ptr(ln_b), H, ptr(G), ptr(C), st)
    a = torch.softmax(alpha, 0) * 1.3
    assert (G - a[:, None] * ln_g).abs().max().item() < 1e-5 and (C - (a[:, None] * ln_b).sum(0)).abs().max().item() < 1e-5
    # items of 1..6 pieces (1-3: pipelined path, 4+: synchronous path), one masked word, one st == ed word
    N, Wd = 300, 3
    words, row_start, t0 = [], [], 0
    for it in range(N):
        row_start.append(t0)
        o = 0
        for j in range(Wd):
            cnt = (it + j) % 6 + 1
            words.append((it, j, o, o + (0 if (it, j) == (7, 1) else cnt)))
            o += cnt
        t0 += o
    assert t0 <= T
    wt = torch.tensor(words, dtype=torch.int32, device="cuda").t().contiguous()
    rs = torch.tensor(row_start, dtype=torch.int32, device="cuda")
    wmask = torch.ones(N, Wd, dtype=torch.uint8, device="cuda")
    wmask[5, 2] = 0
    dst = torch.full((N, Wd, H + 4), 5.0, device="cuda")
    call("ruart_subword_avg_layers_fold", ptr(hs), T * H, ptr(stt), T * 8, eps, ptr(wt), len(words), ptr(rs), ptr(wmask), Wd,
         dst.data_ptr(), H + 4, ptr(G), ptr(C), NL, H, st)
    torch.cuda.synchronize()
    u = x.mean(-1, keepdim=True)
    yn = (x - u) / torch.sqrt((x - u).pow(2).mean(-1, keepdim=True) + eps) * ln_g[:, None, :] + ln_b[:, None, :]
    for it, j, s, e in words[:60] + words[-30:]:
        got = dst[it, j, :H]
        if wmask[it, j] == 0 or e <= s:
            assert (got == 0).all()
            continue
        want = sum(a[l] * yn[l, row_start[it] + s:row_start[it] + e].mean(0) for l in range(NL))
        assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item()), (it, j, s, e)
    assert (dst[..., H:] == 5.0).all()


@pytest.mark.parametrize("long_row", [False, True])
def test_folded_layernorm_encoder_matches_unfolded(long_row):
    """The bf16 encoder with every LayerNorm folded into the neighbouring GEMMs (BertEngine._encode_hidden_fold)
    against the explicit-LayerNorm bf16 encoder and the fp32-mode encoder on the same packed batch.
    long_row: one 200-token row -> the attention cannot be fused into the query/key/value GEMM (a sequence must fit
    one 128-row accumulator tile) and the folded encoder runs with the stand-alone attention kernels."""
    import contextlib, io
    from ruart_b200.Models.Bert.Bert import Bert
    from ruart_b200.bert_engine import Segment
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        bert = Bert({'BERT_LINEAR_COMBINE': True, 'BERT_num_layers': 3, 'BERT_model_file': ''})
    for p in bert.parameters():                     # non-trivial LayerNorm weights
        if p.dim() == 1:
            p.data.add_(torch.randn_like(p) * 0.2)
    N, L, Wd, H = 700, (200 if long_row else 8), 3, 768
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(3, 9, (N,), generator=g)
    if long_row:
        lens[17] = 200
    ids = torch.randint(1000, 30000, (N, L), generator=g)
    mask = torch.arange(L)[None, :] < lens[:, None]
    ids = (ids * mask).cuda()
    offsets = [[[1 + j, 2 + j] for j in range(min(Wd, int(lens[i]) - 2))] for i in range(N)]
    wmask = torch.tensor([[j < len(o) for j in range(Wd)] for o in offsets]).cuda()
    seg = lambda: [Segment(ids, mask.cuda(), offsets, wmask)]
    alpha = torch.randn(3).cuda()
    gamma = torch.tensor([[0.9]]).cuda()
    outs = {}
    for name, prec, fold in (("fold", "bf16", True), ("plain", "bf16", False), ("fp32", "fp32", False)):
        bert.precision = prec
        eng = bert.engine()
        eng.fold = fold and eng.fold
        dst = torch.zeros(N, Wd, H, device="cuda")
        pk = eng.encode(seg(), [(dst, H, 0)], alpha=alpha, gamma=gamma)
        assert ("fold" in pk) == (name == "fold")
        if name == "fold":
            assert pk["fold"]["fused_attention"] == (not long_row)
        outs[name] = dst
        bert._engine = None
    torch.cuda.synchronize()
    scale = outs["fp32"].abs().max().item()
    e_fold = (outs["fold"] - outs["fp32"]).abs().max().item() / scale
    e_plain = (outs["plain"] - outs["fp32"]).abs().max().item() / scale
    assert e_plain < 3e-2 and e_fold < max(1.5 * e_plain, 1e-2), (e_fold, e_plain)


@pytest.mark.parametrize("lens", [[3, 5, 8, 1, 2, 7, 6, 4] * 300, [33, 20, 42, 5, 64, 3, 16, 17, 128, 1, 127, 2, 100, 28] * 40,
                                  [128] * 20 + [1] * 300 + [0, 0, 5] * 50, [7] * 4000])
def test_seq_tiles(lens):
    """ruart_seq_tiles: greedy tiles of whole sequences, <= 128 rows each, and per-token sequence bounds."""
    from ruart_b200._lib import call, current_stream, ptr
    cu = [0]
    for l in lens:
        cu.append(cu[-1] + l)
    T, S = cu[-1], len(lens)
    cu_d = torch.tensor(cu, dtype=torch.int32, device="cuda")
    cap = 2 * T // 128 + 8
    meta = torch.full((cap,), -7, dtype=torch.int32, device="cuda")
    bounds = torch.full((T, 2), -1, dtype=torch.int32, device="cuda")
    call("ruart_seq_tiles", ptr(cu_d), S, ptr(meta), cap, ptr(bounds), current_stream())
    torch.cuda.synchronize()
    m = meta.cpu().tolist()
    n = m[0]
    starts = m[1:n + 2]
    # host restatement of the greedy rule
    want, row0 = [0], 0
    for s in range(S):
        if cu[s + 1] - row0 > 128:
            row0 = cu[s]
            want.append(row0)
    want.append(T)
    assert starts == want and n == len(want) - 1
    assert all(b - a <= 128 for a, b in zip(starts[:-1], starts[1:]))
    b = bounds.cpu()
    for s in (0, 1, S // 2, S - 1):
        if lens[s]:
            assert b[cu[s]:cu[s + 1]].tolist() == [[cu[s], cu[s + 1]]] * lens[s]


@pytest.mark.parametrize("lens", [[3, 5, 8, 1, 2, 7, 6, 4] * 80, [33, 20, 42, 5, 64, 3, 16, 17, 128, 1, 127, 2, 100, 28] * 6,
                                  [50] * 30 + [4, 6] * 200])
def test_qkv_attention_fold(lens):
    """ruart_qkv_attention_fold (query/key/value GEMM with the folded LayerNorm and the attention in its epilogue,
    modeling.py:224-250) against LayerNorm -> three Linears -> per-sequence softmax attention in torch fp32."""
    from ruart_b200._lib import call, current_stream, ptr
    st = current_stream()
    heads, H, eps = 12, 768, 1e-12
    cu = [0]
    for l in lens:
        cu.append(cu[-1] + l)
    T, S = cu[-1], len(lens)
    assert T >= 2048
    g = torch.Generator(device="cuda").manual_seed(T)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    raw = (rnd(T, H) * 1.3 + 0.2).bfloat16()
    x = raw.float()
    stats = torch.zeros(T, 8, 2, device="cuda")
    for k, chunk in enumerate(x.split(128, dim=1)):
        stats[:, k, 0] = chunk.sum(1)
        stats[:, k, 1] = (chunk * chunk).sum(1)
    gamma, beta = rnd(H) * 0.3 + 1.0, rnd(H) * 0.2
    w0, b0 = rnd(3 * H, H) * 0.04, rnd(3 * H) * 0.1
    wf = (w0 * gamma[None, :]).bfloat16()
    colsum = wf.float().sum(1)
    cvec = w0 @ beta + b0
    perm = torch.cat([torch.arange(64, device="cuda") + part * H + h * 64 for h in range(heads) for part in range(3)])
    qs = torch.ones(3 * H, device="cuda")
    qs[:H] = 0.125
    w_p = (wf.float() * qs[:, None])[perm].bfloat16().contiguous()
    s_p, c_p = (colsum * qs)[perm].contiguous(), (cvec * qs)[perm].contiguous()
    cu_d = torch.tensor(cu, dtype=torch.int32, device="cuda")
    cap = 2 * T // 128 + 8
    meta = torch.empty(cap, dtype=torch.int32, device="cuda")
    bounds = torch.empty((T, 2), dtype=torch.int32, device="cuda")
    call("ruart_seq_tiles", ptr(cu_d), S, ptr(meta), cap, ptr(bounds), st)
    ctx = torch.full((T + 1, H), 7.0, device="cuda", dtype=torch.bfloat16)
    call("ruart_qkv_attention_fold", ptr(raw), H, ptr(w_p), H, T, H, heads, ptr(c_p), ptr(s_p), ptr(stats), eps,
         ptr(meta), ptr(bounds), ptr(ctx), H, st)
    torch.cuda.synchronize()
    u = x.mean(-1, keepdim=True)
    ln = (x - u) / torch.sqrt((x - u).pow(2).mean(-1, keepdim=True) + eps) * gamma + beta
    qkv = (ln @ w0.t() + b0).bfloat16()            # the unfused path rounds Q|K|V to bf16 as well
    want = _ref_attention(qkv, cu, heads)
    assert torch.isfinite(ctx[:T].float()).all()
    err = (ctx[:T].float() - want).abs().max().item()
    assert err < 3e-2 * max(1.0, want.abs().max().item()), err
    assert (ctx[T] == 7.0).all()
