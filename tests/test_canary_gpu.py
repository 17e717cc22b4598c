"""GPU: guard-band ("canary") checks of the kernels added in round 2.  compute-sanitizer is closed on this GPU
pool (profiles/r02_sanitizer_unavailable.txt), so out-of-bounds writes are looked for directly: every output
lives inside a larger buffer pre-filled with a sentinel; after the launch the guard bands must be untouched and
the payload fully written (no sentinel left), for ragged sizes that do not divide the tile shapes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PAD = 4096          # guard elements on both sides
SENT = -12345.678   # sentinel (an exact fp32 value none of the kernels produces)


class Guarded(object):
    def __init__(self, shape, dtype=torch.float32):
        n = int(np.prod(shape))
        self.buf = torch.full((PAD + n + PAD,), SENT, dtype=torch.float32, device="cuda").to(dtype)
        self.sent = self.buf[0].clone()
        self.view = self.buf[PAD:PAD + n].view(*shape)
        self.n = n

    def check(self, name, all_written=True):
        torch.cuda.synchronize()
        assert (self.buf[:PAD] == self.sent).all(), "%s wrote BEFORE its output" % name
        assert (self.buf[PAD + self.n:] == self.sent).all(), "%s wrote PAST its output" % name
        if all_written:
            assert not (self.view == self.sent).any(), "%s left part of its output unwritten" % name


def test_new_kernels_stay_inside_their_outputs():
    from ruart_b200 import autograd_ops as A, ops, sdnet_ops as K
    from ruart_b200._lib import current_stream, ptr
    from ruart_b200.ops import call
    g = torch.Generator().manual_seed(0)
    R = lambda *s: torch.randn(*s, generator=g).cuda()
    st = current_stream()

    # bmm_f32: ragged M/N/K, both transposes, strided batches
    a, b = R(3, 37, 29), R(3, 53, 29)
    out = Guarded((3, 37, 53))
    A.bmm_raw(a, b, False, True, 37, 53, 29, out.view)
    out.check("bmm_f32 NT")
    assert torch.allclose(out.view, a.bmm(b.transpose(1, 2)), atol=1e-4)
    out = Guarded((3, 29, 53))
    b2 = R(3, 37, 53)
    A.bmm_raw(a, b2, True, False, 29, 53, 37, out.view)
    out.check("bmm_f32 TN")

    # masked softmax / softmax backward
    x = R(5, 7, 33)
    m = (torch.rand(5, 33, generator=g) > 0.4).cuda()
    m[:, 0] = True
    out = Guarded((5, 7, 33))
    m8 = K.as_u8(m)
    call("ruart_masked_softmax", ptr(x), 33, ptr(m8), 5, 7, 33, ptr(out.view), 33, st)
    out.check("masked_softmax")
    dx = Guarded((5, 7, 33))
    call("ruart_softmax_backward", ptr(out.view), 33, ptr(x), 33, ptr(dx.view), 33, 35, 33, st)
    dx.check("softmax_backward")

    # eltwise (pitched output), mask_fill, colsum
    big = Guarded((19, 40))
    el_a, el_v = R(19, 31), R(31)
    A.eltwise(3, el_a, v=el_v, out=big.view[:, :31])
    big.check("eltwise", all_written=False)
    assert (big.view[:, 31:] == big.sent).all()
    mf = Guarded((6, 30))
    mf_x, mf_m = R(6, 30), K.as_u8(m[:1, :30].expand(6, 30))      # (named: two unnamed temporaries could alias)
    call("ruart_mask_fill", ptr(mf_x), ptr(mf_m), 180, float("-inf"), ptr(mf.view), st)
    mf.check("mask_fill")
    cs = Guarded((77,))
    cs_x = R(1001, 77)
    A.colsum(cs_x, out=cs.view)
    cs.check("colsum")

    # transposed split (ragged rows / K), split_concat
    src = R(70, 45)
    rows_p = 128
    tout = Guarded((45, 3 * rows_p), dtype=torch.bfloat16)
    call("ruart_split_bf16_t", ptr(src), 45, 70, 45, rows_p, 3, ptr(tout.view), st)
    tout.check("split_bf16_t")
    rec = tout.view.float().view(45, 3, rows_p).sum(1)[:, :70].t()
    assert torch.allclose(rec, src, atol=1e-6) and (tout.view.float().view(45, 3, rows_p)[:, :, 70:] == 0).all()
    pieces = [R(4, 9, 300), R(4, 9, 250), R(4, 9, 125)]
    Kp = 704
    sc = Guarded((36, 2 * Kp), dtype=torch.bfloat16)
    got, kp = K.split_concat(pieces, 2)
    assert kp == Kp
    import ctypes
    srcs = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in pieces])
    pit = (ctypes.c_longlong * 3)(300, 250, 125)
    wid = (ctypes.c_int * 3)(300, 250, 125)
    call("ruart_split_concat_bf16", ctypes.addressof(srcs), ctypes.addressof(pit), ctypes.addressof(wid), 3, 36, Kp, 2,
         ptr(sc.view), st)
    sc.check("split_concat_bf16")
    cat = torch.cat(pieces, 2).view(36, 675)
    rec = sc.view.float().view(36, 2, Kp).sum(1)
    assert torch.allclose(rec[:, :675], cat, atol=1e-4) and (rec[:, 675:] == 0).all()
    ref_split, _ = K.split_act(cat, 2)
    assert torch.equal(got, ref_split) and torch.equal(sc.view, ref_split)      # bit-identical to concat + split

    # whole-LN backward, embedding grad (large and small table)
    y, dy = R(3, 17, 250), R(3, 17, 250)
    stats = torch.tensor([0.1, 1.7], device="cuda")
    dxl = Guarded((3, 17, 250))
    ln_ws = torch.empty(4096, dtype=torch.float64, device="cuda")
    call("ruart_whole_layernorm_backward", ptr(y), 250, ptr(dy), 250, 51, 250, ptr(stats), ptr(ln_ws), ptr(dxl.view),
         250, st)
    dxl.check("whole_layernorm_backward")
    for V, D in ((5000, 300), (50, 12)):
        ids = torch.randint(0, V, (777,), generator=g).cuda()
        dyr = R(777, D)
        dyr[::3] = 0
        dw = Guarded((V, D))
        ref = torch.zeros(V, D, device="cuda").index_add_(0, ids, dyr)
        from ruart_b200 import _lib
        sorted_bytes = int(_lib.lib().ruart_embedding_grad_workspace_bytes(777, V, D))
        for ws_bytes in (784 + (64 * V * D * 4 if V < 2048 else 0), sorted_bytes):   # first form, sorted form
            ws = Guarded((ws_bytes // 4,))
            call("ruart_embedding_grad", ptr(ids), 1, 777, ptr(dyr), D, D, V, ptr(ws.view), ws_bytes, ptr(dw.view), D, 0, st)
            dw.check("embedding_grad V=%d" % V)
            ws.check("embedding_grad workspace V=%d" % V, all_written=False)
            assert torch.allclose(dw.view, ref, atol=1e-5)

    # LSTM with saved gates + BPTT: B not a multiple of the CTA's sequence count, H = 125
    B, L, H = 7, 9, 125
    xg = R(B * L, 8 * H) * 0.3
    whh = R(2, 4 * H, H) * 0.05
    o = Guarded((B, L, 2 * H))
    gates = Guarded((B * L, 10 * H))
    call("ruart_lstm_recurrence_train", ptr(xg), 8 * H, ptr(whh), ptr(o.view), 2 * H, B, L, H, 2, ptr(gates.view), 10 * H, st)
    o.check("lstm_recurrence_train out")
    gates.check("lstm_recurrence_train gates")
    plain = torch.empty(B, L, 2 * H, device="cuda")
    call("ruart_lstm_recurrence", ptr(xg), 8 * H, ptr(whh), ptr(plain), 2 * H, B, L, H, 2, st)
    assert torch.allclose(plain, o.view, atol=5e-6)    # inference runs the tensor-core form of the same recurrence
    dxg = Guarded((B * L, 8 * H))
    dout = R(B, L, 2 * H)
    call("ruart_lstm_recurrence_backward", ptr(gates.view), 10 * H, ptr(whh), ptr(dout), 2 * H, ptr(dxg.view),
         8 * H, B, L, H, 2, st)
    dxg.check("lstm_recurrence_backward")

    # pipelined subword mean: words of 1..6 pieces (the > 4 piece fallback included), pitched destination
    NL, Hd, N, W = 12, 768, 5, 4
    lens = [9, 12, 5, 14, 8]
    T = sum(lens)
    hs = R(NL, T, Hd).bfloat16()
    rs = torch.tensor(np.cumsum([0] + lens[:-1]), dtype=torch.int32).cuda()
    words = [(0, 0, 1, 2), (0, 1, 2, 5), (1, 0, 1, 7), (1, 1, 7, 11), (2, 0, 1, 3), (3, 0, 1, 6), (3, 1, 6, 6), (4, 0, 2, 4)]
    wt = torch.tensor(words, dtype=torch.int32).t().contiguous().cuda()
    wmask = torch.ones(N, W, dtype=torch.uint8).cuda()
    dst = Guarded((N * W, 1388))
    alpha, gamma = R(NL), torch.tensor([0.9], device="cuda")
    call("ruart_subword_avg_layers", None, ptr(hs), T * Hd, ptr(wt), len(words), ptr(rs), ptr(wmask), W,
         dst.view.data_ptr() + 4 * 300, 1388, ptr(alpha), NL, ptr(gamma), Hd, st)
    dst.check("subword_avg_layers", all_written=False)
    a = torch.softmax(alpha, 0)
    for (i, j, s0, e0) in words:
        row = dst.view[i * W + j]
        assert (row[:300] == dst.sent).all() and (row[300 + Hd:] == dst.sent).all()     # only the BERT columns
        want = torch.zeros(Hd, device="cuda")
        if e0 > s0:
            for l in range(NL):
                seg = hs[l, int(rs[i]) + s0:int(rs[i]) + e0].float()
                want = want + (seg.sum(0) / float(e0 - s0) if e0 - s0 > 1 else seg[0]) * a[l] * 0.9
        assert torch.allclose(row[300:300 + Hd], want, atol=2e-5), (i, j)

    # bf16 LayerNorm specialisation: T not a multiple of the 8 rows of a CTA
    Tn = 1003
    xb = R(Tn, 768).bfloat16()
    ob = Guarded((Tn, 768), dtype=torch.bfloat16)
    gm, bt = torch.rand(768, generator=g).cuda() + 0.5, R(768)
    call("ruart_add_layernorm", None, ptr(xb), None, None, ptr(gm), ptr(bt), 1e-12, Tn, 768, None, ptr(ob.view), 1, st)
    ob.check("ln_bf16")
    xf = xb.float()
    ref = ((xf - xf.mean(1, keepdim=True)) / torch.sqrt(xf.var(1, unbiased=False, keepdim=True) + 1e-12) * gm + bt)
    assert (ob.view.float() - ref).abs().max().item() < 0.05
