"""GPU: SDNet-stack kernels against the CPU oracle's building blocks on odd shapes."""
import pytest
import torch

from oracle import sdnet_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,I,H,bidir", [(5, 7, 40, 125, True), (1, 1, 300, 125, True), (33, 20, 64, 128, True),
                                           (9, 13, 50, 64, False), (130, 3, 16, 17, True), (4, 6, 30, 300, False),
                                           (256, 9, 32, 125, True), (200, 5, 24, 89, False), (600, 3, 16, 125, True)])
def test_stacked_brnn_layer_matches_oracle(B, L, I, H, bidir):
    from ruart_b200.Models import Layers
    Layers.set_dropout_prob(0.0)
    Layers.set_sdnet_precision(3)
    torch.manual_seed(B * 100 + L)
    rnn = Layers.StackedBRNN(I, H, 1, bidirectional=bidir).cuda().eval()
    x = torch.randn(B, L, I, device="cuda")
    with torch.no_grad():
        got = rnn(x, None, LN=True)
    sd = {"r." + k: v.detach().cpu() for k, v in rnn.state_dict().items()}
    want = sdnet_oracle.stacked_brnn(sd, "r", x.cpu(), 1, bidirectional=bidir, whole_ln=True)[-1]
    assert got.shape == want.shape
    assert (got.cpu() - want).abs().max().item() < 2e-4


@pytest.mark.parametrize("B,L,H,ndir", [(7, 9, 125, 2), (256, 101, 125, 2), (1, 1, 125, 2), (20, 5, 128, 1),
                                        (37, 40, 17, 2), (9, 3, 64, 1), (300, 4, 89, 2)])
def test_lstm_tensor_core_recurrence_matches_fp32_fma_recurrence(B, L, H, ndir):
    """ruart_lstm_recurrence (mma.sync on bf16 hi|lo splits, shared reciprocals) against the fp32-FMA kernel
    that the training path keeps (ruart_lstm_recurrence_train), on a strided output and xg with row padding."""
    from ruart_b200._lib import current_stream, ptr
    from ruart_b200.ops import call
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + L)
    pitch = ndir * 4 * H + 4 * (B % 3)                     # multiple of 4 floats: the TMA path
    xg = torch.randn(B * L, pitch, device="cuda", generator=g) * 1.5
    xg[:, :3] *= 30                                         # saturated gates: the capped exponents
    whh = (torch.rand(ndir, 4 * H, H, device="cuda", generator=g) * 2 - 1) / H ** 0.5
    want = torch.zeros(B, L, ndir * H + 3, device="cuda")
    got = torch.zeros_like(want)
    gates = torch.empty(B * L, ndir * 5 * H, device="cuda")
    st = current_stream()
    call("ruart_lstm_recurrence_train", ptr(xg), pitch, ptr(whh), ptr(want), ndir * H + 3, B, L, H, ndir,
         ptr(gates), ndir * 5 * H, st)
    call("ruart_lstm_recurrence", ptr(xg), pitch, ptr(whh), ptr(got), ndir * H + 3, B, L, H, ndir, st)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert (got[..., ndir * H:] == 0).all()                 # nothing written past the row
    assert (got - want).abs().max().item() < 5e-6
    # rows that cannot be bulk-copied (pitch not a multiple of 4 floats) take the FMA kernel: same result
    if H % 2 == 1 and ndir == 2:
        xg2 = torch.empty(B * L, pitch + 1, device="cuda")
        xg2[:, :pitch] = xg
        got2 = torch.zeros_like(want)
        call("ruart_lstm_recurrence", ptr(xg2), pitch + 1, ptr(whh), ptr(got2), ndir * H + 3, B, L, H, ndir, st)
        assert torch.equal(got2, want)


@pytest.mark.parametrize("B,L,H,ndir", [(7, 9, 125, 2), (33, 12, 128, 2), (1, 1, 125, 2), (9, 2, 17, 1), (20, 3, 64, 2)])
def test_lstm_bptt_tensor_core_kernel_matches_torch_autograd(B, L, H, ndir):
    """ruart_lstm_recurrence_backward (mma.sync on bf16 hi|lo splits of W_hh^T and the gate gradients) on the gates
    saved by ruart_lstm_recurrence_train, against torch autograd through a plain restatement of the recurrence."""
    from ruart_b200._lib import current_stream, ptr
    from ruart_b200.ops import call
    g = torch.Generator(device="cuda").manual_seed(B * 100 + L)
    xg = (torch.randn(B, L, ndir * 4 * H, device="cuda", generator=g) * 0.8).requires_grad_(True)
    whh = (torch.rand(ndir, 4 * H, H, device="cuda", generator=g) * 2 - 1) / H ** 0.5
    dout = torch.randn(B, L, ndir * H, device="cuda", generator=g)
    outs = []
    for d in range(ndir):
        h = torch.zeros(B, H, device="cuda")
        c = torch.zeros(B, H, device="cuda")
        hs = [None] * L
        for s in range(L):
            t = s if d == 0 else L - 1 - s
            pre = xg[:, t, d * 4 * H:(d + 1) * 4 * H] + h @ whh[d].t()
            i, f, gg, o = pre.split(H, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            hs[t] = h
        outs.append(torch.stack(hs, 1))
    ref_out = torch.cat(outs, 2)
    want = torch.autograd.grad(ref_out, xg, dout)[0].reshape(B * L, ndir * 4 * H)
    st = current_stream()
    out = torch.empty(B, L, ndir * H, device="cuda")
    gates = torch.empty(B * L, ndir * 5 * H, device="cuda")
    xg2 = xg.detach().reshape(B * L, -1).contiguous()
    call("ruart_lstm_recurrence_train", ptr(xg2), ndir * 4 * H, ptr(whh), ptr(out), ndir * H, B, L, H, ndir,
         ptr(gates), ndir * 5 * H, st)
    assert (out - ref_out).abs().max().item() < 2e-5
    got = torch.full((B * L, ndir * 4 * H + 2), 7.0, device="cuda")
    call("ruart_lstm_recurrence_backward", ptr(gates), ndir * 5 * H, ptr(whh), ptr(dout), ndir * H, ptr(got),
         ndir * 4 * H + 2, B, L, H, ndir, st)
    torch.cuda.synchronize()
    assert (got[:, ndir * 4 * H:] == 7.0).all()             # nothing written past the row
    err = (got[:, :ndir * 4 * H] - want).abs().max().item()
    assert err < 2e-5 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("rows,widths,parts", [(37, (300, 250, 250), 3), (1000, (768, 300, 12, 8, 300), 3),
                                               (1, (2, 6, 10), 2), (513, (250,) * 7, 3), (5, (1000,) * 8, 2),
                                               (9, (1250,) * 8, 3), (130, (126, 2), 1)])
def test_split_concat_operand_is_the_exact_bf16_split_of_the_concatenation(rows, widths, parts):
    """ruart_split_concat_bf16 (row-tiled kernel; flat kernel beyond Kp = 8192) on pitched sources, 8-column groups
    that straddle two sources, a ragged last tile: every part bit-identical to the torch restatement."""
    from ruart_b200 import sdnet_ops as K
    g = torch.Generator(device="cuda").manual_seed(rows)
    pieces = []
    for i, w in enumerate(widths):
        full = torch.randn(rows, w + 2 * (i % 3), device="cuda", generator=g) * (10.0 ** (i % 4 - 2))
        pieces.append(full[:, :w])                                  # pitched view, even pitch
    got, Kp = K.split_concat(pieces, parts)
    Ksum = sum(widths)
    assert Kp == (Ksum + 63) // 64 * 64 and got.shape == (rows, parts * Kp)
    x = torch.cat(pieces, 1)
    for p in range(parts):
        h = x.to(torch.bfloat16)
        part = got[:, p * Kp:(p + 1) * Kp]
        assert torch.equal(part[:, :Ksum], h), "part %d" % p
        assert (part[:, Ksum:] == 0).all()
        x = x - h.float()


def test_attention_module_matches_oracle():
    from ruart_b200.Models import Layers
    Layers.set_dropout_prob(0.0)
    Layers.set_sdnet_precision(3)
    torch.manual_seed(3)
    att = Layers.Attention(70, 33, correlation_func=3).cuda().eval()
    x1 = torch.randn(6, 19, 70, device="cuda")
    x2 = torch.randn(6, 11, 70, device="cuda")
    x3 = torch.randn(6, 11, 45, device="cuda")
    mask = torch.ones(6, 11, dtype=torch.uint8, device="cuda")
    mask[0, 5:] = 0
    mask[3, 1:] = 0
    with torch.no_grad():
        got = att(x1, x2, mask, x3=x3)
    sd = {"a." + k: v.detach().cpu() for k, v in att.state_dict().items()}
    want = sdnet_oracle.attention(sd, "a", x1.cpu(), x2.cpu(), mask.cpu(), x3.cpu())
    assert (got.cpu() - want).abs().max().item() < 1e-4


def test_whole_layernorm_on_strided_rows():
    from ruart_b200 import sdnet_ops as K
    buf = torch.randn(7, 11, 40, device="cuda") * 3 + 1.5
    view = buf[:, :, 5:30]
    want = view.clone()
    m = want.mean()
    v = (want - m).pow(2).mean()
    want = (want - m) / torch.sqrt(v + 1e-5)
    keep = buf.clone()
    K.whole_layernorm_(view)
    assert (view - want).abs().max().item() < 1e-5
    assert torch.equal(buf[:, :, :5], keep[:, :, :5]) and torch.equal(buf[:, :, 30:], keep[:, :, 30:])


@pytest.mark.parametrize("B,L1,L2,Hd,D3", [(3, 100, 100, 250, 250), (2, 37, 40, 250, 250), (4, 70, 128, 300, 300),
                                           (2, 5, 3, 125, 17), (1, 129, 9, 64, 500)])
def test_attention_tail_tensor_core_form_matches_fp32_form(B, L1, L2, Hd, D3):
    # split_parts = 2: mma.sync with hi+lo bf16 operands (3 terms) vs the fp32 CUDA-core kernel,
    # on strided inputs / outputs like the model's column views
    from ruart_b200 import sdnet_ops as K
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + L1 + L2)
    p1 = torch.relu(torch.randn(B * L1, Hd + 6, device="cuda", generator=g))[:, 2:2 + Hd] * 0.3
    p2 = torch.relu(torch.randn(B * L2, Hd, device="cuda", generator=g)) * 0.3
    x3 = torch.randn(B, L2, D3 + 2, device="cuda", generator=g)[:, :, 1:1 + D3]
    mask = (torch.rand(B, L2, device="cuda", generator=g) > 0.3)
    mask[:, 0] = True
    m8 = K.as_u8(mask)
    outs = []
    for parts in (3, 2):
        buf = torch.full((B, L1, D3 + 2), 0.5, device="cuda")
        K.attention_tail(p1, p2, m8, x3, buf[:, :, 2:], B, L1, L2, add=False, parts=parts)
        K.attention_tail(p1, p2, m8, x3, buf[:, :, 2:], B, L1, L2, add=True, parts=parts)   # accumulate form
        outs.append(buf)
    torch.cuda.synchronize()
    want = torch.softmax((p1.reshape(B, L1, Hd) @ p2.reshape(B, L2, Hd).transpose(1, 2)).masked_fill(~mask[:, None], float("-inf")), -1) @ x3
    scale = want.abs().max().item()
    assert (outs[0][:, :, 2:] - 2 * want).abs().max().item() < 1e-4 * scale
    assert (outs[1][:, :, 2:] - 2 * want).abs().max().item() < 1e-4 * scale
    assert (outs[1][:, :, :2] == 0.5).all()
