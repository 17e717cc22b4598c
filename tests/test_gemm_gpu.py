"""GPU: tcgen05 GEMM against torch fp32 matmul of the same (bf16-rounded) operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, epi, scale):
    c = a.float() @ w.float().t()
    if epi in (1, 2, 4):
        c = c + bias
    if epi == 2:
        c = torch.nn.functional.gelu(c)
    if epi == 3:
        c = torch.relu(c) * scale
    if epi == 4:
        c = torch.relu(c)
    return c


@pytest.mark.parametrize("M,N,K,epi", [
    (128, 256, 64, 0), (128, 256, 768, 0), (256, 768, 768, 1), (1000, 2304, 768, 1),
    (333, 3072, 768, 2), (517, 768, 3072, 1), (77, 1000, 320, 1), (4096, 1200, 1408, 1),
    (130, 250, 832, 3), (64, 125, 256, 3), (1, 32, 64, 0), (5000, 768, 768, 1),
    (40000, 2304, 768, 1),
])
def test_gemm_bf16(M, N, K, epi):
    from ruart_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    out32 = torch.full((M, N), float("nan"), device="cuda")
    out16 = torch.zeros((M, N + 8), device="cuda", dtype=torch.bfloat16)[:, :N]
    ops.gemm(a, w, M, N, K, epi=epi, bias=bias if epi in (1, 2, 4) else None,
             scale=scale if epi == 3 else None, out_f32=out32, out_bf16=out16)
    torch.cuda.synchronize()
    want = _ref(a, w, bias, epi, scale)
    err = (out32 - want).abs().max().item()
    tol = 2e-3 * max(1.0, want.abs().max().item())
    assert err < tol, (err, tol)
    err16 = (out16.float() - want).abs().max().item()
    assert err16 < 1e-2 * max(1.0, want.abs().max().item())
    # bf16-only output: the TMA-store epilogue when the row pitch allows it (N % 8 == 0)
    pitch = N + 8 if N % 8 == 0 else N
    buf = torch.full((M + 1, pitch), 7.0, device="cuda", dtype=torch.bfloat16)
    out_t = buf[:M, :N]
    ops.gemm(a, w, M, N, K, epi=epi, bias=bias if epi in (1, 2, 4) else None,
             scale=scale if epi == 3 else None, out_bf16=out_t)
    torch.cuda.synchronize()
    err_t = (out_t.float() - want).abs().max().item()
    assert err_t < 1e-2 * max(1.0, want.abs().max().item()), err_t
    assert (buf[M] == 7.0).all() and (buf[:, N:] == 7.0).all()  # nothing written out of bounds


def test_gemm_scalar_scale_and_fast_gelu():
    from ruart_b200 import ops
    M, N, K = 300, 300, 320
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    s = torch.tensor([0.0577], device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, M, N, K, epi=3, scale=s, out_f32=out)
    want = torch.relu(a.float() @ w.float().t()) * s
    assert (out - want).abs().max().item() < 1e-3
    bias = torch.randn(N, device="cuda")
    want = torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
    for mode, tol in ((1, 2e-5), (2, 1e-4)):   # A&S 7.1.26 on rcp/ex2 ; tanh form fitted to erf
        ops.gemm(a, w, M, N, K, epi=2, bias=bias, out_f32=out, fast_gelu=mode)
        assert (out - want).abs().max().item() < tol


@pytest.mark.parametrize("terms,tol", [(3, 3e-5), (6, 2e-6)])
def test_gemm_split_terms_reach_fp32_accuracy(terms, tol):
    from ruart_b200 import ops
    M, N, K = 513, 384, 300
    Kp = 320
    parts = 2 if terms == 3 else 3
    a = torch.randn(M, K, device="cuda", dtype=torch.float64)
    w = torch.randn(N, K, device="cuda", dtype=torch.float64) * 0.1

    def split(x):
        x = x.float()
        buf = torch.zeros(x.shape[0], parts * Kp, device="cuda", dtype=torch.bfloat16)
        r = x.clone()
        for p in range(parts):
            h = r.bfloat16()
            buf[:, p * Kp:p * Kp + K] = h
            r = r - h.float()
        return buf

    out = torch.empty(M, N, device="cuda")
    ops.gemm(split(a), split(w), M, N, Kp, a_parts=parts, w_parts=parts, n_terms=terms, out_f32=out)
    want = a.float().double() @ w.float().double().t()
    rel = ((out.double() - want).abs().max() / want.abs().max()).item()
    assert rel < tol, rel


def test_gemm_split_output_parts():
    from ruart_b200 import ops
    M, N, K = 200, 128, 128
    a = (torch.randn(M, K, device="cuda")).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    out = torch.zeros(M, 3 * 192, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, M, N, K, out_bf16=out, out_parts=3, out_part_stride=192)
    want = a.float() @ w.float().t()
    recon = out[:, 0:N].float() + out[:, 192:192 + N].float() + out[:, 384:384 + N].float()
    assert (recon - want).abs().max().item() < 2e-5 * want.abs().max().item() + 1e-6
    assert out[:, N:192].abs().max().item() == 0


def test_gemm_fused_residual():
    from ruart_b200 import ops
    M, N, K = 777, 768, 3072
    a = (torch.randn(M, K, device="cuda") * 0.3).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, M, N, K, epi=1, bias=bias, out_bf16=out, residual=res)
    want = a.float() @ w.float().t() + bias + res.float()
    assert (out.float() - want).abs().max().item() < 1e-2 * want.abs().max().item()


@pytest.mark.parametrize("M,N,K,kind", [(4100, 768, 768, "bias"), (2304, 2304, 768, "bias"), (4100, 3072, 768, "gelu"),
                                        (4100, 768, 3072, "res"), (113664, 768, 3072, "res")])
def test_gemm_cta_pair_path(M, N, K, kind):
    # plain bf16 output, N % 256 == 0, M >= 2048: the cta_group::2 kernel (256 x 256 pair tiles);
    # M = 4100 leaves the second CTA of the last pair entirely out of bounds
    from ruart_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.3).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if kind == "res" else None
    buf = torch.full((M + 1, N + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    out = buf[:M, :N]
    ops.gemm(a, w, M, N, K, epi=2 if kind == "gelu" else 1, bias=bias, out_bf16=out, fast_gelu=2, residual=res)
    torch.cuda.synchronize()
    rows = torch.cat([torch.arange(0, min(M, 700)), torch.arange(max(0, M - 700), M)]).cuda()  # both ends
    want = a[rows].float() @ w.float().t() + bias
    if kind == "gelu":
        want = torch.nn.functional.gelu(want)
    if kind == "res":
        want = want + res[rows].float()
    assert (out[rows].float() - want).abs().max().item() < 1e-2 * max(1.0, want.abs().max().item())
    assert (buf[M] == 7.0).all() and (buf[:, N:] == 7.0).all()
    assert torch.isfinite(out.float()).all()


def _partials(x, slots=6):
    """[M, H] fp32 -> [M, 8, 2] partial (sum, sum of squares) spread over `slots` column groups."""
    M, H = x.shape
    st = torch.zeros(M, 8, 2, device=x.device)
    for k, chunk in enumerate(x.split(H // slots, dim=1)):
        st[:, k, 0] = chunk.sum(1)
        st[:, k, 1] = (chunk * chunk).sum(1)
    return st.contiguous()


def _ln(x, g, b, eps=1e-12):
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * g + b


@pytest.mark.parametrize("M", [2048, 5000, 20011])
def test_gemm_folded_layernorm_forms(M):
    """ruart_gemm_bf16_fold (the CTA-pair GEMM with BertLayerNorm folded in, modeling.py:155-168,260-264,287,299-303)
    against LayerNorm + matmul in torch fp32 on the same bf16-rounded operands."""
    from ruart_b200._lib import call, current_stream, ptr
    H, eps = 768, 1e-12
    g = torch.Generator(device="cuda").manual_seed(M)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    gamma, beta = rnd(H) * 0.5 + 1.0, rnd(H) * 0.3
    # ---- fold 2: out = A W^T + LayerNorm(residual) + b (pre-LayerNorm rows) + partial sums of out
    for K in (768, 3072):
        a = (rnd(M, K) * 0.5).bfloat16()
        w = (rnd(H, K) * 0.03).bfloat16()
        bias = rnd(H)
        raw = (rnd(M, H) * 2.0 + 0.7).bfloat16()                     # stored rows whose LayerNorm is pending
        st_in = _partials(raw.float())
        out = torch.full((M + 1, H), 7.0, device="cuda", dtype=torch.bfloat16)
        st_out = torch.zeros(M + 1, 8, 2, device="cuda")
        call("ruart_gemm_bf16_fold", ptr(a), K, ptr(w), K, M, H, K, 2, 1, ptr((beta + bias).contiguous()), ptr(gamma),
             ptr(st_in), eps, ptr(out), H, ptr(raw), H, ptr(st_out), current_stream())
        torch.cuda.synchronize()
        want = a.float() @ w.float().t() + _ln(raw.float(), gamma, beta, eps) + bias
        scale = want.abs().max().item()
        assert (out[:M].float() - want).abs().max().item() < 1e-2 * scale
        assert (out[M] == 7.0).all() and (st_out[M] == 0).all()      # nothing past the last row
        s1, s2 = st_out[:M, :6, 0].sum(1), st_out[:M, :6, 1].sum(1)
        assert (s1 - want.sum(1)).abs().max().item() < 2e-3 * scale * H ** 0.5
        assert ((s2 - (want * want).sum(1)).abs() / (want * want).sum(1)).max().item() < 1e-3
    # ---- fold 1: out = act(LayerNorm(A) W0^T + b) from the stored rows A, W0 * gamma, colsum and W0 beta + b
    for N, epi in ((2304, 1), (3072, 2)):
        raw = (rnd(M, H) * 1.5 - 0.4).bfloat16()
        w0 = rnd(N, H) * 0.03
        bias = rnd(N)
        wf = (w0 * gamma[None, :]).bfloat16().contiguous()
        colsum = wf.float().sum(1).contiguous()
        cvec = (w0 @ beta + bias).contiguous()
        st_in = _partials(raw.float(), slots=3)
        out = torch.full((M + 1, N), 7.0, device="cuda", dtype=torch.bfloat16)
        call("ruart_gemm_bf16_fold", ptr(raw), H, ptr(wf), H, M, N, H, 1, epi, ptr(cvec), ptr(colsum), ptr(st_in), eps,
             ptr(out), N, None, 0, None, current_stream())
        torch.cuda.synchronize()
        # the same algebra in fp32 on the same rounded operands (tight), and the unfolded definition (bf16-level)
        x = raw.float()
        u = x.mean(-1, keepdim=True)
        r = torch.rsqrt((x * x).mean(-1, keepdim=True) - u * u + eps)
        pre = r * (x @ wf.float().t()) - (r * u) * colsum + cvec
        ref = _ln(x, gamma, beta, eps) @ w0.t() + bias
        if epi == 2:
            pre, ref = torch.nn.functional.gelu(pre), torch.nn.functional.gelu(ref)
        scale = ref.abs().max().item()
        assert (out[:M].float() - pre).abs().max().item() < 6e-3 * scale
        assert (out[:M].float() - ref).abs().max().item() < 2.5e-2 * scale
        assert (out[M] == 7.0).all()
