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
