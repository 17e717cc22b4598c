"""Differentiable building blocks of the SDNet stack (the training step
`SDNetTrainer.update`, reference Models/SDNetTrainer.py:330-376).

The reference differentiates its forward with torch autograd over torch ops.  Here every op is a
`torch.autograd.Function` whose forward AND backward are calls into libruart_b200.so
(csrc/backward_kernels.cu, the tcgen05 GEMM with transposed split operands for dgrad / wgrad, the
persistent LSTM with saved gates + a BPTT kernel): torch only allocates memory and chains the
Functions.  The same ops serve the `Layers.py`-level API (`AttentionScore.forward`,
`LinearSelfAttn.forward`, `BilinearSeqAttn.forward`, `weighted_avg`, ...), which the fused inference
path does not go through.

All tensors are fp32 CUDA.  `parts` is the split width of the tensor-core GEMM operands (3 = fp32
grade, the default for training: gradients within 1e-4 of the fp32 reference).
"""
import torch
from torch.autograd import Function

from . import _lib, ops
from . import sdnet_ops as K
from ._lib import current_stream, ptr
from .ops import call

_TERMS = {1: 1, 2: 3, 3: 6}
_ws_cache = {}


def _ws(dev, n_doubles):
    """fp64 workspace per (device, stream), grown on demand."""
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    w = _ws_cache.get(key)
    if w is None or w.numel() < n_doubles:
        w = torch.empty(max(n_doubles, 1 << 16), dtype=torch.float64, device=dev)
        _ws_cache[key] = w
    return w


def _need(*ts):
    for t in ts:
        if t is not None and torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError("ruart_b200 ops run on CUDA tensors only; there is no CPU fallback")


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _f32(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


# ------------------------------------------------------------------------------------ raw helpers
def split_t(x2d, parts):
    """fp32 [rows, K] -> transposed bf16 split operand [K, parts*rows_p]; returns (tensor, rows_p)."""
    rows, Kc, pitch = K.rows2d(x2d)
    rows_p = ops.round_up(max(rows, 1), 64)
    out = torch.empty((Kc, parts * rows_p), dtype=torch.bfloat16, device=x2d.device)
    call("ruart_split_bf16_t", ptr(x2d), pitch, rows, Kc, rows_p, parts, ptr(out), current_stream())
    return out, rows_p


def eltwise(op, a, b=None, v=None, out=None):
    rows, cols, ap = K.rows2d(a)
    if out is None:
        out = _f32(a.shape, a.device)
    bp = K.rows2d(b)[2] if b is not None else 0
    call("ruart_eltwise", op, ptr(a), ap, ptr(b), bp, ptr(v), 0 if v is None else v.numel(), ptr(out),
         K.rows2d(out)[2], rows, cols, current_stream())
    return out


def colsum(x2d, out=None):
    rows, cols, pitch = K.rows2d(x2d)
    if out is None:
        out = _f32((cols,), x2d.device)
    call("ruart_colsum", ptr(x2d), pitch, rows, cols, ptr(_ws(x2d.device, 64 * cols)), ptr(out), 0, current_stream())
    return out


def bmm_raw(a, b, trans_a, trans_b, M, N, Kd, out, alpha=1.0, accumulate=False):
    """out[batch] (+)= alpha * op(a[batch]) op(b[batch]) for 3-D tensors with unit inner stride."""
    batch = out.shape[0]
    assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
    sa = a.stride(0) if a.shape[0] > 1 else 0
    sb = b.stride(0) if b.shape[0] > 1 else 0
    call("ruart_bmm_f32", ptr(a), a.stride(1), sa, int(trans_a), ptr(b), b.stride(1), sb, int(trans_b), ptr(out),
         out.stride(1), out.stride(0), batch, M, N, Kd, float(alpha), int(accumulate), current_stream())
    return out


def _gemm(a_split, Kp, w_split, rows, N, parts, out, epi=ops.EPI_NONE, bias=None, scale=None):
    return K.linear(a_split, Kp, w_split, rows, N, parts, out, epi=epi, bias=bias, scale=scale)


# ------------------------------------------------------------------------------------ Linear
class LinearFn(Function):
    """y = act(x W^T + b) over the last dim (nn.Linear; act = ReLU for AttentionScore, Layers.py:226-228).
    Forward, dgrad (dX = dY W) and wgrad (dW = dY^T X) all run on the tcgen05 GEMM; outputs narrower
    than 16 columns (ques_merger 250->1, noanswer_w 500->1) use the fp32 batched-GEMM kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, parts):
        _need(x, weight, bias)
        x2 = _c(x.reshape(-1, x.shape[-1]))
        rows, Kin = x2.shape
        N = weight.shape[0]
        w = _c(weight.detach())
        out = _f32((rows, N), x.device)
        small = N < 16
        if small:
            assert not relu
            bmm_raw(x2.unsqueeze(0), w.unsqueeze(0), False, True, rows, N, Kin, out.unsqueeze(0))
            if bias is not None:
                eltwise(5, out, v=_c(bias.detach()), out=out)
        else:
            a, Kp = K.split_act(x2, parts)
            ws, _ = K.split_act(w, parts)
            if relu:
                assert bias is None
                _gemm(a, Kp, ws, rows, N, parts, out, epi=ops.EPI_RELU_SCALE, scale=K.ones(x.device))
            elif bias is not None:
                _gemm(a, Kp, ws, rows, N, parts, out, epi=ops.EPI_BIAS, bias=_c(bias.detach()))
            else:
                _gemm(a, Kp, ws, rows, N, parts, out)
        ctx.save_for_backward(x2, w, out if relu else None)
        ctx.meta = (bool(relu), int(parts), bias is not None, tuple(x.shape), small)
        return out.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, out = ctx.saved_tensors
        relu, parts, has_bias, xshape, small = ctx.meta
        rows, Kin = x2.shape
        N = w.shape[0]
        dz = _c(dy.reshape(rows, N))
        if relu:
            dz = eltwise(1, dz, out)
        dx = dw = db = None
        if small:
            if ctx.needs_input_grad[0]:
                dx = _f32((rows, Kin), dz.device)
                bmm_raw(dz.unsqueeze(0), w.unsqueeze(0), False, False, rows, Kin, N, dx.unsqueeze(0))
            if ctx.needs_input_grad[1]:
                dw = _f32((N, Kin), dz.device)
                bmm_raw(dz.unsqueeze(0), x2.unsqueeze(0), True, False, N, Kin, rows, dw.unsqueeze(0))
        else:
            if ctx.needs_input_grad[0]:
                a, Np = K.split_act(dz, parts)
                wt, _ = split_t(w, parts)                      # [Kin, parts*Np]
                dx = _f32((rows, Kin), dz.device)
                _gemm(a, Np, wt, rows, Kin, parts, dx)
            if ctx.needs_input_grad[1]:
                at, Rp = split_t(dz, parts)                    # [N, parts*Rp]
                xt, _ = split_t(x2, parts)                     # [Kin, parts*Rp]
                dw = _f32((N, Kin), dz.device)
                _gemm(at, Rp, xt, N, Kin, parts, dw)
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum(dz)
        return (dx.view(xshape) if dx is not None else None), dw, db, None, None


def linear(x, weight, bias=None, relu=False, parts=3):
    return LinearFn.apply(x, weight, bias, relu, parts)


# ------------------------------------------------------------------------------------ batched matmul
class BmmFn(Function):
    """C = A op(B) per batch element, fp32 (x1_rep.bmm(x2_rep^T), alpha.bmm(x3): Layers.py:237,288)."""

    @staticmethod
    def forward(ctx, a, b, trans_b):
        _need(a, b)
        a, b = _c(a), _c(b)
        Bn, M, Kd = a.shape
        N = b.shape[1] if trans_b else b.shape[2]
        assert (b.shape[2] if trans_b else b.shape[1]) == Kd and b.shape[0] == Bn
        out = _f32((Bn, M, N), a.device)
        bmm_raw(a, b, False, trans_b, M, N, Kd, out)
        ctx.save_for_backward(a, b)
        ctx.trans_b = bool(trans_b)
        return out

    @staticmethod
    def backward(ctx, dc):
        a, b = ctx.saved_tensors
        dc = _c(dc)
        Bn, M, Kd = a.shape
        N = dc.shape[2]
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _f32(a.shape, a.device)
            # C = A B^T: dA = dC B ;  C = A B: dA = dC B^T
            bmm_raw(dc, b, False, not ctx.trans_b, M, Kd, N, da)
        if ctx.needs_input_grad[1]:
            db = _f32(b.shape, b.device)
            if ctx.trans_b:      # dB = dC^T A   [N, K]
                bmm_raw(dc, a, True, False, N, Kd, M, db)
            else:                # dB = A^T dC   [K, N]
                bmm_raw(a, dc, True, False, Kd, N, M, db)
        return da, db, None


def bmm(a, b, trans_b=False):
    return BmmFn.apply(a, b, trans_b)


# ------------------------------------------------------------------------------------ softmax
class MaskedSoftmaxFn(Function):
    """softmax over the last dim of x [B, L1, L2] with keys mask[b, j] == 0 at -inf (Layers.py:283-288)."""

    @staticmethod
    def forward(ctx, x, mask_u8):
        _need(x, mask_u8)
        x = _c(x)
        Bn, L1, L2 = x.shape
        out = _f32(x.shape, x.device)
        call("ruart_masked_softmax", ptr(x), L2, ptr(mask_u8), Bn, L1, L2, ptr(out), L2, current_stream())
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        dp = _c(dp)
        Bn, L1, L2 = p.shape
        dx = _f32(p.shape, p.device)
        call("ruart_softmax_backward", ptr(p), L2, ptr(dp), L2, ptr(dx), L2, Bn * L1, L2, current_stream())
        return dx, None


def masked_softmax(x, mask):
    """x [B, L1, L2] (or [B, L2]); mask [B, L2] bool/uint8 or None."""
    m = None if mask is None else K.as_u8(mask)
    if x.dim() == 2:
        return MaskedSoftmaxFn.apply(x.unsqueeze(1), m).squeeze(1)
    return MaskedSoftmaxFn.apply(x, m)


class MaskFillFn(Function):
    """x with masked positions set to -inf — the reference's `scores.data.masked_fill_(empty_mask, -inf)`
    (Layers.py:283-284,339,426,463): done on `.data`, so the gradient is the identity."""

    @staticmethod
    def forward(ctx, x, mask_u8):
        x = _c(x)
        out = _f32(x.shape, x.device)
        rows, cols, _ = K.rows2d(x)
        call("ruart_mask_fill", ptr(x), ptr(mask_u8), rows * cols, float("-inf"), ptr(out), current_stream())
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def mask_fill_neg_inf(x, mask):
    m = K.as_u8(mask)
    assert m.shape == x.shape
    return MaskFillFn.apply(x, m)


class MulFn(Function):
    """y = x * m with a constant m of the same shape (dropout masks, Layers.py:23-39)."""

    @staticmethod
    def forward(ctx, x, m):
        x, m = _c(x), _c(m)
        ctx.save_for_backward(m)
        return eltwise(0, x.view(-1, x.shape[-1]), m.view(-1, x.shape[-1])).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (m,) = ctx.saved_tensors
        dy = _c(dy)
        return eltwise(0, dy.view(-1, dy.shape[-1]), m.view(-1, dy.shape[-1])).view(dy.shape), None


def mul_const(x, m):
    return MulFn.apply(x, m)


class AddFn(Function):
    """a + b (x_od_ocr += pos_att, SDNet.py:399-401)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        return eltwise(2, a.view(-1, a.shape[-1]), b.view(-1, a.shape[-1])).view(a.shape)

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return AddFn.apply(a, b)


class LayerMixFn(Function):
    """SDNet.linear_sum (SDNet.py:573-583) on materialised per-layer tensors: sum_l out_l * softmax(alpha)_l * gamma.
    (The fused path never materialises the layers; this is the `Layers.py`/`SDNet.py`-level API form.)"""

    @staticmethod
    def forward(ctx, alpha, gamma, *layers):
        stack = torch.stack([_c(t) for t in layers], 0)              # [NL, ..., D]
        NL = stack.shape[0]
        flat = stack.view(NL, -1)
        a = torch.empty(NL, dtype=torch.float32, device=stack.device)
        call("ruart_masked_softmax", ptr(_c(alpha.detach().float())), NL, None, 1, 1, NL, ptr(a), NL, current_stream())
        coef = eltwise(3, a.view(1, NL), v=_c(gamma.detach().float().reshape(-1)))      # a_l * gamma
        out = _f32((1, flat.shape[1]), stack.device)
        bmm_raw(coef.view(1, 1, NL), flat.unsqueeze(0), False, False, 1, flat.shape[1], NL, out.unsqueeze(0))
        ctx.save_for_backward(a, _c(gamma.detach().float().reshape(-1)), flat)
        ctx.shapes = (tuple(alpha.shape), tuple(gamma.shape), tuple(layers[0].shape))
        return out.view(layers[0].shape)

    @staticmethod
    def backward(ctx, dy):
        a, g, flat = ctx.saved_tensors
        NL = a.numel()
        dy = _c(dy).view(1, -1)
        # s_l = <dy, layer_l>
        s = _f32((1, NL), dy.device)
        bmm_raw(dy.unsqueeze(0), flat.unsqueeze(0), False, True, 1, NL, flat.shape[1], s.unsqueeze(0))
        # d gamma = sum_l a_l s_l ; d alpha = softmax backward of (gamma * s)
        dgamma = _f32((1, 1), dy.device)
        bmm_raw(a.view(1, 1, NL), s.view(1, NL, 1), False, False, 1, 1, NL, dgamma.unsqueeze(0))
        da = eltwise(3, s, v=g)
        dalpha = _f32((1, NL), dy.device)
        call("ruart_softmax_backward", ptr(a), NL, ptr(da), NL, ptr(dalpha), NL, 1, NL, current_stream())
        dlayers = []
        coef = eltwise(3, a.view(1, NL), v=g)
        for l in range(NL):
            if ctx.needs_input_grad[2 + l]:
                dlayers.append(eltwise(3, dy, v=coef.view(-1)[l:l + 1]).view(ctx.shapes[2]))
            else:
                dlayers.append(None)
        return (dalpha.view(ctx.shapes[0]), dgamma.view(ctx.shapes[1])) + tuple(dlayers)


def layer_mix(layers, alpha, gamma):
    return LayerMixFn.apply(alpha, gamma, *layers)


# ------------------------------------------------------------------------------------ diagonal scale
class ScaleColsFn(Function):
    """y = x * d over the last dim; d has 1 element (constant 1/sqrt(h)) or one per column
    (`x1_rep * self.diagonal.expand_as(x1_rep)`, Layers.py:229)."""

    @staticmethod
    def forward(ctx, x, d):
        x = _c(x)
        dv = _c(d.detach().reshape(-1))
        out = eltwise(3, x.view(-1, x.shape[-1]), v=dv).view(x.shape)
        ctx.save_for_backward(x, dv)
        ctx.dshape = tuple(d.shape)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, dv = ctx.saved_tensors
        dy2 = _c(dy).view(-1, x.shape[-1])
        dx = dd = None
        if ctx.needs_input_grad[0]:
            dx = eltwise(3, dy2, v=dv).view(x.shape)
        if ctx.needs_input_grad[1]:
            prod = eltwise(0, dy2, x.view(-1, x.shape[-1]))
            s = colsum(prod)
            dd = (s if dv.numel() > 1 else colsum(s.view(-1, 1))).view(ctx.dshape)
        return dx, dd


def scale_cols(x, d):
    return ScaleColsFn.apply(x, d)


# ------------------------------------------------------------------------------------ whole-tensor LN
class WholeLayerNormFn(Function):
    """F.layer_norm(x, x.size()), eps 1e-5, no affine (Layers.py:167-168)."""

    @staticmethod
    def forward(ctx, x, eps):
        _need(x)
        y = x.contiguous().clone()
        rows, cols, pitch = K.rows2d(y)
        stats = _f32((2,), x.device)
        call("ruart_whole_layernorm_stats", ptr(y), rows, cols, pitch, float(eps), ptr(K.ln_workspace(x.device)),
             ptr(stats), current_stream())
        ctx.save_for_backward(y, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, stats = ctx.saved_tensors
        dy = _c(dy)
        rows, cols, pitch = K.rows2d(y)
        dx = _f32(y.shape, y.device)
        call("ruart_whole_layernorm_backward", ptr(y), pitch, ptr(dy), pitch, rows, cols, ptr(stats),
             ptr(_ws(y.device, 2048)), ptr(dx), pitch, current_stream())
        return dx, None


def whole_layernorm(x, eps=1e-5):
    return WholeLayerNormFn.apply(x, eps)


# ------------------------------------------------------------------------------------ embeddings
class EmbeddingFn(Function):
    """weight[ids] (nn.Embedding, SDNet.py:447-492); the weight gradient is summed per vocabulary row in
    ascending position order by one warp per row (deterministic)."""

    @staticmethod
    def forward(ctx, ids, weight):
        _need(ids, weight)
        flat = _c(ids.reshape(-1))
        w = weight.detach()
        D = w.shape[1]
        out = _f32((flat.numel(), D), w.device)
        K.gather_rows(w, flat, out, None, flat.numel(), D)
        ctx.save_for_backward(flat)
        ctx.wshape = tuple(w.shape)
        return out.view(*ids.shape, D)

    @staticmethod
    def backward(ctx, dy):
        (flat,) = ctx.saved_tensors
        V, D = ctx.wshape
        dy2 = _c(dy).view(-1, D)
        dw = _f32((V, D), dy.device)
        n = flat.numel()
        ws_bytes = int(_lib.lib().ruart_embedding_grad_workspace_bytes(n, V, D))      # the sorted form
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy.device)
        call("ruart_embedding_grad", ptr(flat), 1 if flat.dtype == torch.int64 else 0, n, ptr(dy2), D, D, V,
             ptr(ws), ws_bytes, ptr(dw), D, 0, current_stream())
        return None, dw


def embedding(ids, weight):
    if not weight.requires_grad or not torch.is_grad_enabled():
        w = weight.detach()
        flat = _c(ids.reshape(-1))
        out = _f32((flat.numel(), w.shape[1]), w.device)
        K.gather_rows(w, flat, out, None, flat.numel(), w.shape[1])
        return out.view(*ids.shape, w.shape[1])
    return EmbeddingFn.apply(ids, weight)


class PermuteRowsFn(Function):
    """out = zeros [n_out, D]; out[dst_idx[k]] = src[src_idx[k]] for k < n with UNIQUE indices on both
    sides (pre-align pack / unpack, SDNet.py:504-520,540-550): the backward is the inverse gather."""

    @staticmethod
    def forward(ctx, src, src_idx, dst_idx, n_out):
        src = _c(src)
        D = src.shape[-1]
        s2 = src.view(-1, D)
        out = torch.zeros((n_out, D), dtype=torch.float32, device=src.device)
        n = (src_idx if src_idx is not None else dst_idx).numel()
        K.gather_rows(s2, src_idx, out, dst_idx, n, D)
        ctx.save_for_backward(src_idx, dst_idx)
        ctx.meta = (tuple(src.shape), n, D)
        return out

    @staticmethod
    def backward(ctx, dy):
        src_idx, dst_idx = ctx.saved_tensors
        shape, n, D = ctx.meta
        dy = _c(dy)
        dsrc = torch.zeros(shape, dtype=torch.float32, device=dy.device)
        K.gather_rows(dy, dst_idx, dsrc.view(-1, D), src_idx, n, D)
        return dsrc, None, None, None


def permute_rows(src, src_idx, dst_idx, n_out):
    return PermuteRowsFn.apply(src, src_idx, dst_idx, n_out)


# ------------------------------------------------------------------------------------ BERT layer mix
class SubwordMixFn(Function):
    """out[item, j] = gamma * sum_l softmax(alpha)_l * mean_subwords(h_l)  (Bert.py:149-165 + SDNet.py:573-583)
    over the kept hidden states of the locked encoder; gradients for alpha and gamma only."""

    @staticmethod
    def forward(ctx, alpha, gamma, pack):
        hf, hb, layer_stride, words, n_words, row_start, wmask, N, W, NL, H = pack
        a = _c(alpha.detach().float())
        g = _c(gamma.detach().float().reshape(-1))
        out = torch.zeros((N, W, H), dtype=torch.float32, device=a.device)
        call("ruart_subword_avg_layers", ptr(hf), ptr(hb), layer_stride, ptr(words), n_words, ptr(row_start),
             ptr(wmask), W, ptr(out), H, ptr(a), NL, ptr(g), H, current_stream())
        ctx.pack = pack
        ctx.save_for_backward(a, g)
        ctx.shapes = (tuple(alpha.shape), tuple(gamma.shape))
        return out

    @staticmethod
    def backward(ctx, dy):
        hf, hb, layer_stride, words, n_words, row_start, wmask, N, W, NL, H = ctx.pack
        a, g = ctx.saved_tensors
        dy = _c(dy)
        da = _f32((NL,), dy.device)
        dg = _f32((1,), dy.device)
        call("ruart_subword_layers_backward", ptr(hf), ptr(hb), layer_stride, ptr(words), n_words, ptr(row_start),
             ptr(wmask), W, ptr(dy), H, ptr(a), NL, ptr(g), H, ptr(_ws(dy.device, 256 * NL)), ptr(da), ptr(dg), 0,
             current_stream())
        return da.view(ctx.shapes[0]), dg.view(ctx.shapes[1]), None


def subword_mix(alpha, gamma, pack):
    return SubwordMixFn.apply(alpha, gamma, pack)


# ------------------------------------------------------------------------------------ (Bi)LSTM layer
def _shift_prev(out, H, dir_, reverse):
    """h_{t-1} of every step for one direction: `out` shifted by one step along time (zeros at the start)."""
    B, L, _ = out.shape
    prev = torch.zeros((B, L, H), dtype=torch.float32, device=out.device)
    sl = out[:, :, dir_ * H:(dir_ + 1) * H]
    if L > 1:
        if reverse:
            prev[:, :-1].copy_(sl[:, 1:])
        else:
            prev[:, 1:].copy_(sl[:, :-1])
    return prev


class LstmLayerFn(Function):
    """One layer of StackedBRNN (nn.LSTM batch_first, 1 or 2 directions, H <= 128; Layers.py:137,166) on
    the padded [B, L, in] tensor.  params = (w_ih, w_hh, b_ih, b_hh) per direction, flattened."""

    @staticmethod
    def forward(ctx, x, parts, ndir, *params):
        _need(x)
        x = _c(x)
        B, L, Kin = x.shape
        w_ih = [params[4 * d + 0] for d in range(ndir)]
        w_hh = [params[4 * d + 1] for d in range(ndir)]
        b_ih = [params[4 * d + 2] for d in range(ndir)]
        b_hh = [params[4 * d + 3] for d in range(ndir)]
        H = w_hh[0].shape[1]
        wcat = torch.cat([w.detach() for w in w_ih], 0).contiguous()                       # [ndir*4H, in]
        bias = torch.cat([(bi.detach() + bh.detach()) for bi, bh in zip(b_ih, b_hh)], 0).contiguous()
        whh = torch.stack([w.detach() for w in w_hh], 0).contiguous()                      # [ndir, 4H, H]
        x2 = x.view(B * L, Kin)
        a, Kp = K.split_act(x2, parts)
        ws, _ = K.split_act(wcat, parts)
        xg = _f32((B * L, ndir * 4 * H), x.device)
        _gemm(a, Kp, ws, B * L, ndir * 4 * H, parts, xg, epi=ops.EPI_BIAS, bias=bias)
        out = _f32((B, L, ndir * H), x.device)
        gates = _f32((B * L, ndir * 5 * H), x.device)
        call("ruart_lstm_recurrence_train", ptr(xg), xg.stride(0), ptr(whh), ptr(out), ndir * H, B, L, H, ndir,
             ptr(gates), gates.stride(0), current_stream())
        ctx.save_for_backward(x2, wcat, whh, gates, out)
        ctx.meta = (B, L, Kin, H, ndir, int(parts))
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, wcat, whh, gates, out = ctx.saved_tensors
        B, L, Kin, H, ndir, parts = ctx.meta
        dout = _c(dout)
        rows = B * L
        G = ndir * 4 * H
        dxg = _f32((rows, G), dout.device)
        call("ruart_lstm_recurrence_backward", ptr(gates), gates.stride(0), ptr(whh), ptr(dout), ndir * H, ptr(dxg), G,
             B, L, H, ndir, current_stream())
        # dx = dxg W_ih  (dgrad) ; dW_ih = dxg^T x (wgrad) ; db = column sums ; dW_hh = dxg_dir^T h_prev_dir
        dx = None
        if ctx.needs_input_grad[0]:
            a, Gp = K.split_act(dxg, parts)
            wt, _ = split_t(wcat, parts)
            dx = _f32((rows, Kin), dout.device)
            _gemm(a, Gp, wt, rows, Kin, parts, dx)
            dx = dx.view(B, L, Kin)
        at, Rp = split_t(dxg, parts)                 # [G, parts*Rp]
        xt, _ = split_t(x2, parts)                   # [Kin, parts*Rp]
        dw_all = _f32((G, Kin), dout.device)
        _gemm(at, Rp, xt, G, Kin, parts, dw_all)
        db_all = colsum(dxg)
        grads = []
        for d in range(ndir):
            prev = _shift_prev(out, H, d, reverse=(d == 1)).view(rows, H)
            ht, _ = split_t(prev, parts)             # [H, parts*Rp]
            atd = at[d * 4 * H:(d + 1) * 4 * H]      # rows of dxg^T that belong to this direction
            dwh = _f32((4 * H, H), dout.device)
            _gemm(atd, Rp, ht, 4 * H, H, parts, dwh)
            db = db_all[d * 4 * H:(d + 1) * 4 * H]
            grads += [dw_all[d * 4 * H:(d + 1) * 4 * H], dwh, db, db.clone()]
        return (dx, None, None) + tuple(grads)


def lstm_layer(x, rnn, parts=3):
    """x [B, L, in] through the nn.LSTM parameter holder `rnn` (1 layer, uni- or bidirectional)."""
    sfx = ["", "_reverse"][:2 if rnn.bidirectional else 1]
    params = []
    for s in sfx:
        params += [getattr(rnn, "weight_ih_l0" + s), getattr(rnn, "weight_hh_l0" + s),
                   getattr(rnn, "bias_ih_l0" + s), getattr(rnn, "bias_hh_l0" + s)]
    return LstmLayerFn.apply(x, parts, len(sfx), *params)


# ------------------------------------------------------------------------------------ multi2one
class Multi2OneFn(Function):
    """The word -> item uni-LSTM (hidden 300) over the REAL word steps only, final states scattered into the
    slot tensor (SDNet.py:137,270-271,300-318) — the training form of the step-synchronous path of
    SDNet.forward.  `plan` = (a_rows int32 [n_step_rows], last int32 [n_all], slot_off int64 [n_all] (in
    floats), n_t list, n_slots)."""

    @staticmethod
    def forward(ctx, items_in, w_ih, w_hh, b_ih, b_hh, parts, plan):
        a_rows, last, slot_off, n_t, n_slots = plan
        items_in = _c(items_in)
        XD = items_in.shape[1]
        HS = w_hh.shape[1]
        n_step_rows = int(sum(n_t))
        n_all = int(n_t[0]) if n_t else 0
        dev = items_in.device
        A = _f32((n_step_rows, XD), dev)
        K.gather_rows(items_in, a_rows, A, None, n_step_rows, XD)
        wi = _c(w_ih.detach())
        wh = _c(w_hh.detach())
        bias = (b_ih.detach() + b_hh.detach()).contiguous()
        a_sp, Kp_in = K.split_act(A, parts)
        wis, _ = K.split_act(wi, parts)
        whs, Kp_h = K.split_act(wh, parts)
        gx = _f32((n_step_rows, 4 * HS), dev)
        _gemm(a_sp, Kp_in, wis, n_step_rows, 4 * HS, parts, gx, epi=ops.EPI_BIAS, bias=bias)
        c_state = torch.zeros((n_all, HS), dtype=torch.float32, device=dev)
        h_split = torch.zeros((n_all, parts * Kp_h), dtype=torch.bfloat16, device=dev)
        gh = _f32((max(n_all, 1), 4 * HS), dev)
        slots = torch.zeros((n_slots, HS), dtype=torch.float32, device=dev)
        save = _f32((n_step_rows, 6 * HS), dev)
        st = current_stream()
        row0 = 0
        for t, n in enumerate(n_t):
            n = int(n)
            if t > 0:
                _gemm(h_split, Kp_h, whs, n, 4 * HS, parts, gh)
            call("ruart_lstm_cell_train", gx.data_ptr() + row0 * 4 * HS * 4, ptr(gh) if t > 0 else None, ptr(c_state),
                 ptr(h_split), parts, Kp_h, HS, n, ptr(last), t, ptr(slot_off), ptr(slots),
                 save.data_ptr() + row0 * 6 * HS * 4, st)
            row0 += n
        ctx.save_for_backward(A, wi, wh, save, a_rows, last, slot_off)
        ctx.meta = ([int(n) for n in n_t], int(parts), tuple(items_in.shape), HS, XD)
        return slots

    @staticmethod
    def backward(ctx, dslots):
        A, wi, wh, save, a_rows, last, slot_off = ctx.saved_tensors
        n_t, parts, in_shape, HS, XD = ctx.meta
        dslots = _c(dslots)
        dev = dslots.device
        n_step_rows = int(sum(n_t))
        n_all = n_t[0] if n_t else 0
        st = current_stream()
        dG = _f32((n_step_rows, 4 * HS), dev)
        dc = torch.zeros((n_all, HS), dtype=torch.float32, device=dev)
        starts = [0]
        for n in n_t:
            starts.append(starts[-1] + n)
        wht, _ = split_t(wh, parts)                     # [HS, parts * round_up(4HS, 64)]
        dh_rec, n_rec = None, 0
        for t in range(len(n_t) - 1, -1, -1):
            n = n_t[t]
            sv = save.data_ptr() + starts[t] * 6 * HS * 4
            svp = save.data_ptr() + starts[t - 1] * 6 * HS * 4 if t > 0 else None
            call("ruart_lstm_cell_backward", sv, svp, ptr(dslots), ptr(slot_off), ptr(last), t, ptr(dh_rec), n_rec,
                 ptr(dc), dG.data_ptr() + starts[t] * 4 * HS * 4, HS, n, st)
            if t > 0:
                dGt = dG[starts[t]:starts[t + 1]]
                a, Gp = K.split_act(dGt, parts)
                dh_rec = _f32((n, HS), dev)
                _gemm(a, Gp, wht, n, HS, parts, dh_rec)
                n_rec = n
        # weight gradients over all steps at once
        at, Rp = split_t(dG, parts)
        xt, _ = split_t(A, parts)
        dwi = _f32((4 * HS, XD), dev)
        _gemm(at, Rp, xt, 4 * HS, XD, parts, dwi)
        db = colsum(dG)
        n_rec_rows = n_step_rows - n_all
        if n_rec_rows > 0:
            hprev = _f32((n_rec_rows, HS), dev)
            o = 0
            for t in range(1, len(n_t)):
                n = n_t[t]
                hprev[o:o + n].copy_(save[starts[t - 1]:starts[t - 1] + n, 5 * HS:6 * HS])
                o += n
            at2, Rp2 = split_t(dG[n_all:], parts)
            ht2, _ = split_t(hprev, parts)
            dwh = _f32((4 * HS, HS), dev)
            _gemm(at2, Rp2, ht2, 4 * HS, HS, parts, dwh)
        else:
            dwh = torch.zeros((4 * HS, HS), dtype=torch.float32, device=dev)
        d_items = None
        if ctx.needs_input_grad[0]:
            a, Gp = K.split_act(dG, parts)
            wit, _ = split_t(wi, parts)
            dA = _f32((n_step_rows, XD), dev)
            _gemm(a, Gp, wit, n_step_rows, XD, parts, dA)
            d_items = torch.zeros(in_shape, dtype=torch.float32, device=dev)
            K.gather_rows(dA, None, d_items, a_rows, n_step_rows, XD)
        return d_items, dwi, dwh, db, db.clone(), None, None


def multi2one(items_in, rnn, plan, parts=3):
    return Multi2OneFn.apply(items_in, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0, parts, plan)
