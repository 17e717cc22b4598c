"""Drop-in for the reference's Models/Layers.py: same classes, constructor signatures, parameter
names (state_dict compatible) and forward signatures — with the arithmetic done by the sm_100a
kernels of libruart_b200.so.  Every forward needs CUDA tensors and raises otherwise (no CPU /
eager fallback).

Two execution forms, same results:
  * fused inference form (`SDNet.forward` under `torch.no_grad()` / `.eval()`): projections on the
    tcgen05 GEMM, fused attention tails / scorer / pooling kernels, outputs written into concat buffers;
  * differentiable form (`grad_mode(module)`: autograd recording and the module in train mode — what
    `SDNetTrainer.update` does, SDNetTrainer.py:332-337): every op is a torch.autograd.Function of
    ruart_b200/autograd_ops.py with a hand-written backward kernel.  The `Layers.py`-level forwards
    that the fused form never calls (`AttentionScore.forward`, `LinearSelfAttn.forward`,
    `BilinearSeqAttn.forward`, `GetFinalScores.get_single_score`, `weighted_avg`) are always this form.
Dropout (Layers.py:23-39) is the identity in eval mode; in train mode the masks are drawn with
torch.bernoulli on the device (the RNG is the only library call) and applied by an own kernel.

Reference map: StackedBRNN Layers.py:124-180 | AttentionScore :182-245 | Attention :247-295 |
RNN_from_opt :297-317 | LinearSelfAttn :320-341 | GetFinalScores :352-432 |
BilinearSeqAttn :435-468 | DeepAttention :471-524 | weighted_avg :529-534.
"""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from .. import autograd_ops as A
from .. import ops
from .. import sdnet_ops as K
from .._lib import current_stream, ptr
from ..ops import call

dropout_p = 0.0
do_seq_dropout = False
sdnet_parts = 3  # split parts of the SDNet-stack GEMM operands (3 = fp32-grade)


def set_dropout_prob(p):
    global dropout_p
    dropout_p = p


def set_seq_dropout(option):  # option = True or False
    global do_seq_dropout
    do_seq_dropout = option


def set_sdnet_precision(parts):
    global sdnet_parts
    assert parts in (1, 2, 3)
    sdnet_parts = parts


train_parts = 3  # split width of the GEMM operands in the differentiable form (3 = fp32 grade)


def set_train_precision(parts):
    """Split width of the differentiable form's GEMM operands (forward, dgrad, wgrad): 3 = fp32 grade (6 products),
    2 = ~2^-16 relative (3 products) — what SDNet picks next to a bf16 BERT, like the inference path."""
    global train_parts
    assert parts in (2, 3)
    train_parts = parts


def grad_mode(module=None):
    """True when the differentiable form must run: autograd is recording and the module trains."""
    return torch.is_grad_enabled() and (module is None or module.training)


def _no_training(module, p):
    """The fused inference kernels have no dropout: reaching one in train mode with p > 0 is a bug."""
    if module.training and p > 0 and not torch.is_grad_enabled():
        raise NotImplementedError("train-mode dropout needs the differentiable form (autograd enabled); "
                                  "call .eval() for inference")


def seq_dropout(x, p=0, training=False):
    """Variational dropout: one mask per (batch, feature), shared along the sequence (Layers.py:23-30)."""
    if training == False or p == 0:
        return x
    keep = torch.bernoulli(torch.full((x.size(0), x.size(2)), 1.0 - p, dtype=torch.float32, device=x.device))
    keep = keep * (1.0 / (1.0 - p))
    return A.mul_const(x, keep.unsqueeze(1).expand_as(x).contiguous())


def dropout(x, p=0, training=False):
    """Layers.py:32-39: seq_dropout for 3-D inputs when do_seq_dropout, else element-wise dropout."""
    if training == False or p == 0:
        return x
    if do_seq_dropout and x.dim() == 3:
        return seq_dropout(x, p=p, training=training)
    keep = torch.bernoulli(torch.full(tuple(x.shape), 1.0 - p, dtype=torch.float32, device=x.device))
    return A.mul_const(x, keep * (1.0 / (1.0 - p)))


def _need_cuda(*ts):
    for t in ts:
        if t is not None and torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError("ruart_b200 layers run on CUDA tensors only; there is no CPU fallback")


class StackedBRNN(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers, rnn_type=nn.LSTM, concat_layers=False,
                 bidirectional=True, add_feat=0, LN=False, batch_size=None, max_len=None):
        super(StackedBRNN, self).__init__()
        assert rnn_type is nn.LSTM, "only nn.LSTM is on RUArt's path"
        self.bidir_coef = 2 if bidirectional else 1
        self.num_layers = num_layers
        self.concat_layers = concat_layers
        self.hidden_size = hidden_size
        self.rnns = nn.ModuleList()
        self.LN = LN
        if LN:
            self.ln = nn.LayerNorm([batch_size, max_len, self.bidir_coef * self.hidden_size])
        for i in range(num_layers):
            in_size = input_size if i == 0 else (
                self.bidir_coef * hidden_size + add_feat if i == 1 else self.bidir_coef * hidden_size)
            # parameter holder only (weight_ih_l0, weight_hh_l0, bias_*_l0[_reverse]); never called
            self.rnns.append(nn.LSTM(in_size, hidden_size, num_layers=1, bidirectional=bidirectional,
                                     batch_first=True))

    @property
    def output_size(self):
        if self.concat_layers:
            return self.num_layers * self.bidir_coef * self.hidden_size
        return self.bidir_coef * self.hidden_size

    def _dir_params(self, i):
        r = self.rnns[i]
        sfx = ["", "_reverse"][:self.bidir_coef]
        return ([getattr(r, "weight_ih_l0" + s) for s in sfx], [getattr(r, "weight_hh_l0" + s) for s in sfx],
                [getattr(r, "bias_ih_l0" + s) for s in sfx], [getattr(r, "bias_hh_l0" + s) for s in sfx])

    def run_layer(self, i, x, out=None, LN=None):
        """Layer i on [B, L, in] -> [B, L, ndir*H] (optionally into the strided view `out`).  `x` may be a list of
        [B, L, D_k] tensors standing for their concatenation along the last dim (never materialised)."""
        H = self.hidden_size
        x0 = x[0] if isinstance(x, (list, tuple)) else x
        if out is None:
            out = torch.empty((x0.shape[0], x0.shape[1], self.bidir_coef * H), dtype=torch.float32, device=x0.device)
        w_ih, w_hh, b_ih, b_hh = self._dir_params(i)
        if H <= 128:
            K.lstm_layer(self, x, i, w_ih, w_hh, b_ih, b_hh, H, sdnet_parts, out, whole_ln=bool(LN))
        else:
            self._run_layer_stepwise(i, x, out)
            if LN:
                K.whole_layernorm_(out)
        return out

    def _run_layer_stepwise(self, i, x, out):
        """Hidden sizes that do not fit the persistent kernel (multi2one: 300): one recurrent GEMM
        + cell kernel per step.  Unidirectional only (the shipped conf's multi2one)."""
        assert self.bidir_coef == 1, "step-synchronous path is unidirectional"
        H = self.hidden_size
        B, L, _ = x.shape
        w_ih, w_hh, b_ih, b_hh = self._dir_params(i)
        parts = sdnet_parts
        a, Kp_in = K.split_act(x, parts)
        wi, _ = K.prep_weight(self, (i, "w_ih"), w_ih, parts)
        wh, Kp_h = K.prep_weight(self, (i, "w_hh"), w_hh, parts)
        bias = K.prep_vector(self, (i, "bias"), lambda: b_ih[0] + b_hh[0], b_ih + b_hh)
        gx = torch.empty((B * L, 4 * H), dtype=torch.float32, device=x.device)
        K.linear(a, Kp_in, wi, B * L, 4 * H, parts, gx, epi=ops.EPI_BIAS, bias=bias)
        c = torch.zeros((B, H), dtype=torch.float32, device=x.device)
        hs = torch.zeros((B, parts * Kp_h), dtype=torch.bfloat16, device=x.device)
        gh = torch.empty((B, 4 * H), dtype=torch.float32, device=x.device)
        last = torch.zeros(B, dtype=torch.int32, device=x.device)
        op = K.rows2d(out)[2]
        rows_b = torch.arange(B, device=x.device, dtype=torch.int32) * L
        offs_b = torch.arange(B, device=x.device, dtype=torch.int64) * L * op
        for t in range(L):
            if t > 0:
                K.linear(hs, Kp_h, wh, B, 4 * H, parts, gh)
            last.fill_(t)
            # keep both index tensors alive across the call: two unnamed temporaries would share
            # one freed allocator block
            rg, so = rows_b + t, offs_b + t * op
            call("ruart_lstm_cell", ptr(gx), ptr(rg), ptr(gh) if t > 0 else None, ptr(c), ptr(hs), parts, Kp_h, H,
                 B, ptr(last), t, ptr(so), ptr(out), current_stream())

    def forward(self, x, x_mask, return_list=False, x_additional=None, LN=None):
        _need_cuda(x)
        if grad_mode(self):
            return self._forward_differentiable(x, return_list, x_additional, LN)
        _no_training(self, dropout_p)
        hiddens = [x]
        for i in range(self.num_layers):
            rnn_input = hiddens[-1]
            if i == 1 and x_additional is not None:
                rnn_input = K.concat_cols([rnn_input, x_additional])
            hiddens.append(self.run_layer(i, rnn_input.contiguous(), LN=LN))
        output = K.concat_cols(hiddens[1:]) if self.concat_layers else hiddens[-1]
        if return_list:
            return output, hiddens[1:]
        return output


def _stacked_brnn_differentiable(self, x, return_list=False, x_additional=None, LN=None):
    """StackedBRNN.forward (Layers.py:156-180) with autograd: per layer dropout -> (Bi)LSTM over the padded
    tensor -> whole-tensor LayerNorm; hidden sizes <= 128 (the persistent kernel + BPTT kernel)."""
    if self.hidden_size > 128:
        raise NotImplementedError("differentiable StackedBRNN needs hidden_size <= 128 (multi2one goes through "
                                  "autograd_ops.multi2one inside SDNet.forward)")
    hiddens = [x]
    for i in range(self.num_layers):
        rnn_input = hiddens[-1]
        if i == 1 and x_additional is not None:
            rnn_input = torch.cat((rnn_input, x_additional), 2)
        if dropout_p > 0:
            rnn_input = dropout(rnn_input, p=dropout_p, training=self.training)
        out = A.lstm_layer(rnn_input, self.rnns[i], parts=train_parts)
        if LN:
            out = A.whole_layernorm(out)
        hiddens.append(out)
    output = torch.cat(hiddens[1:], 2) if self.concat_layers else hiddens[-1]
    if return_list:
        return output, hiddens[1:]
    return output


StackedBRNN._forward_differentiable = _stacked_brnn_differentiable


class AttentionScore(nn.Module):
    """correlation_func 3: s_ij = relu(W x1_i) D relu(W x2_j) (the only variant RUArt constructs)."""

    def __init__(self, input_size, hidden_size, correlation_func=1, do_similarity=False):
        super(AttentionScore, self).__init__()
        self.correlation_func = correlation_func
        self.hidden_size = hidden_size
        if correlation_func == 2 or correlation_func == 3:
            self.linear = nn.Linear(input_size, hidden_size, bias=False)
            if do_similarity:
                self.diagonal = Parameter(torch.ones(1, 1, 1) / (hidden_size ** 0.5), requires_grad=False)
            else:
                self.diagonal = Parameter(torch.ones(1, 1, hidden_size), requires_grad=True)
        if correlation_func == 4:
            self.linear = nn.Linear(input_size, input_size, bias=False)
        if correlation_func == 5:
            self.linear = nn.Linear(input_size, hidden_size, bias=False)

    def project(self, x, with_diag, a_split=None):
        """relu(x W^T) (* diagonal) for [B, L, D] rows -> fp32 [B*L, hidden]; returns (proj, split)."""
        assert self.correlation_func == 3, "only correlation_func=3 is on RUArt's path"
        xs = list(x) if isinstance(x, (list, tuple)) else [x]     # a list = its column-wise concatenation
        rows = xs[0].shape[0] * xs[0].shape[1]
        if a_split is None:
            a_split = K.split_concat(xs, sdnet_parts)
        a, Kp = a_split
        w, _ = K.prep_weight(self, ("w"), [self.linear.weight], sdnet_parts)
        out = torch.empty((rows, self.hidden_size), dtype=torch.float32, device=xs[0].device)
        if with_diag:
            d = K.prep_vector(self, ("d"), lambda: self.diagonal.reshape(-1), [self.diagonal])
        else:
            d = K.ones(xs[0].device)
        K.linear(a, Kp, w, rows, self.hidden_size, sdnet_parts, out, epi=ops.EPI_RELU_SCALE, scale=d)
        return out, a_split

    def forward(self, x1, x2):
        """scores [B, L1, L2] = relu(x1 W^T) D relu(x2 W^T)^T (Layers.py:208-245, correlation_func 3).  The fused
        inference path never materialises this matrix (Attention.forward); this is the differentiable /
        `Layers.py`-level form."""
        _need_cuda(x1, x2)
        assert self.correlation_func == 3, "only correlation_func=3 is on RUArt's path"
        x1 = dropout(x1, p=dropout_p, training=self.training)
        x2 = dropout(x2, p=dropout_p, training=self.training)
        x1_rep = A.linear(x1, self.linear.weight, None, relu=True, parts=train_parts)
        x2_rep = A.linear(x2, self.linear.weight, None, relu=True, parts=train_parts)
        x1_rep = A.scale_cols(x1_rep, self.diagonal)
        return A.bmm(x1_rep, x2_rep, trans_b=True)


class Attention(nn.Module):
    def __init__(self, input_size, hidden_size, correlation_func=1, do_similarity=False):
        super(Attention, self).__init__()
        self.scoring = AttentionScore(input_size, hidden_size, correlation_func, do_similarity)

    def forward(self, x1, x2, x2_mask, x3=None, drop_diagonal=False, return_score=False, out=None,
                add_to_out=False, p2_cache=None):
        """attended[b, i] = sum_j softmax_j(score(x1_i, x2_j) | x2_mask) x3_j   (Layers.py:253-295).
        `out` (strided view) / `add_to_out` / `p2_cache` are extensions used by SDNet.forward."""
        _need_cuda(x1, x2, x3)
        if drop_diagonal or return_score:
            raise NotImplementedError("drop_diagonal / return_score are not used on RUArt's path")
        if x3 is None:
            x3 = x2
        if grad_mode(self):
            # differentiable form (Layers.py:272-288): scores -> masked softmax over keys -> alpha.bmm(x3)
            alpha = A.masked_softmax(self.scoring(x1, x2), x2_mask)
            res = A.bmm(alpha, x3)
            if out is not None:
                raise NotImplementedError("`out=` is an inference-path extension")
            return res
        _no_training(self, dropout_p)
        # x1 / x2 may be lists of tensors (their concatenation along the last dim, never materialised)
        x1_0 = x1[0] if isinstance(x1, (list, tuple)) else x1
        x2_0 = x2[0] if isinstance(x2, (list, tuple)) else x2
        B, L1, L2 = x1_0.shape[0], x1_0.shape[1], x2_0.shape[1]
        if x2 is x1 and p2_cache is None:
            # self-attention (SDNet.py:380-390,411): both sides project the SAME rows through the same Linear, so
            # relu(x W^T) is computed once and the query side is that result times the diagonal — the product the
            # GEMM epilogue would form (relu(acc) * d), hence bit-identical, for a 10 us pass instead of a GEMM
            p2, _ = self.scoring.project(x1, False)
            d = K.prep_vector(self.scoring, ("d"), lambda: self.scoring.diagonal.reshape(-1), [self.scoring.diagonal])
            p1 = torch.empty_like(p2)
            hid = p2.shape[1]
            call("ruart_eltwise", 3, ptr(p2), hid, None, 0, ptr(d), d.numel(), ptr(p1), hid, p2.shape[0], hid,
                 current_stream())
        else:
            p1, sp = self.scoring.project(x1, True)
            if p2_cache is not None and "p2" in p2_cache:
                p2 = p2_cache["p2"]
            else:
                p2, _ = self.scoring.project(x2, False)
                if p2_cache is not None:
                    p2_cache["p2"] = p2
        if out is None:
            out = torch.empty((B, L1, x3.shape[2]), dtype=torch.float32, device=x1_0.device)
        K.attention_tail(p1, p2, K.as_u8(x2_mask), x3, out, B, L1, L2, add=add_to_out, parts=sdnet_parts)
        return out


def RNN_from_opt(input_size_, hidden_size_, num_layers=1, concat_rnn=False, add_feat=0, bidirectional=True,
                 rnn_type=nn.LSTM, LN=False, batch_size=None, max_len=None):
    new_rnn = StackedBRNN(input_size=input_size_, hidden_size=hidden_size_, num_layers=num_layers,
                          rnn_type=rnn_type, concat_layers=concat_rnn, bidirectional=bidirectional,
                          add_feat=add_feat, LN=False, batch_size=batch_size, max_len=max_len)
    output_size = hidden_size_
    if bidirectional:
        output_size *= 2
    if concat_rnn:
        output_size *= num_layers
    return new_rnn, output_size


class LinearSelfAttn(nn.Module):
    """alpha = softmax(mask(W x_i)) over the sequence (Layers.py:320-341)."""

    def __init__(self, input_size):
        super(LinearSelfAttn, self).__init__()
        self.linear = nn.Linear(input_size, 1)

    def pooled(self, x, x_mask):
        """weighted_avg(x, self(x, x_mask)) in one kernel (SDNet.py:414-415)."""
        _need_cuda(x)
        B, L, D = x.shape
        out = torch.empty((B, D), dtype=torch.float32, device=x.device)
        mask8 = K.as_u8(x_mask)  # named: must outlive the launch
        w = self.linear.weight.detach().reshape(-1).contiguous()
        call("ruart_self_attn_pool", ptr(x), K.rows2d(x)[2], B, L, D, ptr(mask8), ptr(w),
             ptr(self.linear.bias.detach()), ptr(out), D, current_stream())
        return out

    def forward(self, x, x_mask):
        """alpha [B, L] = softmax(mask(W x_i + b)) (Layers.py:328-341) — the differentiable / API form; the fused
        inference path uses `pooled` (alpha and weighted_avg in one kernel)."""
        _need_cuda(x)
        x = dropout(x, p=dropout_p, training=self.training)
        B, L, D = x.shape
        scores = A.linear(x.reshape(B * L, D), self.linear.weight, self.linear.bias, parts=train_parts).view(B, L)
        return A.masked_softmax(scores, x_mask)


def generate_mask(new_data, dropout_p=0.0):
    raise NotImplementedError("training-time helper; not part of the inference path")


class GetFinalScores(nn.Module):
    def __init__(self, x_size, h_size, yesno, no_answer, useES):
        super(GetFinalScores, self).__init__()
        self.no_answer = no_answer
        self.yesno = yesno
        self.useES = useES
        if no_answer:
            self.noanswer_linear = nn.Linear(h_size, x_size)
            self.noanswer_w = nn.Linear(x_size, 1, bias=True)
        if yesno:
            self.no_linear = nn.Linear(h_size, x_size)
            self.no_w = nn.Linear(x_size, 1, bias=True)
            self.yes_linear = nn.Linear(h_size, x_size)
            self.yes_w = nn.Linear(x_size, 1, bias=True)
            self.no_read_linear = nn.Linear(h_size, x_size)
            self.no_read_w = nn.Linear(x_size, 1, bias=True)
        self.attn = BilinearSeqAttn(x_size, h_size)
        self.rnn = nn.GRUCell(x_size, h_size)  # parameters kept; its output is unused (Layers.py:395-397)
        self.attn2 = BilinearSeqAttn(x_size, h_size)
        self.x_size = x_size
        self.last_logits = None

    def forward(self, x, h0, x_mask, ES_len, mask_flag=None, nan_flag=None, want_logits=False):
        """softmax([ES scores | OCR scores | no-answer]) (Layers.py:373-419)."""
        _need_cuda(x, h0)
        if grad_mode(self):
            return self._forward_differentiable(x, h0, x_mask, ES_len, mask_flag)
        _no_training(self, dropout_p)
        if self.yesno or not self.no_answer or not self.useES or not mask_flag:
            raise NotImplementedError("only the shipped conf's scorer (useES, label_no_answer, mask_score, "
                                      "no yes/no heads) is implemented")
        B, M, X = x.shape
        parts = sdnet_parts
        w, _ = K.prep_weight(self, ("w3"), [self.attn.linear.weight, self.attn2.linear.weight,
                                                 self.noanswer_linear.weight], parts)
        b3 = K.prep_vector(self, ("b3"), lambda: torch.cat([self.attn.linear.bias, self.attn2.linear.bias,
                                                                 self.noanswer_linear.bias]),
                           [self.attn.linear.bias, self.attn2.linear.bias, self.noanswer_linear.bias])
        a, Kp = K.split_act(h0, parts)
        wy = torch.empty((B, 3 * X), dtype=torch.float32, device=x.device)
        K.linear(a, Kp, w, B, 3 * X, parts, wy, epi=ops.EPI_BIAS, bias=b3)
        probs = torch.empty((B, M + 1), dtype=torch.float32, device=x.device)
        logits = torch.empty((B, M + 1), dtype=torch.float32, device=x.device) if want_logits else None
        mask8 = K.as_u8(x_mask)  # named: must outlive the launch
        nw = self.noanswer_w.weight.detach().reshape(-1).contiguous()
        call("ruart_final_scores", ptr(x), K.rows2d(x)[2], B, M, X, ptr(wy), ptr(mask8), int(ES_len),
             ptr(nw), ptr(self.noanswer_w.bias.detach()), ptr(probs), ptr(logits), ptr(nan_flag), current_stream())
        self.last_logits = logits
        return probs

    def _forward_differentiable(self, x, h0, x_mask, ES_len, mask_flag=None):
        """GetFinalScores.forward (Layers.py:373-419) op by op with autograd.  The GRUCell step of
        :393-397 does not influence the output (its parameters get no gradient in the reference either)."""
        if self.yesno:
            raise NotImplementedError("yes/no heads are outside the shipped conf")
        if self.useES:
            x_es, x_es_mask = x[:, :ES_len], x_mask[:, :ES_len]
            x_ocr, x_ocr_mask = x[:, ES_len:], x_mask[:, ES_len:]
            score_ocr = self.attn(x_ocr, h0, x_ocr_mask, mask_flag=mask_flag)
            score_es = self.attn2(x_es, h0, x_es_mask, mask_flag=mask_flag)
            score_s = torch.cat([score_es, score_ocr], dim=-1)
        else:
            score_s = self.attn(x, h0, x_mask, mask_flag=mask_flag)
        if self.no_answer:
            h0 = dropout(h0, p=dropout_p, training=self.training)
            score_noanswer = self.get_single_score(x, h0, x_mask, self.noanswer_linear, self.noanswer_w)
            score_s = torch.cat([score_s, score_noanswer], dim=-1)
        self.last_logits = score_s.detach()
        return A.masked_softmax(score_s, None)

    def get_single_score(self, x, h, x_mask, linear, w):
        """w(softmax(mask(x . linear(h))) . x) -> [B, 1]  (Layers.py:421-432)."""
        _need_cuda(x, h)
        Wh = A.linear(h, linear.weight, linear.bias, parts=train_parts)
        xWh = A.bmm(x, Wh.unsqueeze(2)).squeeze(2)
        beta = A.masked_softmax(xWh, x_mask)
        attn_x = A.bmm(beta.unsqueeze(1), x)                      # [B, 1, x_size]
        return A.linear(attn_x, w.weight, w.bias, parts=train_parts).squeeze(2)


class BilinearSeqAttn(nn.Module):
    """o_i = x_i' (W y + b) (Layers.py:435-468).  The fused inference path evaluates it inside
    GetFinalScores' kernel; `forward` is the differentiable / API form."""

    def __init__(self, x_size, y_size, identity=False):
        super(BilinearSeqAttn, self).__init__()
        self.linear = nn.Linear(y_size, x_size) if not identity else None

    def forward(self, x, y, x_mask, mask_flag=True):
        _need_cuda(x, y)
        x = dropout(x, p=dropout_p, training=self.training)
        y = dropout(y, p=dropout_p, training=self.training)
        Wy = A.linear(y, self.linear.weight, self.linear.bias, parts=train_parts) if self.linear is not None else y
        xWy = A.bmm(x, Wy.unsqueeze(2)).squeeze(2)                 # [B, len]
        if mask_flag:
            xWy = A.mask_fill_neg_inf(xWy, x_mask)
        return xWy


class DeepAttention(nn.Module):
    def __init__(self, opt, abstr_list_cnt, deep_att_hidden_size_per_abstr, correlation_func=1,
                 word_hidden_size=None):
        super(DeepAttention, self).__init__()
        word_hidden_size = opt['embedding_dim'] if word_hidden_size is None else word_hidden_size
        abstr_hidden_size = opt['hidden_size'] * 2
        if 'no_DeepAttention' in opt:
            att_size = 0
            rnn_input_size = abstr_hidden_size * abstr_list_cnt
        else:
            att_size = abstr_hidden_size * abstr_list_cnt + word_hidden_size
            self.int_attn_list = nn.ModuleList()
            for i in range(abstr_list_cnt + 1):
                self.int_attn_list.append(Attention(att_size, deep_att_hidden_size_per_abstr,
                                                    correlation_func=correlation_func))
            rnn_input_size = abstr_hidden_size * abstr_list_cnt * 2 + (opt['highlvl_hidden_size'] * 2)
        self.att_size = att_size
        self.rnn_input_size = rnn_input_size
        self.rnn, self.output_size = RNN_from_opt(rnn_input_size, opt['highlvl_hidden_size'], num_layers=1)
        self.opt = opt

    def _fused_weights(self):
        """The abstr_list_cnt+1 attention heads share their inputs, so their projections run as ONE
        GEMM per side with the weights stacked along N (and the trainable diagonals concatenated)."""
        lins = [a.scoring.linear.weight for a in self.int_attn_list]
        diags = [a.scoring.diagonal for a in self.int_attn_list]
        w, _ = K.prep_weight(self, ("w_all",), lins, sdnet_parts)
        hid = self.int_attn_list[0].scoring.hidden_size
        d = K.prep_vector(self, ("d_all",), lambda: torch.cat([
            x.reshape(-1) if x.numel() > 1 else x.reshape(-1).expand(hid) for x in diags]), diags)
        return w, d, hid

    def project_x2(self, x2_word, x2_abstr):
        """relu(x2_att W_i^T) for every head i: [B*L2, heads*hid] (shared by every x1 it is paired with)."""
        srcs = x2_word + x2_abstr[:-1]                 # x2_att = torch.cat(srcs, 2), never materialised
        w, _, hid = self._fused_weights()
        n = len(self.int_attn_list) * hid
        a, Kp = K.split_concat(srcs, sdnet_parts)
        rows = srcs[0].shape[0] * srcs[0].shape[1]
        p2 = torch.empty((rows, n), dtype=torch.float32, device=srcs[0].device)
        K.linear(a, Kp, w, rows, n, sdnet_parts, p2, epi=ops.EPI_RELU_SCALE, scale=K.ones(srcs[0].device))
        return p2

    def forward(self, x1_word, x1_abstr, x2_word, x2_abstr, x1_mask, x2_mask, return_bef_rnn=False,
                return_score=False, x2_proj=None):
        """History-of-word multi-level inter-attention (Layers.py:493-524).  x2_proj: optional
        result of project_x2 (the question side is the same for the OCR and the OD call)."""
        _need_cuda(*x1_abstr)
        if return_score or 'no_DeepAttention' in self.opt:
            raise NotImplementedError("return_score / no_DeepAttention are not on the shipped conf's path")
        if grad_mode(self):
            # differentiable form (Layers.py:498-524)
            x1_att = torch.cat(x1_word + x1_abstr, 2)
            x2_att = torch.cat(x2_word + x2_abstr[:-1], 2)
            x1 = torch.cat(x1_abstr, 2)
            for i in range(len(x2_abstr)):
                x1 = torch.cat((x1, self.int_attn_list[i](x1_att, x2_att, x2_mask, x3=x2_abstr[i])), 2)
            x1_hiddens = self.rnn(x1, x1_mask)
            return (x1_hiddens, x1) if return_bef_rnn else x1_hiddens
        _no_training(self, dropout_p)
        # fused inference form.  Neither x1_att = cat(x1_word + x1_abstr) nor x1 = cat(x1_abstr + attended) is
        # materialised: the projection GEMM and the BiLSTM's input GEMM read split-bf16 operands built straight
        # from the pieces (K.split_concat), and `before-rnn` is returned as the LIST of its pieces.
        srcs = x1_word + x1_abstr
        B, L1, L2 = srcs[0].shape[0], srcs[0].shape[1], x2_abstr[0].shape[1]
        dev = srcs[0].device
        w, d, hid = self._fused_weights()
        n = len(self.int_attn_list) * hid
        a1, Kp = K.split_concat(srcs, sdnet_parts)
        p1 = torch.empty((B * L1, n), dtype=torch.float32, device=dev)
        K.linear(a1, Kp, w, B * L1, n, sdnet_parts, p1, epi=ops.EPI_RELU_SCALE, scale=d)
        p2 = x2_proj if x2_proj is not None else self.project_x2(x2_word, x2_abstr)
        mask = K.as_u8(x2_mask)
        att = torch.empty((B, L1, sum(t.shape[2] for t in x2_abstr)), dtype=torch.float32, device=dev)
        nh = len(x2_abstr)
        x3s = [K.rows2d(t) for t in x2_abstr]          # (2-D view, width, pitch)
        if (sdnet_parts == 2 and L2 <= 128 and nh <= 4 and len({(r[1], r[2]) for r in x3s}) == 1):
            # the heads share p1 / p2 / mask and have equally shaped x3: ONE launch with the head on grid.z
            # (three 66 us launches of a latency-bound kernel -> one; same arithmetic per head)
            import ctypes
            ptrs = (ctypes.c_void_p * nh)(*[t.data_ptr() for t in x2_abstr])
            call("ruart_attention_tail_heads", ptr(p1), n, ptr(p2), K.rows2d(p2)[2], hid, nh, ptr(mask), ptrs,
                 x3s[0][2], x3s[0][1], ptr(att), K.rows2d(att)[2], B, L1, L2, current_stream())
        else:
            col = 0
            for i in range(nh):
                x3 = x2_abstr[i]
                K.attention_tail(p1[:, i * hid:(i + 1) * hid], p2[:, i * hid:(i + 1) * hid], mask, x3,
                                 att[:, :, col:col + x3.shape[2]], B, L1, L2, parts=sdnet_parts)
                col += x3.shape[2]
        x1 = list(x1_abstr) + [att]
        x1_hiddens = self.rnn.run_layer(0, x1) if self.rnn.num_layers == 1 else self.rnn(K.concat_cols(x1), x1_mask)
        if return_bef_rnn:
            return x1_hiddens, x1
        return x1_hiddens


def weighted_avg(x, weights):
    """x [B, len, d], weights [B, len] -> [B, d]  (Layers.py:529-534)."""
    return A.bmm(weights.unsqueeze(1), x).squeeze(1)


# Present in the reference but never constructed with the shipped conf.
class CNN(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("CNN is unused on RUArt's inference path")


class MaxPooling(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("MaxPooling is unused on RUArt's inference path")


class AveragePooling(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("AveragePooling is unused on RUArt's inference path")


class LinearTransform(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("LinearTransform is unused on RUArt's inference path")
