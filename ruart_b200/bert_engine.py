"""Packed (pad-free) BERT encoder on sm_100a kernels.

Replaces BertModel.forward (reference Models/Bert/modeling.py:585-614) + Bert.combine_forward
(Models/Bert/Bert.py:130-176) + SDNet.linear_sum (Models/SDNet.py:573-583) for inference:

    rows of real wordpieces  --embed+LN-->  h
    12 x [ QKV GEMM (tcgen05) -> varlen attention -> out-proj GEMM -> add+LN
           -> FFN-up GEMM + erf-GELU epilogue -> FFN-down GEMM -> add+LN
           -> subword mean * softmax(alpha)[l] * gamma accumulated into the word slots ]

The 12 per-layer outputs are never materialised ([N, L, 9216] in the reference, Bert.py:137).

Three numeric modes
  "bf16"  : activations and GEMM operands bf16, fp32 accumulation (tensor cores, 1 term)
  "bf16x2": activations fp32; GEMM operands are 2-part bf16 splits multiplied as 3 terms (~2^-16): the
            accuracy of a TF32-class GEMM at 3x the bf16 tensor time — for weight sets on which plain bf16
            operands move the answers (the "pretrained-like" chaotic random net: 96.5 % agreement at B=256)
  "fp32"  : activations fp32; GEMM operands are 3-part bf16 splits multiplied as 6 terms (~2^-24)
"""
import itertools
import os

import numpy as np
import torch

from . import ops
from .sdnet_ops import upload
from ._lib import current_stream, ptr
from .ops import call

WINDOW = 512  # BERT_MAX_LEN, Models/Bert/Bert.py:18


class Segment(object):
    """One Bert.forward call's worth of input: ids/mask [N, L], word offsets, word mask [N, W]."""

    def __init__(self, ids, mask, offsets, word_mask, totals=None):
        self.ids, self.mask, self.offsets, self.word_mask = ids, mask, offsets, word_mask
        self.N, self.L = ids.shape
        self.W = word_mask.shape[1]
        # (real tokens of the list, longest 512-token window) when the collate step already counted
        # them on the host (Utils.collate.attach_index_tensors): the packing then needs no read-back
        self.totals = None if totals is None else (int(totals[0]), int(totals[1]))


def flatten_offsets(offsets, n_rows):
    """list[N][n_words][2] -> int32 [4, n_words_total] rows (item, j, st, ed).  An int32 [4, n] array
    (or CPU tensor) — the CSR form Utils.collate.attach_index_tensors precomputes — passes through."""
    if torch.is_tensor(offsets):
        offsets = offsets.numpy()
    if isinstance(offsets, np.ndarray):
        if offsets.ndim != 2 or offsets.shape[0] != 4 or offsets.dtype != np.int32:
            raise ValueError("precomputed word offsets must be int32 [4, n_words]")
        if offsets.shape[1] and not (0 <= int(offsets[0].min()) and int(offsets[0].max()) < n_rows):
            raise ValueError("precomputed word offsets name rows outside the batch")
        return np.ascontiguousarray(offsets)
    try:
        counts = np.fromiter(map(len, offsets), dtype=np.int64, count=n_rows)
        total = int(counts.sum())
        flat = np.fromiter(itertools.chain.from_iterable(itertools.chain.from_iterable(offsets)),
                           dtype=np.int32, count=2 * total)
    except TypeError:
        # VQA_Dataset.bertify emits a flat [1, 1] for an item without words (VQA_Dataset.py:426-427)
        norm = [[o] if (len(o) == 2 and not isinstance(o[0], (list, tuple))) else o for o in offsets]
        counts = np.fromiter(map(len, norm), dtype=np.int64, count=n_rows)
        total = int(counts.sum())
        flat = np.fromiter(itertools.chain.from_iterable(itertools.chain.from_iterable(norm)),
                           dtype=np.int32, count=2 * total)
    words = np.empty((4, total), dtype=np.int32)
    words[0] = np.repeat(np.arange(n_rows, dtype=np.int32), counts)
    starts = np.cumsum(counts) - counts
    words[1] = np.arange(total, dtype=np.int64) - np.repeat(starts, counts)
    pair = flat.reshape(total, 2)
    words[2] = pair[:, 0]
    words[3] = pair[:, 1]
    return words


class BertEngine(object):
    def __init__(self, bert_model, mode="bf16", residual_fp32=False):
        assert mode in ("bf16", "bf16x2", "fp32")
        self.model = bert_model
        self.mode = mode
        self.parts = {"bf16": 1, "bf16x2": 2, "fp32": 3}[mode]     # split width of the GEMM operands
        self.terms = {1: 1, 2: 3, 3: 6}[self.parts]
        # bf16 mode: keep the residual stream / LayerNorm outputs in fp32 as well (what autocast
        # does); the GEMM operands stay bf16.  Measured (tools/bert_error.py): no consistent gain —
        # the bf16 error is dominated by operand rounding — so it is off by default.
        self.residual_fp32 = residual_fp32
        # bf16 mode GELU epilogue: 1 = erf via Abramowitz-Stegun 7.1.26 (2 MUFU ops, |err| < 1e-5),
        # 2 = erf matched by a fitted tanh form (1 MUFU op, |err| < 4e-5, 11 % faster FFN-up GEMM: the
        # epilogue, not the MMA, bounds that GEMM); fp32 mode always uses erff()
        self.gelu_mode = 2
        cfg = bert_model.config
        self.H = cfg.hidden_size
        self.I = cfg.intermediate_size
        self.heads = cfg.num_attention_heads
        self.n_layers = cfg.num_hidden_layers
        assert self.H // self.heads == 64, "attention kernel is specialised for head_dim 64"
        self._key = None
        self._w = None
        self._params = None
        # bf16 mode: the 24 BertLayerNorm passes are folded into the neighbouring GEMMs (fold_weights / the FOLD forms
        # of the CTA-pair GEMM); needs the pair kernel (T >= FOLD_MIN_T) and 768-wide rows.  RUART_NO_LN_FOLD: A/B aid.
        self.fold = (mode == "bf16" and not residual_fp32 and self.H == 768
                     and os.environ.get("RUART_NO_LN_FOLD") is None)
        # ... and the attention runs inside the query/key/value GEMM's epilogue (ruart_qkv_attention_fold) when every
        # sequence fits one 128-row accumulator tile.  RUART_NO_ATTN_FUSE: A/B aid.
        self.fuse_attn = self.fold and os.environ.get("RUART_NO_ATTN_FUSE") is None

    FOLD_MIN_T = 2048

    # ------------------------------------------------------------------ weights
    def _weights_key(self, dev):
        # the Parameter OBJECTS are fixed after construction (load_state_dict / .to() / optimizers change their
        # data in place): walking the module tree on every forward cost ~1.3 ms of host time
        ps = self._params
        if ps is None:
            ps = self._params = tuple(self.model.parameters())
        return (str(dev), self.mode, tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))

    def _prep_matrix(self, w):
        """fp32 [N, K] -> bf16 GEMM operand ([N, K] or 3-part split [N, 3K])."""
        w = w.detach().float().contiguous()
        N, K = w.shape
        assert K % 64 == 0
        P = self.parts
        out = torch.empty((N, P * K), dtype=torch.bfloat16, device=w.device)
        call("ruart_split_bf16", ptr(w), K, None, N, K, K, P, ptr(out), current_stream())
        return out

    def prepare(self, dev):
        key = self._weights_key(dev)
        if key == self._key:
            return self._w
        m = self.model
        f = lambda t: t.detach().float().contiguous()
        W = {"word": f(m.embeddings.word_embeddings.weight), "pos": f(m.embeddings.position_embeddings.weight),
             "type": f(m.embeddings.token_type_embeddings.weight),
             "eg": f(m.embeddings.LayerNorm.gamma), "eb": f(m.embeddings.LayerNorm.beta),
             "eps": float(m.embeddings.LayerNorm.variance_epsilon), "layers": []}
        for lay in m.encoder.layer:
            s = lay.attention.self
            wqkv = torch.cat([s.query.weight, s.key.weight, s.value.weight], 0)
            bqkv = torch.cat([s.query.bias, s.key.bias, s.value.bias], 0)
            W["layers"].append({
                "wqkv": self._prep_matrix(wqkv), "bqkv": f(bqkv),
                "wo": self._prep_matrix(lay.attention.output.dense.weight), "bo": f(lay.attention.output.dense.bias),
                "g1": f(lay.attention.output.LayerNorm.gamma), "b1": f(lay.attention.output.LayerNorm.beta),
                "wi": self._prep_matrix(lay.intermediate.dense.weight), "bi": f(lay.intermediate.dense.bias),
                "wd": self._prep_matrix(lay.output.dense.weight), "bd": f(lay.output.dense.bias),
                "g2": f(lay.output.LayerNorm.gamma), "b2": f(lay.output.LayerNorm.beta),
                "eps": float(lay.output.LayerNorm.variance_epsilon),
            })
        if self.fold:
            self.fold_weights(W)
        self._key, self._w = key, W
        return W

    def fold_weights(self, W):
        """Operands of the folded-LayerNorm encoder (ruart_gemm_bf16_fold).  For a LayerNorm (g, b) whose output
        feeds a dense layer (W0, b0):  LayerNorm(v) W0^T + b0 = r (v (W0 g)^T) - r mu colsum(W0 g) + (W0 b + b0);
        and where it feeds a residual add, LayerNorm(v) + b0' = (v r - mu r) g + (b + b0')."""
        m = self.model
        f = lambda t: t.detach().float().contiguous()

        def consumer(w0, b0, g, b):
            w0 = f(w0)
            wf = self._prep_matrix(w0 * g[None, :])                  # bf16(W0 * gamma)
            return wf, wf.float().sum(1).contiguous(), (w0 @ b + f(b0)).contiguous()

        g_prev, b_prev = W["eg"], W["eb"]                             # layer 0 consumes the embedding LayerNorm
        ln_g, ln_b = [], []
        for lay, lw in zip(m.encoder.layer, W["layers"]):
            s = lay.attention.self
            wqkv = torch.cat([s.query.weight, s.key.weight, s.value.weight], 0)
            lw["wqkv_f"], lw["sqkv"], lw["cqkv"] = consumer(wqkv, lw["bqkv"], g_prev, b_prev)
            # head-major row order [q_h / 8 ; k_h ; v_h] for the fused attention epilogue (the scale 1 / sqrt(64) is a
            # power of two: folding it into the query rows is exact)
            H, nh = self.H, self.heads
            perm = torch.cat([torch.arange(64, device=wqkv.device) + part * H + h * 64
                              for h in range(nh) for part in range(3)])
            qscale = torch.ones(3 * H, device=wqkv.device)
            qscale[:H] = 0.125
            lw["wqkv_p"] = (lw["wqkv_f"].float() * qscale[:, None])[perm].to(torch.bfloat16).contiguous()
            lw["sqkv_p"] = (lw["sqkv"] * qscale)[perm].contiguous()
            lw["cqkv_p"] = (lw["cqkv"] * qscale)[perm].contiguous()
            lw["res1_g"], lw["res1_b"] = g_prev, (b_prev + lw["bo"]).contiguous()
            lw["wi_f"], lw["si"], lw["ci"] = consumer(lay.intermediate.dense.weight, lw["bi"], lw["g1"], lw["b1"])
            lw["res2_g"], lw["res2_b"] = lw["g1"], (lw["b1"] + lw["bd"]).contiguous()
            g_prev, b_prev = lw["g2"], lw["b2"]
            ln_g.append(lw["g2"])
            ln_b.append(lw["b2"])
        W["ln2_g"] = torch.stack(ln_g).contiguous()                   # [n_layers, H] for ruart_subword_coef
        W["ln2_b"] = torch.stack(ln_b).contiguous()

    # ------------------------------------------------------------------ packing
    _pack_streams = {}

    @classmethod
    def pack_begin(cls, segments, want_tiles=False):
        """First half of the token packing: row / window lengths (ruart_seq_lengths) and their prefix
        sums + totals (ruart_seq_scan) on a SIDE stream, and an async copy of the totals into pinned
        memory.  The caller can queue unrelated work on the compute stream before pack_finish()."""
        dev = segments[0].ids.device
        key = str(dev)
        side = cls._pack_streams.get(key)
        if side is None:
            side = torch.cuda.Stream(device=dev)
            cls._pack_streams[key] = side
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)  # inputs produced on the compute stream are visible
        with torch.cuda.stream(side):
            nwins = [(sg.L + WINDOW - 1) // WINDOW for sg in segments]
            row0 = np.zeros(len(segments) + 1, dtype=np.int32)
            seq0 = np.zeros(len(segments) + 1, dtype=np.int32)
            for k, sg in enumerate(segments):
                row0[k + 1] = row0[k] + sg.N
                seq0[k + 1] = seq0[k] + sg.N * nwins[k]
            R, S = int(row0[-1]), int(seq0[-1])
            row_len = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
            win_len = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
            cu_rows = torch.empty(R + 1, dtype=torch.int32, device=dev)
            cu_seq = torch.empty(S + 1, dtype=torch.int32, device=dev)
            totals = torch.empty(1 + len(segments), dtype=torch.int32, device=dev)
            st = current_stream()
            masks = []
            for k, sg in enumerate(segments):
                m8 = sg.mask.contiguous().view(torch.uint8) if sg.mask.dtype == torch.bool \
                    else sg.mask.to(torch.uint8).contiguous()
                masks.append(m8)
                call("ruart_seq_lengths", ptr(m8), sg.N, sg.L, WINDOW, row_len.data_ptr() + 4 * int(row0[k]),
                     win_len.data_ptr() + 4 * int(seq0[k]), st)
            host_known = all(sg.totals is not None for sg in segments)
            # a host-side token count sizes the buffers: the device clamps every offset to it
            max_total = sum(sg.totals[0] for sg in segments) if host_known else -1
            call("ruart_seq_scan", ptr(row_len), R, ptr(win_len), S, len(segments), row0.ctypes.data,
                 seq0.ctypes.data, ptr(cu_rows), ptr(cu_seq), ptr(totals), max_total, st)
            if host_known:
                # counted on the host by the collate step: no device -> host read-back, no host wait
                host = [max_total] + [sg.totals[1] for sg in segments]
                check = torch.empty(totals.shape, dtype=totals.dtype, pin_memory=True)
                check.copy_(totals, non_blocking=True)   # compared by check_totals() after the forward
            else:
                host = torch.empty(totals.shape, dtype=totals.dtype, pin_memory=True)
                host.copy_(totals, non_blocking=True)
                host_known = False
                check = None
            done = torch.cuda.Event()
            done.record(side)
            keep = (row_len, win_len, totals, masks, row0, seq0)
            tiles = None
            if want_tiles:
                # row tiles of whole sequences + per-token sequence bounds for the fused attention epilogue
                # (ruart_seq_tiles: one CTA, ~0.2 ms at cfg-3) — queued AFTER `done`, so that it runs under the token
                # packing and the embedding kernel; sized by the padded token count (the real one is not known here)
                padded = sum(sg.N * sg.L for sg in segments)
                cap = 2 * padded // 128 + 8
                meta = torch.empty(cap, dtype=torch.int32, device=dev)
                bounds = torch.empty((max(padded, 1), 2), dtype=torch.int32, device=dev)
                call("ruart_seq_tiles", ptr(cu_seq), S, ptr(meta), cap, ptr(bounds), st)
                ev = torch.cuda.Event()
                ev.record(side)
                tiles = (meta, bounds, ev)
        return {"segments": segments, "cu_rows": cu_rows, "cu_seq": cu_seq, "host": host, "done": done,
                "nwins": nwins, "keep": keep, "host_known": host_known, "check": check, "tiles": tiles}

    @staticmethod
    def pack_finish(h):
        """Second half: ONE host wait (for the side stream only: total token count and the longest
        sequence per segment size the buffers and pick the kernels), then a pack_tokens kernel
        per segment on the compute stream.

        Returns dict: ids/pos int32 [T]; cu_seqlens int32 [S+1] over 512-token windows; per
        segment: row_start int32 [N], its range of sequences, its longest sequence."""
        segments = h["segments"]
        dev = segments[0].ids.device
        main = torch.cuda.current_stream(dev)
        if not h["host_known"]:
            h["done"].synchronize()
        main.wait_event(h["done"])
        cu_rows, cu_seq, host, nwins = h["cu_rows"], h["cu_seq"], h["host"], h["nwins"]
        cu_rows.record_stream(main)
        cu_seq.record_stream(main)
        T = int(host[0])
        if h["host_known"]:  # host-supplied count: slots the masks do not fill stay valid ids
            ids = torch.zeros(T, dtype=torch.int32, device=dev)
            pos = torch.zeros(T, dtype=torch.int32, device=dev)
        else:
            ids = torch.empty(T, dtype=torch.int32, device=dev)
            pos = torch.empty(T, dtype=torch.int32, device=dev)
        segs = []
        r0 = s0 = 0
        st = current_stream()
        for k, sg in enumerate(segments):
            row_start = cu_rows[r0:r0 + sg.N].contiguous()
            row_start.record_stream(main)
            mask8 = sg.mask.contiguous().view(torch.uint8) if sg.mask.dtype == torch.bool else sg.mask.to(torch.uint8).contiguous()
            ids64 = sg.ids.contiguous()
            call("ruart_pack_tokens", ptr(ids64), ptr(mask8), sg.N, sg.L, ptr(row_start), WINDOW,
                 ptr(ids), ptr(pos), T, st)
            segs.append({"row_start": row_start, "seq0": s0, "seq1": s0 + sg.N * nwins[k],
                         "max_len": int(host[1 + k])})
            r0 += sg.N
            s0 += sg.N * nwins[k]
        return {"ids": ids, "pos": pos, "cu_seqlens": cu_seq, "T": T, "segments": segs,
                "check": (h["check"], list(host)) if h["host_known"] else None, "tiles": h.get("tiles")}

    @staticmethod
    def check_totals(pk):
        """After a device sync: the host-supplied token counts (collate step) must equal what the device
        counted from the masks; anything else means the batch dicts were edited after collate."""
        if pk is None or pk.get("check") is None:
            return
        dev_totals, host = pk["check"]
        if [int(v) for v in dev_totals] != [int(v) for v in host]:
            raise RuntimeError("bert_totals do not match bert_mask (got %s, masks give %s): rebuild them with "
                               "Utils.collate.attach_index_tensors" % (host, [int(v) for v in dev_totals]))

    @classmethod
    def pack(cls, segments, want_tiles=False):
        return cls.pack_finish(cls.pack_begin(segments, want_tiles))

    # ------------------------------------------------------------------ forward
    def _gemm(self, a, w, bias, N, K, epi, out_kind, fast_gelu=False, residual=None):
        """a: activation GEMM operand; returns (f32, bf16) outputs according to mode/out_kind.
        residual (bf16 mode): added in the GEMM epilogue before the output is rounded."""
        T = a.shape[0]
        dev = a.device
        if self.mode == "bf16":
            out = torch.empty((T, N), dtype=torch.bfloat16, device=dev)
            ops.gemm(a, w, T, N, K, epi=epi, bias=bias, out_bf16=out, fast_gelu=fast_gelu, residual=residual)
            return None, out
        P, NT = self.parts, self.terms
        if out_kind == "split":  # consumer is another GEMM only
            out = torch.empty((T, P * N), dtype=torch.bfloat16, device=dev)
            ops.gemm(a, w, T, N, K, a_parts=P, w_parts=P, n_terms=NT, epi=epi, bias=bias,
                     out_bf16=out, out_parts=P, out_part_stride=N)
            return None, out
        out = torch.empty((T, N), dtype=torch.float32, device=dev)
        ops.gemm(a, w, T, N, K, a_parts=P, w_parts=P, n_terms=NT, epi=epi, bias=bias, out_f32=out)
        return out, None

    def encode(self, segments, sinks, alpha=None, gamma=None, pack_handle=None):
        """segments: list[Segment]; sinks: per segment (dst fp32 tensor, dst_stride floats, col_off)
        or a list of n_layers such triples when alpha is None (per-layer outputs).

        With alpha/gamma: dst[item, j, col_off:col_off+H] = sum_l mean_subwords(h_l) * softmax(alpha)_l * gamma.

        All encoder-layer outputs are kept ([n_layers, T, H]) so that the subword averaging runs
        once at the end; the Python word-offset lists are flattened on the host AFTER the whole
        encoder has been queued, i.e. while the GPU is busy.
        """
        pk, hs_f, hs_b = self.encode_hidden(segments, pack_handle, allow_fold=alpha is not None)
        self.apply_sinks(segments, pk, hs_f, hs_b, sinks, alpha, gamma)
        return pk

    def apply_sinks(self, segments, pk, hs_f, hs_b, sinks, alpha=None, gamma=None):
        """Second half of encode(): subword mean (+ learned layer sum) of the kept hidden states into the sinks.
        Host work first (flattening the Python offset lists), then one kernel per segment."""
        T, H, NL = pk["T"], self.H, self.n_layers
        st = current_stream()
        keep32 = hs_f is not None
        layer_stride = T * H
        fold = pk.get("fold")
        if fold is not None:
            # hs_b holds the rows BEFORE each layer's output LayerNorm: the subword kernel normalises them
            if alpha is None:
                raise RuntimeError("per-layer outputs need the unfolded encoder (encode_hidden(allow_fold=False))")
            W = self._w
            dev = hs_b.device
            G = torch.empty((NL, H), dtype=torch.float32, device=dev)
            C = torch.empty((H,), dtype=torch.float32, device=dev)
            call("ruart_subword_coef", ptr(alpha), ptr(gamma), NL, ptr(W["ln2_g"]), ptr(W["ln2_b"]), H, ptr(G), ptr(C), st)
            stats = fold["stats"]
            for k, (sg, (wt, nw, rs, wmask)) in enumerate(zip(segments, self.word_tables(segments, pk))):
                dst, stride, col = sinks[k]
                call("ruart_subword_avg_layers_fold", ptr(hs_b[1:]), layer_stride, ptr(stats[1:]), T * 8,
                     W["layers"][0]["eps"], ptr(wt), nw, ptr(rs), ptr(wmask), sg.W, dst.data_ptr() + 4 * col, stride,
                     ptr(G), ptr(C), NL, H, st)
            return
        for k, (sg, (wt, nw, rs, wmask)) in enumerate(zip(segments, self.word_tables(segments, pk))):
            hf1 = hs_f[1:] if keep32 else None
            hb1 = None if keep32 else hs_b[1:]
            if alpha is not None:
                dst, stride, col = sinks[k]
                call("ruart_subword_avg_layers", ptr(hf1), ptr(hb1), layer_stride, ptr(wt), nw, ptr(rs), ptr(wmask),
                     sg.W, dst.data_ptr() + 4 * col, stride, ptr(alpha), NL, ptr(gamma), H, st)
            else:
                for li in range(NL):
                    dst, stride, col = sinks[k][li]
                    call("ruart_subword_avg_accum", ptr(hf1[li]) if keep32 else None,
                         None if keep32 else ptr(hb1[li]), ptr(wt), nw, ptr(rs), ptr(wmask), sg.W,
                         dst.data_ptr() + 4 * col, stride, None, NL, None, li, 1, H, st)

    def word_tables(self, segments, pk):
        """Per segment: (flattened word offsets int32 [4, n] on the device, n, row_start, uint8 word mask).
        Host work (flattening the Python lists) — call it after the encoder has been queued."""
        dev = segments[0].ids.device
        out = []
        for k, sg in enumerate(segments):
            wt_np = flatten_offsets(sg.offsets, sg.N)
            wt = upload(wt_np, dev)
            wmask = sg.word_mask.contiguous().view(torch.uint8) if sg.word_mask.dtype == torch.bool \
                else sg.word_mask.to(torch.uint8).contiguous()
            out.append((wt, wt_np.shape[1], pk["segments"][k]["row_start"], wmask))
        return out

    def subword_packs(self, segments, pk, hs_f, hs_b):
        """What autograd_ops.SubwordMixFn needs per segment (the differentiable form of the layer mix)."""
        T, H, NL = pk["T"], self.H, self.n_layers
        hf1 = hs_f[1:] if hs_f is not None else None
        hb1 = None if hs_f is not None else hs_b[1:]
        return [(hf1, hb1, T * H, wt, nw, rs, wmask, sg.N, sg.W, NL, H)
                for sg, (wt, nw, rs, wmask) in zip(segments, self.word_tables(segments, pk))]

    def _gemm_fold(self, a, w, T, N, K, fold, epi, vec, vec2, in_stats, eps, residual=None, out=None, out_stats=None):
        if out is None:
            out = torch.empty((T, N), dtype=torch.bfloat16, device=a.device)
        call("ruart_gemm_bf16_fold", ptr(a), a.stride(0), ptr(w), w.stride(0), T, N, K, fold, epi, ptr(vec), ptr(vec2),
             ptr(in_stats), eps, ptr(out), out.stride(0), ptr(residual), 0 if residual is None else residual.stride(0),
             ptr(out_stats), current_stream())
        return out

    def _encode_hidden_fold(self, pk, W):
        """bf16 encoder with every BertLayerNorm folded into its neighbours: 4 GEMMs + the attention per layer.
        Returns hs_raw bf16 [NL + 1, T, H] (rows BEFORE the embedding / each layer's output LayerNorm) and their
        partial sums [NL + 1, T, 8, 2]."""
        T, H, I, NL = pk["T"], self.H, self.I, self.n_layers
        dev = pk["ids"].device
        st = current_stream()
        hs = torch.empty((NL + 1, T, H), dtype=torch.bfloat16, device=dev)
        stats = torch.empty((NL + 1, T, 8, 2), dtype=torch.float32, device=dev)
        call("ruart_bert_embed_raw", ptr(pk["ids"]), ptr(pk["pos"]), ptr(W["word"]), ptr(W["pos"]), ptr(W["type"]),
             T, H, ptr(hs[0]), ptr(stats[0]), st)
        eps0 = W["eps"]
        stats1 = torch.empty((T, 8, 2), dtype=torch.float32, device=dev)
        raw1 = torch.empty((T, H), dtype=torch.bfloat16, device=dev)
        scale = 1.0 / 8.0
        fuse = self.fuse_attn and max([sg["max_len"] for sg in pk["segments"]] + [0]) <= 128
        if fuse:
            # row tiles of whole sequences (<= 128 rows) + per-token sequence bounds, once per batch: normally
            # already queued on the packing stream (pack_begin(want_tiles=True))
            if pk.get("tiles") is not None:
                meta, bounds, ev = pk["tiles"]
                main = torch.cuda.current_stream(dev)
                main.wait_event(ev)
                meta.record_stream(main)
                bounds.record_stream(main)
            else:
                cap = 2 * T // 128 + 8
                meta = torch.empty(cap, dtype=torch.int32, device=dev)
                bounds = torch.empty((T, 2), dtype=torch.int32, device=dev)
                call("ruart_seq_tiles", ptr(pk["cu_seqlens"]), pk["cu_seqlens"].numel() - 1, ptr(meta), cap, ptr(bounds), st)
        for li, lw in enumerate(W["layers"]):
            eps_in = eps0 if li == 0 else W["layers"][li - 1]["eps"]
            ctx = torch.empty((T, H), dtype=torch.bfloat16, device=dev)
            if fuse:
                call("ruart_qkv_attention_fold", ptr(hs[li]), H, ptr(lw["wqkv_p"]), H, T, H, self.heads,
                     ptr(lw["cqkv_p"]), ptr(lw["sqkv_p"]), ptr(stats[li]), eps_in, ptr(meta), ptr(bounds), ptr(ctx), H, st)
            else:
                qkv = self._gemm_fold(hs[li], lw["wqkv_f"], T, 3 * H, H, 1, ops.EPI_BIAS, lw["cqkv"], lw["sqkv"],
                                      stats[li], eps_in)
                for s0, s1, mlen in pk["att_groups"]:
                    cu = pk["cu_seqlens"][s0:s1 + 1]
                    call("ruart_bert_attention", None, ptr(qkv), ptr(cu), s1 - s0, self.heads, scale, mlen, None,
                         ptr(ctx), 1, st)
            self._gemm_fold(ctx, lw["wo"], T, H, H, 2, ops.EPI_BIAS, lw["res1_b"], lw["res1_g"], stats[li], eps_in,
                            residual=hs[li], out=raw1, out_stats=stats1)
            ff = self._gemm_fold(raw1, lw["wi_f"], T, I, H, 1, ops.EPI_BIAS_GELU, lw["ci"], lw["si"], stats1, lw["eps"])
            self._gemm_fold(ff, lw["wd"], T, H, I, 2, ops.EPI_BIAS, lw["res2_b"], lw["res2_g"], stats1, lw["eps"],
                            residual=raw1, out=hs[li + 1], out_stats=stats[li + 1])
        pk["fold"] = {"stats": stats, "fused_attention": fuse}
        return pk, None, hs

    def encode_hidden(self, segments, pack_handle=None, allow_fold=True):
        """The encoder alone: returns (pack, hs_f32 or None, hs_bf16 or None) with hs [n_layers + 1, T, H]
        (index 0 = embedding output, 1.. = encoder layers).  With the folded LayerNorms (bf16 mode, T >= FOLD_MIN_T,
        allow_fold) hs_bf16 holds the rows BEFORE each LayerNorm and pack["fold"]["stats"] their partial sums —
        apply_sinks() is the only consumer that understands that form."""
        dev = segments[0].ids.device
        if dev.type != "cuda":
            raise RuntimeError("ruart_b200 BERT runs on CUDA only; there is no CPU fallback")
        W = self.prepare(dev)
        pk = self.pack_finish(pack_handle) if pack_handle is not None else self.pack(segments, self.fuse_attn)
        T, H, I, NL = pk["T"], self.H, self.I, self.n_layers
        st = current_stream()
        scale = 1.0 / 8.0
        # attention launches: adjacent segments of short sequences (<= 16 tokens: the paired-MMA
        # kernel) are merged into one call.  (Running the question segment on a side stream next to
        # them was measured: no gain, its 192 KB CTAs and the item kernel's CTAs exclude each other.)
        att_groups = []
        for sgm in pk["segments"]:
            if sgm["seq1"] == sgm["seq0"]:
                continue
            short = sgm["max_len"] <= 16
            if att_groups and short and att_groups[-1][2] <= 16 and att_groups[-1][1] == sgm["seq0"]:
                g = att_groups[-1]
                att_groups[-1] = (g[0], sgm["seq1"], max(g[2], sgm["max_len"]))
            else:
                att_groups.append((sgm["seq0"], sgm["seq1"], sgm["max_len"]))
        pk["att_groups"] = att_groups
        if self.fold and allow_fold and T >= self.FOLD_MIN_T and self.gelu_mode == 2:
            return self._encode_hidden_fold(pk, W)
        fp32 = self.mode != "bf16"            # fp32 activations + split GEMM operands ("fp32" and "bf16x2")
        keep32 = fp32 or self.residual_fp32   # fp32 copy of the residual stream
        parts = self.parts
        # per-layer outputs: fp32 [NL, T, H] when an fp32 stream exists, else bf16 [NL, T, H]
        hs_f = torch.empty((NL + 1, T, H), dtype=torch.float32, device=dev) if keep32 else None
        if fp32:
            hs_b = None  # split operands are transient per layer
        else:
            hs_b = torch.empty((NL + 1, T, H), dtype=torch.bfloat16, device=dev)

        def layer_bufs(i):
            f = hs_f[i] if keep32 else None
            b = torch.empty((T, parts * H), dtype=torch.bfloat16, device=dev) if fp32 else hs_b[i]
            return f, b

        h_f, h_b = layer_bufs(0)
        call("ruart_bert_embed_ln", ptr(pk["ids"]), ptr(pk["pos"]), ptr(W["word"]), ptr(W["pos"]),
             ptr(W["type"]), ptr(W["eg"]), ptr(W["eb"]), W["eps"], T, H, ptr(h_f), ptr(h_b), parts, st)
        for li, lw in enumerate(W["layers"]):
            q_f, q_b = self._gemm(h_b, lw["wqkv"], lw["bqkv"], 3 * H, H, ops.EPI_BIAS, "act")
            ctx = torch.empty((T, parts * H), dtype=torch.bfloat16, device=dev)
            for s0, s1, mlen in att_groups:
                cu = pk["cu_seqlens"][s0:s1 + 1]
                call("ruart_bert_attention", ptr(q_f), ptr(q_b), ptr(cu), s1 - s0, self.heads, scale,
                     mlen, None, ptr(ctx), parts, st)
            fuse_res = not keep32  # bf16 residual stream: the GEMM epilogue adds it
            a_f, a_b = self._gemm(ctx, lw["wo"], lw["bo"], H, H, ops.EPI_BIAS, "act",
                                  residual=h_b if fuse_res else None)
            h1_f = torch.empty((T, H), dtype=torch.float32, device=dev) if keep32 else None
            h1_b = torch.empty((T, parts * H), dtype=torch.bfloat16, device=dev)
            call("ruart_add_layernorm", ptr(a_f), ptr(a_b), ptr(h_f), None,
                 ptr(lw["g1"]), ptr(lw["b1"]), lw["eps"], T, H, ptr(h1_f), ptr(h1_b), parts, st)
            _, ff = self._gemm(h1_b, lw["wi"], lw["bi"], I, H, ops.EPI_BIAS_GELU, "split",
                               fast_gelu=self.gelu_mode if not fp32 else 0)
            d_f, d_b = self._gemm(ff, lw["wd"], lw["bd"], H, I, ops.EPI_BIAS, "act",
                                  residual=h1_b if fuse_res else None)
            h_f, h_b = layer_bufs(li + 1)
            call("ruart_add_layernorm", ptr(d_f), ptr(d_b), ptr(h1_f), None,
                 ptr(lw["g2"]), ptr(lw["b2"]), lw["eps"], T, H, ptr(h_f), ptr(h_b), parts, st)
        return pk, hs_f, hs_b
