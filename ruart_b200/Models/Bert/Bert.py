"""Drop-in for the reference's Models/Bert/Bert.py (`Bert(opt).forward(...)`), running the
encoder as packed sm_100a kernels (ruart_b200.bert_engine) instead of torch ops.

Same constructor contract (Bert.py:15-45): reads opt['BERT_LINEAR_COMBINE'], opt['BERT_LARGE'],
os.path.join(opt['datadir'], opt['BERT_model_file']) holding bert_config.json +
pytorch_model.bin; the model is put on the GPU and in eval mode.  opt['BERT_MAX_BatchSize'] is
accepted and ignored: rows are independent and nothing of size [N, L, 12*768] is materialised, so
chunking is never needed (it is numerically neutral in the reference, Bert.py:65-85).

Extra (not in the reference): opt['BERT_precision'] in {'bf16' (default), 'bf16x2', 'fp32'}, and
`encode_into` — the fused entry SDNet uses (subword mean + learned layer sum written straight
into the embedding concat buffer).
"""
import os

import torch
import torch.nn as nn

from ...bert_engine import BertEngine, Segment
from .modeling import BertConfig, BertModel


class Bert(nn.Module):
    def __init__(self, opt):
        super(Bert, self).__init__()
        print('Loading BERT model...')
        self.BERT_MAX_LEN = 512
        self.linear_combine = 'BERT_LINEAR_COMBINE' in opt
        self.BERT_MAX_BS = opt.get('BERT_MAX_BatchSize')
        large = 'BERT_LARGE' in opt
        model_file = os.path.join(opt.get('datadir', ''),
                                  opt['BERT_large_model_file'] if large else opt.get('BERT_model_file', ''))
        if model_file and os.path.isfile(os.path.join(model_file, 'bert_config.json')):
            self.bert_model = BertModel.from_pretrained(model_file)
        else:
            # no checkpoint directory (synthetic / benchmark use): random init of the right shape
            cfg = BertConfig(30522, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                             intermediate_size=4096) if large else BertConfig(30522)
            if 'BERT_num_layers' in opt:
                cfg.num_hidden_layers = int(opt['BERT_num_layers'])
            self.bert_model = BertModel(cfg)
        self.bert_dim = self.bert_model.config.hidden_size
        self.bert_layer = self.bert_model.config.num_hidden_layers
        self.precision = opt.get('BERT_precision', 'bf16')
        self.residual_fp32 = bool(opt.get('BERT_residual_fp32', False))
        if torch.cuda.is_available():
            self.bert_model.cuda()
        self.bert_model.eval()
        self._engine = None
        print('Finished loading')

    def engine(self):
        if self._engine is None or self._engine.mode != self.precision \
                or self._engine.residual_fp32 != self.residual_fp32:
            self._engine = BertEngine(self.bert_model, self.precision, self.residual_fp32)
        return self._engine

    def forward(self, x_bert, x_bert_mask, x_bert_offset, x_mask, device=None):
        """Reference signature (Bert.py:56-90).  Returns the list of per-layer word tensors
        [N, W, bert_dim] (BERT_LINEAR_COMBINE) or the last layer's tensor."""
        if x_bert_offset is None:
            raise NotImplementedError("x_bert_offset=None (raw wordpiece outputs) is not on RUArt's path")
        N, W = x_mask.shape
        dev = x_bert.device
        seg = Segment(x_bert, x_bert_mask, x_bert_offset, x_mask)
        outs = [torch.zeros((N, W, self.bert_dim), dtype=torch.float32, device=dev)
                for _ in range(self.bert_layer)]
        sinks = [[(o, self.bert_dim, 0) for o in outs]]
        self.engine().encode([seg], sinks, alpha=None, gamma=None)
        if device is not None:
            outs = [o.to(device) for o in outs]
        return outs if self.linear_combine else outs[-1]

    def pack_begin(self, segments):
        """Start the token packing of `segments` on a side stream (see BertEngine.pack_begin)."""
        segs = [Segment(*s) for s in segments]
        eng = self.engine()
        return eng.pack_begin(segs, want_tiles=eng.fuse_attn)

    def encode_into(self, segments, sinks, alpha, gamma, pack_handle=None):
        """Fused path: segments = [(ids, mask, offsets, word_mask)], sinks = [(dst, stride, col)].
        dst[item, j, col:col+dim] = sum_l softmax(alpha)_l * gamma * mean_subwords(layer_l)."""
        segs = pack_handle["segments"] if pack_handle is not None else [Segment(*s) for s in segments]
        return self.engine().encode(segs, sinks, alpha=alpha.detach().float().contiguous(),
                                    gamma=gamma.detach().float().contiguous(), pack_handle=pack_handle)

    def encode_hidden(self, pack_handle):
        """First half of encode_into: queue the packed encoder; returns (pack, hs_f32 | None, hs_bf16 | None)."""
        return self.engine().encode_hidden(pack_handle["segments"], pack_handle)

    def mix_into(self, pack_handle, hidden, sinks, alpha, gamma):
        """Second half: subword mean + learned layer sum of `hidden` (from encode_hidden) into the sinks."""
        pk, hs_f, hs_b = hidden
        self.engine().apply_sinks(pack_handle["segments"], pk, hs_f, hs_b, sinks,
                                  alpha=alpha.detach().float().contiguous(), gamma=gamma.detach().float().contiguous())
        return pk
