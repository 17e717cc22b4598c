"""Parameter containers with the reference's BERT state_dict layout.

The reference vendors pytorch-pretrained-BERT (Models/Bert/modeling.py).  Here the modules only
HOLD parameters under the same names (so `bert_config.json` + `pytorch_model.bin` checkpoints and
SDNet checkpoints load unchanged, Models/Bert/modeling.py:497-521); the arithmetic lives in the
sm_100a kernels driven by ruart_b200.bert_engine.BertEngine.  There is no torch forward here.

Name map (reference file:line):
  embeddings.{word,position,token_type}_embeddings.weight, embeddings.LayerNorm.{gamma,beta}  :171-199
  encoder.layer.N.attention.self.{query,key,value}.{weight,bias}                             :202-227
  encoder.layer.N.attention.output.dense.{weight,bias}, .LayerNorm.{gamma,beta}              :253-264
  encoder.layer.N.intermediate.dense.{weight,bias}                                           :279-289
  encoder.layer.N.output.dense.{weight,bias}, .LayerNorm.{gamma,beta}                        :292-303
  pooler.dense.{weight,bias}  (kept for checkpoint compatibility; its output is discarded by
                               Models/Bert/Bert.py:136, so it is never evaluated)            :337-349
"""
import json
import os

import torch
import torch.nn as nn


class BertConfig(object):
    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02):
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r") as f:
                self.__dict__.update(json.load(f))
        else:
            self.vocab_size = vocab_size_or_config_json_file
            self.hidden_size = hidden_size
            self.num_hidden_layers = num_hidden_layers
            self.num_attention_heads = num_attention_heads
            self.intermediate_size = intermediate_size
            self.hidden_act = hidden_act
            self.hidden_dropout_prob = hidden_dropout_prob
            self.attention_probs_dropout_prob = attention_probs_dropout_prob
            self.max_position_embeddings = max_position_embeddings
            self.type_vocab_size = type_vocab_size
            self.initializer_range = initializer_range

    @classmethod
    def from_json_file(cls, path):
        return cls(path)

    def to_dict(self):
        return dict(self.__dict__)


class BertLayerNorm(nn.Module):
    """gamma/beta holder; eps 1e-12 inside the sqrt (reference modeling.py:155-168)."""

    def __init__(self, hidden, eps=1e-12):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(hidden))
        self.beta = nn.Parameter(torch.zeros(hidden))
        self.variance_epsilon = eps


class _Holder(nn.Module):
    pass


def _linear(i, o):
    return nn.Linear(i, o)


class BertModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        H, I = config.hidden_size, config.intermediate_size
        emb = _Holder()
        emb.word_embeddings = nn.Embedding(config.vocab_size, H)
        emb.position_embeddings = nn.Embedding(config.max_position_embeddings, H)
        emb.token_type_embeddings = nn.Embedding(config.type_vocab_size, H)
        emb.LayerNorm = BertLayerNorm(H)
        self.embeddings = emb
        enc = _Holder()
        layers = []
        for _ in range(config.num_hidden_layers):
            lay = _Holder()
            att = _Holder()
            att.self = _Holder()
            att.self.query = _linear(H, H)
            att.self.key = _linear(H, H)
            att.self.value = _linear(H, H)
            att.output = _Holder()
            att.output.dense = _linear(H, H)
            att.output.LayerNorm = BertLayerNorm(H)
            lay.attention = att
            lay.intermediate = _Holder()
            lay.intermediate.dense = _linear(H, I)
            lay.output = _Holder()
            lay.output.dense = _linear(I, H)
            lay.output.LayerNorm = BertLayerNorm(H)
            layers.append(lay)
        enc.layer = nn.ModuleList(layers)
        self.encoder = enc
        self.pooler = _Holder()
        self.pooler.dense = _linear(H, H)
        self._init_weights()

    @torch.no_grad()
    def _init_weights(self):
        # same distributions as the reference's init_bert_weights (modeling.py:432-443)
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.normal_(0.0, std)
            elif isinstance(m, BertLayerNorm):
                m.beta.normal_(0.0, std)
                m.gamma.normal_(0.0, std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.zero_()

    @classmethod
    def from_pretrained(cls, model_dir):
        """Directory with bert_config.json + pytorch_model.bin whose keys carry the 'bert.' prefix
        (reference modeling.py:445-531)."""
        config = BertConfig.from_json_file(os.path.join(model_dir, "bert_config.json"))
        model = cls(config)
        sd = torch.load(os.path.join(model_dir, "pytorch_model.bin"), map_location="cpu")
        own = model.state_dict()
        has_prefix = any(k.startswith("bert.") for k in sd)
        picked = {}
        for k, v in sd.items():
            kk = k[5:] if (has_prefix and k.startswith("bert.")) else k
            if kk in own:
                picked[kk] = v
        model.load_state_dict(picked, strict=False)
        return model

    def forward(self, *a, **k):
        raise RuntimeError("ruart_b200 BertModel has no torch forward; use Models.Bert.Bert.Bert "
                           "(sm_100a kernels via ruart_b200.bert_engine)")
