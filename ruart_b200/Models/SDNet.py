"""Drop-in for the reference's Models/SDNet.py: `SDNet(opt, embedding)` with the same constructor
contract, attribute / state_dict names and `forward(q_list, ocr_list, od_list) -> (score_s,
att_score)`, whose arithmetic runs on the sm_100a kernels of libruart_b200.so.

Scope: the inference path of the shipped `conf`: q embedding
glove,pos,ent,bert; ocr/od embedding fasttext,pos,ent,bert; PRE_ALIGN_befor_rnn; LN;
position_mod qk+; pos_att_merge_mod cat; useES (ES_using_way as_ocr); label_no_answer; mask_score.
Other option combinations raise NotImplementedError at construction / forward.

How it differs from the reference's forward (same results, different schedule):
  * one packed, pad-free BERT pass for question + OCR + OD tokens; the subword mean and the learned
    layer sum (Bert.py:149-165, SDNet.py:573-583) are accumulated layer by layer straight into
    the [word300 | bert768 | pos12 | ent8 | prealign300] concat buffers;
  * the host loops of SDNet.py:300-318 and :498-550 become index tensors built once per batch
    with numpy from `num_cnt` / `len_cnt` (ruart_b200/host_index.py; optionally already in the
    collate function, Utils/collate.py), consumed by gather kernels;
  * `multi2one` only runs the real word steps (the LSTM is causal and only step len-1 is read,
    SDNet.py:304,310) and scatters straight into the slot tensors;
  * the NaN asserts (Layers.py:169,290,430,462,467) are one device flag checked once per forward.
"""
import logging

import torch
import torch.nn as nn

from .. import autograd_ops as A
from .. import host_index, ops
from .. import sdnet_ops as K
from .._lib import current_stream, ptr
from ..ops import call
from . import Layers
from ..bert_engine import BertEngine
from .Bert.Bert import Bert
from .Layers import (Attention, DeepAttention, GetFinalScores, LinearSelfAttn, RNN_from_opt, dropout,
                     set_dropout_prob, set_seq_dropout)

log = logging.getLogger(__name__)

_REQUIRED = ("PRE_ALIGN", "PRE_ALIGN_befor_rnn", "BERT", "BERT_LINEAR_COMBINE", "useES", "label_no_answer",
             "mask_score", "position_dim", "GLOVE", "FastText")
_UNSUPPORTED = ("img_feature", "fixed_answers", "ModelParallel", "label_yesno", "no_Context_Self_Attention",
                "no_DeepAttention", "PRE_ALIGN_after_rnn", "BERT_LARGE_only")


def _pos_ent_sizes(opt):
    """len(POS), len(ENT): the reference takes them from spaCy at import (Utils/CoQAUtils.py:31-32)."""
    if "pos_size" in opt and "ent_size" in opt:
        return int(opt["pos_size"]), int(opt["ent_size"])
    try:  # running inside the reference tree: use its tables
        from Utils.CoQAUtils import ENT, POS
        return len(POS), len(ENT)
    except Exception as e:
        raise RuntimeError("opt['pos_size'] / opt['ent_size'] are required when Utils.CoQAUtils (spaCy) is "
                           "not importable: %r" % (e,))


class SDNet(nn.Module):
    def __init__(self, opt, embedding):
        super(SDNet, self).__init__()
        print('SDNet model\n')
        self.opt = opt
        for k in _REQUIRED:
            if k not in opt:
                raise NotImplementedError("ruart_b200.SDNet implements the shipped conf; option %s is required" % k)
        for k in _UNSUPPORTED:
            if k in opt:
                raise NotImplementedError("ruart_b200.SDNet: option %s is outside the implemented path" % k)
        if opt['position_mod'] != 'qk+' or opt['pos_att_merge_mod'] != 'cat' or opt.get('ES_using_way') != 'as_ocr':
            raise NotImplementedError("only position_mod qk+, pos_att_merge_mod cat, ES_using_way as_ocr")
        q_names, ocr_names = set(opt['q_embedding'].split(',')), set(opt['ocr_embedding'].split(','))
        if q_names - {'phoc'} != {'glove', 'pos', 'ent', 'bert'} or \
                ocr_names - {'phoc'} != {'fasttext', 'pos', 'ent', 'bert'}:
            raise NotImplementedError("only q_embedding glove,pos,ent,bert / ocr_embedding fasttext,pos,ent,bert "
                                      "(each optionally with phoc)")
        if 'phoc' in (q_names | ocr_names) and 'PHOC' not in opt:
            raise KeyError("'phoc' in q_embedding / ocr_embedding needs the PHOC option (SDNet.py:51-55)")
        self.vocab_dim = 300
        self.use_cuda = (opt['cuda'] == True)
        self.q_embedding = opt['q_embedding'].split(',')
        self.ocr_embedding = opt['ocr_embedding'].split(',')
        self.LN_flag = 'LN' in opt
        self.LN = 'LN' in opt
        self.drop_emb = False
        set_dropout_prob(0.0 if 'DROPOUT' not in opt else float(opt['DROPOUT']))
        set_seq_dropout('VARIATIONAL_DROPOUT' in opt)
        # split parts of the SDNet-stack GEMM operands: fp32-grade (3 parts, 6 products) when BERT runs
        # in fp32 mode, 2 parts (3 products, ~2^-16 relative) next to a bf16 BERT whose own error is
        # 1e-4..1e-2; override with opt['SDNET_precision'].
        default_prec = 'fp32' if opt.get('BERT_precision', 'bf16') == 'fp32' else 'bf16x2'
        self.sdnet_parts = {'fp32': 3, 'bf16x2': 2, 'bf16': 1}[opt.get('SDNET_precision', default_prec)]
        Layers.set_sdnet_precision(self.sdnet_parts)

        self.vocab_size = int(opt['vocab_size'])
        if 'PHOC' in opt:  # registered first, like the reference (SDNet.py:51-55), so state_dict order matches
            self.phoc_dim = int(opt['phoc_dim'])
            if self.phoc_dim != ops.PHOC_DIM:
                raise NotImplementedError("phoc_dim must be %d" % ops.PHOC_DIM)
            self.phoc_embed = nn.Embedding(self.vocab_size, self.phoc_dim, padding_idx=1)
            self.phoc_embed.weight.data = embedding['phoc_embedding']
        self.phoc_q = self.phoc_dim if 'phoc' in self.q_embedding else 0
        self.phoc_x = self.phoc_dim if 'phoc' in self.ocr_embedding else 0
        self.fast_dim = int(opt['fast_dim'])
        self.glove_dim = int(opt['glove_dim'])
        self.fast_embed = nn.Embedding(self.vocab_size, self.fast_dim, padding_idx=1)
        self.fast_embed.weight.data = embedding['fast_embedding']
        self.glove_embed = nn.Embedding(self.vocab_size, self.glove_dim, padding_idx=1)
        self.glove_embed.weight.data = embedding['glove_embedding']
        if 'TUNE_PARTIAL' in opt:
            print('TUNE_PARTIAL')
            self.fixed_embedding_fast = embedding['fast_embedding'][opt['tune_partial']:]
            self.fixed_embedding_glove = embedding['glove_embedding'][opt['tune_partial']:]
        else:
            self.fast_embed.weight.requires_grad = False
            self.glove_embed.weight.requires_grad = False

        print('Using BERT')
        self.Bert = Bert(opt)
        if 'LOCK_BERT' in opt:
            print('Lock BERT\'s weights')
            for p in self.Bert.parameters():
                p.requires_grad = False
        bert_dim = self.Bert.bert_dim
        bert_layers = self.Bert.bert_layer
        print('BERT dim:', bert_dim, 'BERT_LAYERS:', bert_layers)
        self.alphaBERT = nn.Parameter(torch.Tensor(bert_layers), requires_grad=True)
        self.gammaBERT = nn.Parameter(torch.Tensor(1, 1), requires_grad=True)
        torch.nn.init.constant_(self.alphaBERT, 1.0)
        torch.nn.init.constant_(self.gammaBERT, 1.0)

        pos_dim, ent_dim = opt['pos_dim'], opt['ent_dim']
        n_pos, n_ent = _pos_ent_sizes(opt)
        x_input_size = self.phoc_x + self.fast_dim + bert_dim + self.vocab_dim + pos_dim + ent_dim
        ques_input_size = self.phoc_q + self.glove_dim + bert_dim + pos_dim + ent_dim
        self.pre_align = Attention(self.vocab_dim, opt['prealign_hidden'], correlation_func=3, do_similarity=True)
        self.pos_embedding = nn.Embedding(n_pos, pos_dim)
        self.ent_embedding = nn.Embedding(n_ent, ent_dim)
        print('Initially, the vector_sizes [ocr, query] are', x_input_size, ques_input_size)
        self.x_input_size, self.ques_input_size, self.bert_dim = x_input_size, ques_input_size, bert_dim

        self.multi2one, multi2one_output_size = RNN_from_opt(
            x_input_size, opt['multi2one_hidden_size'], num_layers=1, concat_rnn=opt['concat_rnn'], add_feat=0,
            bidirectional=opt['multi2one_bidir'])
        if opt['multi2one_bidir']:
            raise NotImplementedError("multi2one_bidir True is outside the implemented path")
        self.multi2one_output_size = multi2one_output_size
        self.context_rnn, context_rnn_output_size = RNN_from_opt(
            multi2one_output_size, opt['hidden_size'], num_layers=opt['in_rnn_layers'],
            concat_rnn=opt['concat_rnn'], add_feat=0)
        self.ques_rnn, ques_rnn_output_size = RNN_from_opt(
            ques_input_size, opt['hidden_size'], num_layers=opt['in_rnn_layers'], concat_rnn=opt['concat_rnn'],
            add_feat=0)
        print('After Input LSTM, the vector_sizes [doc, query] are [', context_rnn_output_size,
              ques_rnn_output_size, '] *', opt['in_rnn_layers'])
        self.deep_attn = DeepAttention(opt, abstr_list_cnt=opt['in_rnn_layers'],
                                       deep_att_hidden_size_per_abstr=opt['deep_att_hidden_size_per_abstr'],
                                       correlation_func=3, word_hidden_size=multi2one_output_size)
        self.deep_attn_input_size = self.deep_attn.rnn_input_size
        self.deep_attn_output_size = self.deep_attn.output_size
        self.high_lvl_ques_rnn, high_lvl_ques_rnn_output_size = RNN_from_opt(
            ques_rnn_output_size * opt['in_rnn_layers'], opt['highlvl_hidden_size'],
            num_layers=opt['question_high_lvl_rnn_layers'], concat_rnn=True)
        self.after_deep_attn_size = self.deep_attn_output_size + self.deep_attn_input_size + multi2one_output_size
        self.self_attn_input_size = self.after_deep_attn_size
        self.highlvl_self_att = Attention(self.self_attn_input_size, opt['deep_att_hidden_size_per_abstr'],
                                          correlation_func=3)
        self.high_lvl_context_rnn, high_lvl_context_rnn_output_size = RNN_from_opt(
            self.deep_attn_output_size * 2, opt['highlvl_hidden_size'], num_layers=1, concat_rnn=False)
        context_final_size = high_lvl_context_rnn_output_size
        self.ques_self_attn = Attention(high_lvl_ques_rnn_output_size, opt['query_self_attn_hidden_size'],
                                        correlation_func=3)
        ques_final_size = high_lvl_ques_rnn_output_size
        self.od_ocr_attn = Attention(context_final_size, opt['hidden_size'], correlation_func=3, do_similarity=True)
        self.position_attn = Attention(opt['position_dim'], opt['hidden_size'], correlation_func=3,
                                       do_similarity=True)
        self.ques_merger = LinearSelfAttn(ques_final_size)
        ocr_final_size = context_final_size * 2
        self.get_answer = GetFinalScores(ocr_final_size, ques_final_size, yesno=False, no_answer=True, useES=True)
        # NaN guard (the reference's isnan asserts, Layers.py:169,290,430,462,467): True / 'sync' = read the device
        # flag at the end of every forward (one host sync, what predict()'s `.cpu()` would do anyway);
        # 'deferred' = copy the flag to pinned memory asynchronously and raise at the START of the next forward /
        # in check_pending(), so back-to-back forwards overlap (serving loops); False = no check
        self.check_nan = opt.get('CHECK_NAN', True)
        self._pending = []
        self.use_streams = bool(opt.get('USE_STREAMS', True))
        self._side = None
        self._warm_version = None
        log.debug('Network build successes')

    # host-side index building lives in ruart_b200/host_index.py (shared with Utils/collate.py)
    _item_index = staticmethod(host_index.item_index)

    def _phoc_channel(self, lst, dst, n_rows, errs):
        """PHOC columns of one item list (SDNet.py:441-446).  With `lst['phoc']` word ids: the
        reference's `[V, 604]` table lookup.  With `lst['phoc_chars']` (uint8) / `lst['phoc_offsets']`
        (int32 [n_rows+1]; see Utils.phoc.encode_tokens) the PHOC kernel writes the vectors of the
        word strings straight into the embedding buffer — no table, no out-of-vocabulary loss
        (SURVEY.md §8f-3).  Word slots with an empty string get zeros."""
        if 'phoc_chars' in lst:
            offs = lst['phoc_offsets']
            if offs.numel() != n_rows + 1 or offs.dtype != torch.int32:
                raise ValueError("phoc_offsets must be int32 [word slots + 1]")
            _, err = ops.phoc_batch(lst['phoc_chars'], offs, out=dst, check=False)
            errs.append(err)
        else:
            K.gather_rows(self.phoc_embed.weight.detach(), lst['phoc'].reshape(-1), dst, None, n_rows,
                          self.phoc_dim)

    # ------------------------------------------------------------------ forward
    phase_log = None  # set to a list to collect (label, seconds since forward start) with device syncs

    phase_events = None  # set to a list to collect (label, cuda event) on the main stream, no syncs
    _phase_stream = None

    def _phase(self, label):
        if self.phase_log is not None:
            import time
            torch.cuda.synchronize()
            self.phase_log.append((label, time.perf_counter()))
        if self.phase_events is not None:   # always on the MAIN stream (some phases are queued on side streams)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(self._phase_stream if self._phase_stream is not None else torch.cuda.current_stream())
            self.phase_events.append((label, ev))

    def check_pending(self, wait=True):
        """Raise what 'deferred' forwards found (NaN scores, stale bert_totals, unknown PHOC unigram).
        wait=False: only look at forwards whose flag copy has already landed (no host block) — what the next
        forward does on entry, so back-to-back forwards overlap; at most two forwards stay unchecked."""
        while self._pending:
            ev, host_flags, bert_pack, n_phoc = self._pending[0]
            if not wait and len(self._pending) <= 2 and not ev.query():
                return
            ev.synchronize()
            self._pending.pop(0)
            self._raise_flags(host_flags, bert_pack, n_phoc)

    @staticmethod
    def _raise_flags(host_flags, bert_pack, n_phoc):
        if int(host_flags[0]) != 0:
            raise AssertionError("NaN in answer scores (reference: assert torch.sum(torch.isnan(...)) == 0)")
        BertEngine.check_totals(bert_pack)
        for k in range(n_phoc):  # table-free PHOC channel: unknown unigram, like Utils/cphoc.c:45-50
            key = int(host_flags[1 + k])
            if key != -1:
                raise RuntimeError("Error: unigram %s is unknown" % chr(key & 0xFF))

    def forward(self, q_list, ocr_list, od_list, return_score=False):
        self.check_pending(wait=False)
        att_score = {} if return_score else None
        dev = ocr_list['fasttext'].device
        if dev.type != 'cuda':
            raise RuntimeError("ruart_b200.SDNet.forward needs CUDA inputs (ToCUDA, SDNetTrainer.py:208-230); "
                               "there is no CPU fallback")
        if Layers.grad_mode(self):
            # SDNetTrainer.update (SDNetTrainer.py:332-337): network.train(), autograd recording
            Layers.set_train_precision(max(2, self.sdnet_parts))
            return self._forward_differentiable(q_list, ocr_list, od_list), att_score
        if self.training and (Layers.dropout_p > 0 or (self.drop_emb and self.opt.get('dropout_emb', 0) > 0)):
            raise NotImplementedError("train-mode dropout needs autograd enabled (the differentiable form); "
                                      "for inference call .eval() and set drop_emb=False")
        opt = self.opt
        st = current_stream()
        Layers.set_sdnet_precision(self.sdnet_parts)
        f32 = dict(dtype=torch.float32, device=dev)
        B = len(ocr_list['num_cnt'])
        M, M_od = ocr_list['position'].size(1), od_list['position'].size(1)
        Wq, Wo, Wd = q_list['glove'].size(1), ocr_list['fasttext'].size(1), od_list['fasttext'].size(1)
        N_ocr, N_od = ocr_list['fasttext'].size(0), od_list['fasttext'].size(0)
        XD, QD, BD, VD = self.x_input_size, self.ques_input_size, self.bert_dim, self.vocab_dim
        pos_dim, ent_dim = opt['pos_dim'], opt['ent_dim']
        H = opt['hidden_size']

        self._phase('start')
        # token-length bookkeeping of the BERT pass starts on a side stream; its one host sync is
        # taken after the embedding gathers below have been queued
        def word_offsets(lst):  # the reference's nested list, or its CSR form from Utils.collate
            return lst['bert_offsets_csr'] if 'bert_offsets_csr' in lst else lst['bert_offsets']

        bert_segments = [
            (lst['bert'], lst['bert_mask'], word_offsets(lst), lst[wmask], lst.get('bert_totals'))
            for lst, wmask in ((q_list, 'glove_mask'), (ocr_list, 'fasttext_mask'), (od_list, 'fasttext_mask'))]
        pack_handle = self.Bert.pack_begin(bert_segments)
        # ---- embeddings: [(phoc |) word | bert | pos | ent (| prealign)]  (SDNet.py:439-493) ----
        q_in = torch.zeros((B, Wq, QD), **f32)
        # no zero-fill (2 GB at cfg-3): every column of every REAL word row is written below (word / pos /
        # ent gathers, BERT sink incl. explicit zeros for masked words, pre-align); pad-word rows of an
        # item are never read (multi2one consumes real word steps only)
        items_in = torch.empty((N_ocr * Wo + N_od * Wd, XD), **f32)
        ocr_in = items_in[:N_ocr * Wo].view(N_ocr, Wo, XD)
        od_in = items_in[N_ocr * Wo:].view(N_od, Wd, XD)
        q_word = torch.empty((B, Wq, VD), **f32)
        ocr_word = torch.empty((N_ocr, Wo, VD), **f32)
        od_word = torch.empty((N_od, Wd, VD), **f32)
        PQ, PX = self.phoc_q, self.phoc_x   # width of the leading PHOC channel (0 without it)
        phoc_errs = []

        def embed(lst, key, table, buf, raw, n_rows, P):
            if P:
                self._phoc_channel(lst, buf.view(n_rows, -1)[:, :P], n_rows, phoc_errs)
            c_pos, c_ent = P + VD + BD, P + VD + BD + pos_dim
            K.gather_rows(table.weight.detach(), lst[key].reshape(-1), buf[..., P:], None, n_rows, VD, dst2=raw)
            K.gather_rows(self.pos_embedding.weight.detach(), lst['pos'].reshape(-1), buf[..., c_pos:], None,
                          n_rows, pos_dim)
            K.gather_rows(self.ent_embedding.weight.detach(), lst['ent'].reshape(-1), buf[..., c_ent:], None,
                          n_rows, ent_dim)

        q_mask = K.as_u8(q_list['glove_mask'])
        # anything that invalidates the prepared-weight caches: parameter versions and the split width.  A forward
        # whose caches are cold (first call / weights changed) runs serially on the main stream so that no branch
        # reads a prepared weight another branch is still writing.
        main = torch.cuda.current_stream(dev)
        self._phase_stream = main
        plist = self.__dict__.get('_plist')
        if plist is None or self.__dict__.get('_plist_n') != len(self._modules):
            # Parameter objects are fixed after construction (state_dict loads / optimizers write in place); walking
            # the module tree costs ~1.3 ms of host time per forward
            plist = tuple(self.parameters())
            self.__dict__['_plist'] = plist
            self.__dict__['_plist_n'] = len(self._modules)
        ver = (sum(p._version for p in plist), self.sdnet_parts)
        concurrent = self.use_streams and self.phase_log is None and self._warm_version == ver
        if concurrent and self._side is None:
            # question branch (gates both context branches): high priority; s_emb: embeddings + pre-alignment,
            # which do not depend on BERT and run underneath its GEMMs
            self._side = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=-1),
                          torch.cuda.Stream(device=dev))
        s_emb = self._side[2] if concurrent else main
        fork_emb = torch.cuda.Event()
        fork_emb.record(main)      # inputs + the zero-filled q_in are ready; recorded BEFORE the encoder is queued

        # ---- BERT encoder: queued first on the main stream (compute-bound, ~3/4 of the step) -------
        hidden = self.Bert.encode_hidden(pack_handle)
        self._phase('bert_layers')

        # ---- embeddings, host indices and word-level pre-alignment on the side stream -------------
        s_emb.wait_event(fork_emb)
        with torch.cuda.stream(s_emb):
            embed(q_list, 'glove', self.glove_embed, q_in, q_word, B * Wq, PQ)
            embed(ocr_list, 'fasttext', self.fast_embed, ocr_in, ocr_word, N_ocr * Wo, PX)
            embed(od_list, 'fasttext', self.fast_embed, od_in, od_word, N_od * Wd, PX)
            q_list['glove_emb'] = q_word          # side effect of the reference (SDNet.py:449-450,458-459)
            ocr_list['fasttext_emb'] = ocr_word
            od_list['fasttext_emb'] = od_word
            self._phase('embed')
            # host indices (one upload each)
            plan = ocr_list.get('ruart_plan')   # precomputed by Utils.collate.attach_index_tensors, else here
            if plan is None:
                plan = host_index.forward_plan(ocr_list['num_cnt'], ocr_list['len_cnt'], od_list['num_cnt'],
                                               od_list['len_cnt'], Wo, Wd, M, M_od)
            if plan['key'] != (B, N_ocr, N_od, Wo, Wd, M, M_od):
                raise ValueError("num_cnt / len_cnt do not match the item rows of this batch")
            n_t, n_step_rows, n_all = plan['n_t'], plan['n_step_rows'], plan['n_items']
            i32_d = K.upload(plan['i32'], dev)
            i64_d = K.upload(plan['slots'] * self.multi2one_output_size, dev)
            masks_d = K.upload(plan['masks'], dev)
            for t in (i32_d, i64_d, masks_d):
                t.record_stream(main)
            cuts = plan['cuts']
            ocr_wsrc, ocr_wdst, od_wsrc, od_wdst, a_rows_d, last_d = [i32_d[cuts[i]:cuts[i + 1]] for i in range(6)]
            ocr_mask = masks_d[:B * M].view(B, M)
            od_mask = masks_d[B * M:].view(B, M_od)
            self._phase('host_index')
            # word-level pre-alignment (SDNet.py:495-551)
            c_pre = PX + VD + BD + pos_dim + ent_dim
            p2_cache = {}
            for k, word, wsrc, wdst, buf in ((0, ocr_word, ocr_wsrc, ocr_wdst, ocr_in),
                                             (1, od_word, od_wsrc, od_wdst, od_in)):
                T_max, n_words = plan['T_max'][k], plan['total_words'][k]
                packed = torch.zeros((B, T_max, VD), **f32)
                K.gather_rows(word, wsrc, packed, wdst, n_words, VD)
                att = self.pre_align(packed, q_word, q_mask, p2_cache=p2_cache)
                K.gather_rows(att, wdst, buf[..., c_pre:], wsrc, n_words, VD)
            ev_emb = torch.cuda.Event()
            ev_emb.record(s_emb)

        # ---- BERT tail: subword mean + layer sum into the concat buffers (host: flatten the offset lists) ----
        bert_pack = self.Bert.mix_into(pack_handle, hidden,
                                       [(q_in, QD, PQ + VD), (ocr_in, XD, PX + VD), (od_in, XD, PX + VD)],
                                       self.alphaBERT, self.gammaBERT)
        del hidden
        self._phase('bert')
        main.wait_event(ev_emb)

        self._phase('prealign')
        # ---- multi2one: real word steps only, last step -> slot (SDNet.py:270-271,300-318) ----
        HS = self.multi2one_output_size
        slots = torch.zeros((B * M + B * M_od, HS), **f32)
        m2o = self.multi2one
        w_ih, w_hh, b_ih, b_hh = m2o._dir_params(0)
        mp = self.sdnet_parts
        a_sp, Kp_in = K.split_act(items_in, mp, row_idx=a_rows_d, n_rows=n_step_rows)
        wi, _ = K.prep_weight(m2o, (0, "w_ih"), w_ih, mp)
        wh, Kp_h = K.prep_weight(m2o, (0, "w_hh"), w_hh, mp)
        bias = K.prep_vector(m2o, (0, "bias"), lambda: b_ih[0] + b_hh[0], b_ih + b_hh)
        gx = torch.empty((n_step_rows, 4 * HS), **f32)
        K.linear(a_sp, Kp_in, wi, n_step_rows, 4 * HS, mp, gx, epi=ops.EPI_BIAS, bias=bias)
        c_state = torch.zeros((n_all, HS), **f32)
        h_split = torch.zeros((n_all, mp * Kp_h), dtype=torch.bfloat16, device=dev)
        gh = torch.empty((n_all, 4 * HS), **f32)
        row0 = 0
        for t, n in enumerate(n_t):
            n = int(n)
            if t > 0:
                K.linear(h_split, Kp_h, wh, n, 4 * HS, mp, gh)
            call("ruart_lstm_cell", gx.data_ptr() + row0 * 4 * HS * 4, None, ptr(gh) if t > 0 else None,
                 ptr(c_state), ptr(h_split), mp, Kp_h, HS, n, ptr(last_d), t, ptr(i64_d), ptr(slots), st)
            row0 += n
        ocr_x = slots[:B * M].view(B, M, HS)
        od_x = slots[B * M:].view(B, M_od, HS)

        self._phase('multi2one')
        # ---- the OCR, OD and question branches are independent until they meet: they run on three
        # streams (main = OCR, side = OD, question).  Temporaries are allocated and consumed inside
        # their branch's stream; tensors that cross branches live until the joins below.  A forward
        # whose weight caches are cold (first call / weights changed) runs serially on the main
        # stream so that no branch reads a prepared weight another branch is still writing.
        L_in = opt['in_rnn_layers']
        if concurrent:
            s_od, s_q = self._side[0], self._side[1]
            fork = torch.cuda.Event()
            fork.record(main)
            s_od.wait_event(fork)
            s_q.wait_event(fork)
        else:
            s_od = s_q = main

        def encode(rnn, x, n_layers):
            outs, cur = [], x
            for i in range(n_layers):
                cur = rnn.run_layer(i, cur, LN=True)
                outs.append(cur)
            return outs

        def context_branch(x, layers, mask, Mx):
            # deep inter-attention + context self-attention (SDNet.py:376-390)
            after, before = self.deep_attn([x], layers, [q_word], q_layers, mask, q_mask, return_bef_rnn=True,
                                           x2_proj=q_proj)
            # s_in = cat(after, before-rnn, x) and cat(after, self-attention output) are never materialised:
            # `before` is the list of its pieces and the GEMM operands are built from the pieces (K.split_concat)
            s_in = [after] + list(before) + [x]
            s_out = self.highlvl_self_att(s_in, s_in, mask, x3=after)
            return self.high_lvl_context_rnn.run_layer(0, [after, s_out], LN=True)

        # encoders with whole-tensor LN (SDNet.py:338-350)
        with torch.cuda.stream(s_q):
            q_layers = encode(self.ques_rnn, q_in, L_in)
            q_high = encode(self.high_lvl_ques_rnn, list(q_layers), opt['question_high_lvl_rnn_layers'])[-1]
            q_layers = q_layers + [q_high]
            q_proj = self.deep_attn.project_x2([q_word], q_layers)  # shared by the OCR and OD branches
            ev_q = torch.cuda.Event()
            ev_q.record(s_q)
        ocr_layers = encode(self.context_rnn, ocr_x, L_in)   # critical path: queued before the OD branch
        with torch.cuda.stream(s_od):
            od_layers = encode(self.context_rnn, od_x, L_in)
        self._phase('encoders')
        main.wait_event(ev_q)
        s_od.wait_event(ev_q)
        ocr_high = context_branch(ocr_x, ocr_layers, ocr_mask, M)
        with torch.cuda.stream(s_od):
            od_high = context_branch(od_x, od_layers, od_mask, M_od)
            ev_od = torch.cuda.Event()
            ev_od.record(s_od)
        with torch.cuda.stream(s_q):
            # question summary (SDNet.py:411-415)
            q_final = self.ques_self_attn(q_high, q_high, q_mask)
            q_merged = self.ques_merger.pooled(q_final, q_mask)
            ev_q2 = torch.cuda.Event()
            ev_q2.record(s_q)
        self._phase('deep_self_attn')
        main.wait_event(ev_od)
        # ---- OD <-> OCR + position attention (SDNet.py:393-405) -------------------------------
        CF = ocr_high.shape[2]
        ocr_final = torch.empty((B, M, 2 * CF), **f32)
        K.copy_cols(ocr_high, ocr_final[:, :, :CF])
        x_od_ocr = ocr_final[:, :, CF:]
        self.od_ocr_attn(ocr_high, od_high, od_mask, out=x_od_ocr)
        self.position_attn(ocr_list['position'].float(), od_list['position'].float(), od_mask, x3=od_high,
                           out=x_od_ocr, add_to_out=True)
        self._phase('od_ocr')
        main.wait_event(ev_q2)
        # ---- answer scores (SDNet.py:428-431) -------------------------------------------------
        nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        score_s = self.get_answer(ocr_final, q_merged, ocr_mask, opt['ES_ocr_len'], mask_flag='mask_score' in opt,
                                  nan_flag=nan_flag, want_logits=bool(opt.get('KEEP_LOGITS', False)))
        self._warm_version = ver
        self._phase('scores')
        if self.check_nan or phoc_errs:
            # ONE read-back of all device flags: [nan flag | PHOC error words] -> pinned host memory
            flags = torch.cat([nan_flag.to(torch.int64)] + [e.reshape(1) for e in phoc_errs]) if phoc_errs \
                else nan_flag.to(torch.int64)
            host_flags = torch.empty(flags.shape, dtype=torch.int64, pin_memory=True)
            host_flags.copy_(flags, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            if self.check_nan == 'deferred':
                self._pending.append((ev, host_flags, bert_pack, len(phoc_errs)))
            else:
                ev.synchronize()
                self._raise_flags(host_flags, bert_pack if self.check_nan else None, len(phoc_errs))
        return score_s, att_score

    def linear_sum(self, output, alpha, gamma):
        """sum_l output[l] * softmax(alpha)_l * gamma, then dropout_emb (SDNet.py:573-583).  The fused forward
        never materialises the per-layer tensors (Bert.encode_into); this is the API / differentiable form."""
        res = A.layer_mix(list(output), alpha, gamma)
        return dropout(res, p=self.opt['dropout_emb'], training=self.drop_emb)

    # ------------------------------------------------------------------ differentiable forward
    def _forward_differentiable(self, q_list, ocr_list, od_list):
        """SDNet.forward (SDNet.py:253-437) op by op on autograd Functions whose forward and backward are
        kernels of libruart_b200.so (ruart_b200/autograd_ops.py) — what `SDNetTrainer.update` differentiates
        (SURVEY.md §8 a-19).  BERT is locked (SDNet.py:91-94): the packed encoder runs as in inference and stays
        in eval mode (the reference's `network.train()` also switches the locked BERT's dropout on,
        SDNetTrainer.py:332 vs Bert.py:43 — not reproduced: a frozen encoder is kept deterministic); only
        alphaBERT / gammaBERT receive gradients from it."""
        opt = self.opt
        if 'LOCK_BERT' not in opt:
            raise NotImplementedError("training needs LOCK_BERT (the BERT encoder has no backward kernels)")
        if self.phoc_q or self.phoc_x:
            raise NotImplementedError("the PHOC channel is inference-only")
        P = Layers.train_parts
        dev = ocr_list['fasttext'].device
        B = len(ocr_list['num_cnt'])
        M, M_od = ocr_list['position'].size(1), od_list['position'].size(1)
        Wo, Wd = ocr_list['fasttext'].size(1), od_list['fasttext'].size(1)
        N_ocr, N_od = ocr_list['fasttext'].size(0), od_list['fasttext'].size(0)
        VD = self.vocab_dim
        p_emb = opt['dropout_emb'] if 'dropout_emb' in opt else 0

        # ---- locked BERT: packed encoder, hidden states kept for the layer-mix gradients --------
        def word_offsets(lst):
            return lst['bert_offsets_csr'] if 'bert_offsets_csr' in lst else lst['bert_offsets']

        lists = ((q_list, 'glove_mask'), (ocr_list, 'fasttext_mask'), (od_list, 'fasttext_mask'))
        with torch.no_grad():
            from ..bert_engine import Segment
            segs = [Segment(l['bert'], l['bert_mask'], word_offsets(l), l[m], l.get('bert_totals')) for l, m in lists]
            eng = self.Bert.engine()
            pk, hs_f, hs_b = eng.encode_hidden(segs, allow_fold=False)   # SubwordMixFn reads normalised rows
            packs = eng.subword_packs(segs, pk, hs_f, hs_b)

        def embed(lst, key, table, pack):   # get_embedding_from_list, SDNet.py:439-493
            word = A.embedding(lst[key], table.weight)
            bert = A.subword_mix(self.alphaBERT, self.gammaBERT, pack)
            bert = dropout(bert, p=p_emb, training=self.drop_emb)
            parts = [dropout(word, p=p_emb, training=self.drop_emb), bert,
                     A.embedding(lst['pos'], self.pos_embedding.weight),
                     A.embedding(lst['ent'], self.ent_embedding.weight)]
            return word, torch.cat(parts, -1)

        q_word, q_in = embed(q_list, 'glove', self.glove_embed, packs[0])
        ocr_word, ocr_in = embed(ocr_list, 'fasttext', self.fast_embed, packs[1])
        od_word, od_in = embed(od_list, 'fasttext', self.fast_embed, packs[2])
        q_list['glove_emb'] = q_word
        ocr_list['fasttext_emb'] = ocr_word
        od_list['fasttext_emb'] = od_word
        q_mask = K.as_u8(q_list['glove_mask'])

        plan = ocr_list.get('ruart_plan')
        if plan is None:
            plan = host_index.forward_plan(ocr_list['num_cnt'], ocr_list['len_cnt'], od_list['num_cnt'],
                                           od_list['len_cnt'], Wo, Wd, M, M_od)
        if plan['key'] != (B, N_ocr, N_od, Wo, Wd, M, M_od):
            raise ValueError("num_cnt / len_cnt do not match the item rows of this batch")
        i32_d = K.upload(plan['i32'], dev)
        i64_d = K.upload(plan['slots'] * self.multi2one_output_size, dev)
        masks_d = K.upload(plan['masks'], dev)
        cuts = plan['cuts']
        ocr_wsrc, ocr_wdst, od_wsrc, od_wdst, a_rows_d, last_d = [i32_d[cuts[i]:cuts[i + 1]] for i in range(6)]
        ocr_mask = masks_d[:B * M].view(B, M)
        od_mask = masks_d[B * M:].view(B, M_od)

        # ---- word-level pre-alignment before the item RNN (SDNet.py:265-268,495-551) -------------
        def prealign(word, wsrc, wdst, k, n_items, W):
            T_max = plan['T_max'][k]
            packed = A.permute_rows(word.reshape(-1, VD), wsrc, wdst, B * T_max).view(B, T_max, VD)
            att = self.pre_align(packed, q_word, q_mask)
            return A.permute_rows(att.reshape(-1, VD), wdst, wsrc, n_items * W).view(n_items, W, VD)

        ocr_in = torch.cat([ocr_in, prealign(ocr_word, ocr_wsrc, ocr_wdst, 0, N_ocr, Wo)], -1)
        od_in = torch.cat([od_in, prealign(od_word, od_wsrc, od_wdst, 1, N_od, Wd)], -1)

        # ---- multi2one over the real word steps, last step -> slot (SDNet.py:270-271,300-318) -----
        XD = ocr_in.shape[-1]
        if Layers.dropout_p > 0:   # StackedBRNN.forward's input dropout (Layers.py:163-164), one call per list
            ocr_in = dropout(ocr_in, p=Layers.dropout_p, training=self.training)
            od_in = dropout(od_in, p=Layers.dropout_p, training=self.training)
        items_in = torch.cat([ocr_in.reshape(-1, XD), od_in.reshape(-1, XD)], 0)
        slots = A.multi2one(items_in, self.multi2one.rnns[0],
                            (a_rows_d, last_d, i64_d, plan['n_t'], B * M + B * M_od), parts=P)
        HS = self.multi2one_output_size
        ocr_x = slots[:B * M].view(B, M, HS)
        od_x = slots[B * M:].view(B, M_od, HS)

        # ---- encoders, deep attention, self attention (SDNet.py:338-390) ----------------------
        _, ocr_layers = self.context_rnn(ocr_x, ocr_mask, return_list=True, LN=True)
        _, q_layers = self.ques_rnn(q_in, q_mask, return_list=True, LN=True)
        _, od_layers = self.context_rnn(od_x, od_mask, return_list=True, LN=True)
        q_high = self.high_lvl_ques_rnn(torch.cat(q_layers, 2), q_mask, LN=True)
        q_layers = q_layers + [q_high]
        ocr_after, ocr_before = self.deep_attn([ocr_x], ocr_layers, [q_word], q_layers, ocr_mask, q_mask,
                                               return_bef_rnn=True)
        od_after, od_before = self.deep_attn([od_x], od_layers, [q_word], q_layers, od_mask, q_mask,
                                             return_bef_rnn=True)

        def self_attend(after, before, x, mask):
            s_in = torch.cat([after, before, x], 2)
            s_out = self.highlvl_self_att(s_in, s_in, mask, x3=after)
            return self.high_lvl_context_rnn(torch.cat([after, s_out], 2), mask, LN=True)

        ocr_high = self_attend(ocr_after, ocr_before, ocr_x, ocr_mask)
        od_high = self_attend(od_after, od_before, od_x, od_mask)
        # ---- OD <-> OCR + position attention (SDNet.py:393-405) -------------------------------
        x_od_ocr = self.od_ocr_attn(ocr_high, od_high, od_mask)
        pos_att = self.position_attn(ocr_list['position'].float(), od_list['position'].float(), od_mask, x3=od_high)
        ocr_final = torch.cat([ocr_high, A.add(x_od_ocr, pos_att)], 2)
        # ---- question summary + scores (SDNet.py:408-431) -------------------------------------
        q_final = self.ques_self_attn(q_high, q_high, q_mask)
        q_merged = Layers.weighted_avg(q_final, self.ques_merger(q_final, q_mask))
        score_s = self.get_answer(ocr_final, q_merged, ocr_mask, opt['ES_ocr_len'], mask_flag='mask_score' in opt)
        return score_s
