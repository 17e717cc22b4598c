"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference (xiaojino/RUArt under /root/reference).

Only usable in the build container (the reference tree does not travel to the GPU box).  It is
used to (a) pin oracle/sdnet_oracle.py and (b) generate the golden fixtures under tests/golden/
(see oracle/gen_model_golden.py).  Nothing in ruart_b200/ imports this.

Shims, all outside the reference tree:
  (i)   stub modules `spacy`, `fasttext`, `h5py` — Models/SDNet.py:12 imports POS/ENT from
        Utils/CoQAUtils.py, which loads spaCy at import (Utils/GeneralUtils.py:7,13) and fastText
        (Utils/CoQAUtils.py:26); |POS|, |ENT| are fixed to synth.POS_SIZE / synth.ENT_SIZE;
  (ii)  on a CPU-only host `.cuda()` becomes the identity (hard-coded `.cuda()` calls at
        Models/SDNet.py:288-299, Models/Bert/Bert.py:42,173);
  (iii) a temp dir with bert_config.json + pytorch_model.bin ('bert.'-prefixed BertModel state);
  (iv)  `opt` from ruart_b200.synth.make_opt (the shipped conf restated) ;
  (v)   net.drop_emb = False, net.eval();
  (vi)  training-step goldens only (oracle/gen_train_golden.py): `fixed_embedding_*` cloned, which is what
        `network.cuda()` does implicitly on a GPU (see the comment there), and all dropout at 0.
"""
import contextlib
import json
import os
import sys
import tempfile
import types

import torch

from ruart_b200 import synth

REF_ROOT = os.environ.get("RUART_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "Models", "SDNet.py"))


def _install_stubs():
    if "spacy" not in sys.modules:
        spacy = types.ModuleType("spacy")

        class _Tagger:
            labels = tuple("P%d" % i for i in range(synth.POS_SIZE - 1))

        class _Entity:
            move_names = ["E%d" % i for i in range(synth.ENT_SIZE - 1)]

        class _NLP:
            tagger = _Tagger()
            entity = _Entity()

        spacy.load = lambda *a, **k: _NLP()
        sys.modules["spacy"] = spacy
    if "fasttext" not in sys.modules:
        ft = types.ModuleType("fasttext")
        ft.load_model = lambda *a, **k: None
        sys.modules["fasttext"] = ft
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")


@contextlib.contextmanager
def _cpu_cuda_identity():
    """(ii): make `.cuda()` a no-op when there is no GPU."""
    if torch.cuda.is_available():
        yield
        return
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def import_reference():
    """Import Models.SDNet etc. from the reference tree; returns the `Models.SDNet` module."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # our package mirrors the names Models/ and Utils/ *inside* ruart_b200/, so there is no clash
    import Models.SDNet as ref_sdnet  # noqa: E402
    return ref_sdnet


def build_reference(opt, embedding=None, seed=1033, bert_init="random", bert_layers=12, bert_dropout=None):
    """Construct the reference SDNet with weights from synth.fill_state_dict (same names as ours)."""
    ref_sdnet = import_reference()
    from Models.Bert.modeling import BertConfig, BertModel
    if embedding is None:
        embedding = synth.make_embedding(seed)
    opt = dict(opt)
    with tempfile.TemporaryDirectory() as tmp, _cpu_cuda_identity():
        cfg = BertConfig(synth.BERT_VOCAB, num_hidden_layers=bert_layers)
        if bert_dropout is not None:  # training-step goldens: deterministic
            cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = float(bert_dropout)
        with open(os.path.join(tmp, "bert_config.json"), "w") as f:
            json.dump(cfg.to_dict(), f)
        bm = BertModel(cfg)
        torch.save({"bert." + k: v for k, v in bm.state_dict().items()}, os.path.join(tmp, "pytorch_model.bin"))
        del bm
        opt["datadir"] = ""
        opt["BERT_model_file"] = tmp
        net = ref_sdnet.SDNet(opt, {k: v.clone() for k, v in embedding.items()})
    synth.fill_state_dict(net, seed=seed, bert_init=bert_init)
    net.eval()
    net.drop_emb = False
    return net


def run_reference(net, batch, capture=()):
    """Forward of the unmodified reference on a synth batch (CPU or CUDA).  Returns
    (probs, logits, captured) — logits are the pre-softmax row (hook on F.softmax input at
    Models/Layers.py:416-418), captured maps module names to their outputs."""
    import copy
    q, ocr, od = copy.deepcopy(batch)
    captured = {}
    hooks = []
    mods = dict(net.named_modules())
    for name in capture:
        def _mk(nm):
            def hook(_m, _inp, out):
                captured.setdefault(nm, []).append(out)
            return hook
        hooks.append(mods[name].register_forward_hook(_mk(name)))
    logits = {}
    F = torch.nn.functional
    orig_softmax = F.softmax

    def spy_softmax(x, dim=None, **kw):
        if x.dim() == 2 and dim == -1:
            logits["final"] = x.detach().clone()
        return orig_softmax(x, dim=dim, **kw)

    F.softmax = spy_softmax
    try:
        with torch.no_grad(), _cpu_cuda_identity():
            probs, _ = net(q, ocr, od)
    finally:
        F.softmax = orig_softmax
        for h in hooks:
            h.remove()
    return probs.detach(), logits.get("final"), captured



def run_reference_update(net, opt, batch, targets):
    """One UNMODIFIED `SDNetTrainer.update` (Models/SDNetTrainer.py:330-376) on `net`, driven through a
    stand-in trainer object that carries exactly the attributes `update` touches (the real constructor
    needs the ST-VQA data files).  Returns (loss, {name: grad})."""
    import types

    import torch.optim as optim
    import Models.SDNetTrainer as trainer_mod

    class _Meter(object):
        def update(self, *a, **k):
            pass

    fake = types.SimpleNamespace(network=net, opt=opt, train_loss=_Meter(), updates=0)
    fake.instance_bce_with_logits = types.MethodType(trainer_mod.SDNetTrainer.instance_bce_with_logits, fake)
    fake.loss_func = fake.instance_bce_with_logits
    params = [p for p in net.parameters() if p.requires_grad]
    fake.optimizer = optim.Adamax(params, lr=opt["lr"] if "lr" in opt else 2e-3)
    import copy
    q, ocr, od = copy.deepcopy(batch)
    with _cpu_cuda_identity():
        loss = trainer_mod.SDNetTrainer.update(fake, (q, ocr, od, targets, [{"q_id": i} for i in range(len(targets))]), 0)
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.requires_grad and p.grad is not None}
    return loss, grads


def stand_in_trainer(net, opt):
    """An object carrying exactly the attributes the UNMODIFIED `SDNetTrainer.predict` / `load_model` /
    `save_for_predict` / `update` methods touch (the real constructor needs the ST-VQA data files and
    spaCy).  The methods themselves are the reference's, bound to this object."""
    import types

    import Models.SDNetTrainer as trainer_mod

    class _Meter(object):
        val = avg = sum = count = 0

        def update(self, *a, **k):
            pass

    T = trainer_mod.SDNetTrainer
    fake = types.SimpleNamespace(network=net, opt=opt, train_loss=_Meter(), updates=0, fixed_answers_len=0,
                                 fixed_answers_entry=None)
    for name in ("instance_bce_with_logits", "predict", "update", "load_model", "save_for_predict", "save"):
        setattr(fake, name, types.MethodType(getattr(T, name), fake))
    fake.loss_func = fake.instance_bce_with_logits
    return fake


def predict_extras(batch):
    """gt_list / extra_info of a synth batch in the form `predict` reads them (SDNetTrainer.py:378-451):
    extra_info[i]['ocr_list'] = the image's OCR token strings incl. the '<OCR>' end token (its LENGTH
    is what the index rule uses, :409), answers None (no ANLS), q_id."""
    q, ocr, od = batch
    B = len(ocr["num_cnt"])
    M1 = ocr["position"].size(1) + 1
    gt = torch.zeros(B, M1)
    extra = []
    for i in range(B):
        n = int(ocr["num_cnt"][i])
        gt[i, i % max(1, n - 1)] = 1.0
        extra.append({"q_id": i, "answers": None, "ocr_list": ["tok%d" % k for k in range(n - 1)] + ["<OCR>"]})
    return gt, extra


def run_reference_predict(net, opt, batch, capture=()):
    """ONE forward of the unmodified reference driven through the unmodified `SDNetTrainer.predict`
    (Models/SDNetTrainer.py:378-451).  Returns (probs, logits, picks, captured): probs = what
    `self.network(...)` returned inside predict, logits = the pre-softmax row, picks = `save_res[i]['idx']`,
    i.e. the reference's own answer-index rule (:402-412) — not a restatement of it."""
    import copy
    q, ocr, od = copy.deepcopy(batch)
    gt, extra = predict_extras(batch)
    fake = stand_in_trainer(net, opt)
    captured = {}
    hooks = []
    mods = dict(net.named_modules())
    for name in capture:
        def _mk(nm):
            def hook(_m, _inp, out):
                captured.setdefault(nm, []).append(out)
            return hook
        hooks.append(mods[name].register_forward_hook(_mk(name)))
    seen = {}
    hooks.append(net.register_forward_hook(lambda _m, _i, out: seen.__setitem__("probs", out[0].detach().clone())))
    F = torch.nn.functional
    orig_softmax = F.softmax

    def spy_softmax(x, dim=None, **kw):
        if x.dim() == 2 and dim == -1:
            seen["logits"] = x.detach().clone()
        return orig_softmax(x, dim=dim, **kw)

    F.softmax = spy_softmax
    try:
        with torch.no_grad(), _cpu_cuda_identity():
            _loss, _anls, _acc, res, save_res = fake.predict((q, ocr, od, gt, extra))
    finally:
        F.softmax = orig_softmax
        for h in hooks:
            h.remove()
    picks = [int(r["idx"]) for r in save_res]
    answers = [r["answer"] for r in res]
    return seen["probs"], seen.get("logits"), picks, captured, answers
