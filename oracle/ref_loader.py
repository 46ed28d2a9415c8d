"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference implementation.

Puts /root/reference on sys.path (with two tiny shims for modules that are not
installed here) and exposes the reference's own functions for the hot path so
that `oracle/make_golden.py` can run them on seeded inputs and commit the
results under tests/golden/.  /root/reference does not exist on the GPU box:
nothing in `-m gpu` tests, smoke() or bench.py may import this module.

Accommodations, none of which edits a reference file:
  * `easydict`, `ipdb` are shimmed (oracle/shims/).
  * no GPU here -> gloo process group, `Tensor.cuda` patched to identity,
    `torch.cuda.ByteTensor = torch.ByteTensor` (utils/distributed.py:75-77).
  * VAST is never constructed (needs pretrained weights, and bert.py does not
    import under transformers 5.x): the *unbound* methods `VAST.batch_get`
    (model/vast.py:82) and `VAST.forward_ret` (model/vast.py:383) run on a stub
    `self` whose batch dict is pre-seeded with the pooled features.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VAST_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


_loaded = None


def load(init_pg: bool = True):
    """Import the reference. Returns a namespace with VAST, E (evaluation_mm),
    D (utils.distributed), G (general_module)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    import torch
    import torch.distributed as dist

    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.ByteTensor = torch.ByteTensor
    if init_pg and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)

    # The reference's model/__init__ pulls every encoder; bert.py fails under
    # transformers 5.x.  Import model.vast with a stub for the bert module only.
    try:
        from model.vast import VAST  # noqa
    except Exception:
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        bert_stub = types.ModuleType("model.text_encoders.bert.bert")
        bert_stub.BertForMaskedLM = object
        bert_stub.BertConfig = object
        sys.modules["model.text_encoders.bert.bert"] = bert_stub
        from model.vast import VAST  # noqa
    import evaluation.evaluation_mm as E
    import utils.distributed as D
    import model.general_module as G

    _loaded = types.SimpleNamespace(VAST=VAST, E=E, D=D, G=G)
    return _loaded


def make_stub_model(contra_temp=0.07, itm_ratio=0.1, hidden=16, seed=0,
                    vision_encoder_type="evaclip01_giant", audio_encoder_type="beats"):
    """A stand-in for `self` that drives the reference's unbound methods.

    `multimodal_encoder.bert` is a tiny deterministic cross-encoder so that the
    ITM branch of forward_ret (model/vast.py:449-457) and compute_slice_scores
    (model/vast.py:373-380) execute; its output only depends on
    (input_ids, attention_mask, encoder_hidden_states) so gathers can be checked.
    """
    import torch
    import torch.nn as nn
    ns = load()
    VAST, G = ns.VAST, ns.G

    class TinyCross(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            self.emb = nn.Parameter(torch.randn(30522, hidden, generator=g) * 0.1)
            self.proj = nn.Parameter(torch.randn(hidden, hidden, generator=g) * 0.3)

        def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None):
            x = self.emb[input_ids] * attention_mask.unsqueeze(-1).to(self.emb.dtype)
            ctx = encoder_hidden_states.float().mean(dim=1, keepdim=True)[..., :hidden]
            h = torch.tanh((x + ctx) @ self.proj)
            return types.SimpleNamespace(last_hidden_state=h)

    class ItmHead(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(seed + 1)
            self.w = nn.Parameter(torch.randn(hidden, 2, generator=g))

        def forward(self, x):  # accepts the .half() of vast.py:453
            return x.float() @ self.w

    class Stub(nn.Module):
        batch_get = VAST.batch_get
        forward_ret = VAST.forward_ret
        compute_slice_scores = VAST.compute_slice_scores
        pool_vision_for_contra = G.MMGeneralModule.pool_vision_for_contra
        pool_text_for_contra = G.MMGeneralModule.pool_text_for_contra
        pool_audio_for_contra = G.MMGeneralModule.pool_audio_for_contra

        def __init__(self):
            super().__init__()
            self.contra_temp = nn.Parameter(torch.tensor(float(contra_temp)))
            self.itm_ratio = itm_ratio
            self.multimodal_encoder = types.SimpleNamespace(bert=TinyCross())
            self.itm_head = ItmHead()
            self.config = types.SimpleNamespace(
                vision_encoder_type=vision_encoder_type,
                audio_encoder_type=audio_encoder_type,
                itm_rerank_num=50, ret_bidirection_evaluation=False)

    return Stub()
