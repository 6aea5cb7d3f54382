"""Imports the UNMODIFIED reference model (/root/reference/src/model.py, or its byte-identical copy
oracle/_ref/model.py made by oracle/build_ref.py where /root/reference does not exist: the GPU box)
under the installed transformers 5.5 / torch 2.11 — TEST / BENCH INFRASTRUCTURE ONLY.

Used by oracle/make_golden.py (fixture generation), by CPU tests that are skipped when no reference
copy is present, and by bench.py's reference arms (oracle/ref_runner.py), which time the reference
itself on the host cores and in torch eager on the B200.  The shims follow SURVEY.md §8(c): they only
restore names transformers 4.26.1 had and 5.5 removed, and keep model.py:401-408's
hard-coded .to("cuda") from failing on a CPU-only host.  No reference file is edited.
"""
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_src():
    if os.path.isfile("/root/reference/src/model.py"):
        return "/root/reference/src"
    from . import build_ref
    if build_ref.verify():
        return os.path.join(_HERE, "_ref")
    return None


REF_SRC = _ref_src()


def available():
    return REF_SRC is not None


_cached = {}


def load(no_caption_guard=False):
    """Returns the reference `model` module.  no_caption_guard=True execs the source with the
    one-line guard `caption_embeds = None` inserted before model.py:460 so that
    caption_ids=None (what main.py:147 passes) runs instead of raising UnboundLocalError."""
    key = bool(no_caption_guard)
    if key in _cached:
        return _cached[key]
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    if not hasattr(mu, "SequenceSummary"):
        mu.SequenceSummary = type("SequenceSummary", (torch.nn.Module,), {})
    for name in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):
        if not hasattr(pu, name):
            setattr(pu, name, None)
    if "transformers.utils.model_parallel_utils" not in sys.modules:
        m = types.ModuleType("transformers.utils.model_parallel_utils")
        m.assert_device_map = m.get_device_map = None
        sys.modules["transformers.utils.model_parallel_utils"] = m
    if not hasattr(mu.PreTrainedModel, "get_head_mask"):
        mu.PreTrainedModel.get_head_mask = lambda self, head_mask, n, is_attention_chunked=False: [None] * n
    if not torch.cuda.is_available() and not getattr(torch.nn.Module.to, "_ergm_cpu_redirect", False):
        _to = torch.nn.Module.to

        def to(self, *a, **k):
            a = ["cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in a]
            return _to(self, *a, **k)

        to._ergm_cpu_redirect = True
        torch.nn.Module.to = to
    path = os.path.join(REF_SRC, "model.py")
    src = open(path).read()
    modname = "ergm_reference_model_guarded" if no_caption_guard else "ergm_reference_model"
    if no_caption_guard:
        needle = "        if caption_ids is not None:\n            caption_ids = caption_ids.view"
        assert needle in src
        src = src.replace(needle, "        caption_embeds = None\n" + needle, 1)
    mod = types.ModuleType(modname)
    mod.__file__ = path
    sys.modules[modname] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    mod.GPT2LMHeadModel._tied_weights_keys = {"lm_head.weight": "transformer.wte.weight"}
    _cached[key] = mod
    return mod


def build_reference_model(cfg, state_dict, no_caption_guard=False, dropout=0.0):
    """Instantiates reference GPT2LMHeadModel for an OracleConfig and loads `state_dict`."""
    from transformers import GPT2Config
    mod = load(no_caption_guard)
    hf = GPT2Config(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.n_embd,
                    n_layer=cfg.n_layer, n_head=cfg.n_head, n_inner=None if cfg.n_inner == 4 * cfg.n_embd else cfg.n_inner,
                    layer_norm_epsilon=cfg.layer_norm_epsilon, initializer_range=cfg.initializer_range,
                    attn_pdrop=dropout, resid_pdrop=dropout, embd_pdrop=dropout)
    m = mod.GPT2LMHeadModel(hf)
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith(".attn.bias") or k.endswith(".masked_bias") or k == "lm_head.weight" for k in missing), missing
    m.lm_head.weight = m.transformer.wte.weight
    return m.eval() if dropout == 0.0 else m
