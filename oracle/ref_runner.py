"""Runs the UNMODIFIED reference `modeling.py` (staged under oracle/_ref/ by oracle/make_ref.py) on the host CPUs.

TEST INFRASTRUCTURE, NOT PRODUCT CODE: imported only by tests/, __graft_entry__ and bench.py's `cpu_baseline` /
`--impl reference` legs.  Nothing under macaw-llm_b200/ may import this module.

Shims, applied to the importing process, never to the file:
  1. reference modeling.py:25 imports PretrainedConfig from transformers.modeling_utils (no longer re-exported);
  2. reference modeling.py:939 calls init_weights() without post_init() (transformers 5.x needs all_tied_weights_keys).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(HERE, "_ref", "modeling.py")
_MOD = None


def available() -> bool:
    return os.path.exists(REF_FILE)


def load_reference():
    """Import oracle/_ref/modeling.py under a private module name (the repo's own drop-in `modeling` is not shadowed)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    if not available():
        raise RuntimeError("oracle/_ref/modeling.py is absent: run `python oracle/make_ref.py` where /root/reference exists")
    meta = os.path.join(HERE, "_ref", "SOURCE.json")
    if os.path.exists(meta):
        from oracle.make_ref import sha256

        want = json.load(open(meta))["sha256"]
        if sha256(REF_FILE) != want:
            raise RuntimeError("oracle/_ref/modeling.py differs from the staged reference (sha256 mismatch)")
    import transformers.modeling_utils as mu
    from transformers import PretrainedConfig, PreTrainedModel

    mu.PretrainedConfig = PretrainedConfig  # shim 1
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("_macaw_reference_modeling", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_macaw_reference_modeling"] = mod
    spec.loader.exec_module(mod)
    _orig = PreTrainedModel.init_weights

    def _iw(self):  # shim 2
        return self.post_init() if not hasattr(self, "all_tied_weights_keys") else _orig(self)

    mod.MM_LLMs.init_weights = _iw
    _MOD = mod
    return mod


def random_state_dict(model_meta, seed: int = 0, std: float = 0.02, dtype=torch.float32):
    """Random weights of the family the reference's init_weights() draws from (N(0, std) matrices, unit norm gains, zero
    biases), as a state_dict for `build_model` — generated directly so the 8 B-parameter model is initialised ONCE."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model_meta.state_dict().items():
        if k.endswith(("inv_freq", "position_ids")):
            continue
        shp = tuple(v.shape)
        if len(shp) == 0:
            sd[k] = torch.tensor(2.6592, dtype=dtype)
        elif len(shp) == 1 and "norm" in k.lower() and k.endswith("weight"):
            sd[k] = torch.ones(shp, dtype=dtype)
        elif len(shp) == 1:
            sd[k] = torch.zeros(shp, dtype=dtype)
        else:
            sd[k] = torch.empty(shp, dtype=dtype).normal_(0.0, std, generator=g)
    return sd


def build_model(clip_cfg, whisper_cfg, llama_cfg, hyper: dict, state_dict=None, dtype=torch.float32, seed: int = 0):
    """Reference `MM_LLMs(config)` in eval mode on the CPU, every parameter taken from `state_dict` (or, when None, from
    `random_state_dict`).  The module tree is built by the reference's own constructor on the META device (no 32 GB
    allocation + random init that would be overwritten anyway), parameters are attached with `load_state_dict(assign=True)`,
    and the non-persistent buffers the constructor computes (CLIP position_ids, LLaMA rotary cos/sin caches, reference
    modeling.py:93-105) are rebuilt by calling the reference's own LlamaRotaryEmbedding constructor on the CPU."""
    mod = load_reference()
    for c in (clip_cfg, whisper_cfg):  # eager attention inside the HF encoders: deterministic
        try:
            c._attn_implementation = "eager"
            for sub in ("vision_config", "text_config"):
                if hasattr(c, sub):
                    getattr(c, sub)._attn_implementation = "eager"
        except Exception:
            pass
    cfg = mod.MM_LLMs_Config(clip_config=clip_cfg, whisper_config=whisper_cfg, llm_config=llama_cfg, **hyper)
    with torch.device("meta"):
        model = mod.MM_LLMs(cfg)
    if state_dict is None:
        state_dict = random_state_dict(model, seed=seed, dtype=dtype)
    own = model.state_dict()
    sd = {}
    for k, v in state_dict.items():
        if k in own and not k.endswith(("inv_freq", "position_ids")):
            sd[k] = v.detach().to("cpu", dtype) if v.is_floating_point() else v.detach().cpu()
    missing, unexpected = model.load_state_dict(sd, strict=False, assign=True)
    bad = [m for m in missing if not m.endswith(("inv_freq", "position_ids"))]
    assert not bad and not unexpected, (bad[:5], unexpected[:5])
    # buffers computed by constructors
    hd = llama_cfg.hidden_size // llama_cfg.num_attention_heads
    for m in model.modules():
        if isinstance(m, mod.LlamaRotaryEmbedding):
            fresh = mod.LlamaRotaryEmbedding(hd, max_position_embeddings=llama_cfg.max_position_embeddings)
            for n in ("inv_freq", "cos_cached", "sin_cached"):
                m._buffers[n] = fresh._buffers[n]
    for n, b in list(model.named_buffers()):
        if b.is_meta:
            if n.endswith("position_ids"):
                owner = model.get_submodule(n.rsplit(".", 1)[0])
                owner._buffers["position_ids"] = torch.arange(b.shape[-1]).expand(b.shape).clone()
            else:
                raise RuntimeError(f"reference buffer {n} was not materialised")
    leftover = [n for n, p in model.named_parameters() if p.is_meta]
    assert not leftover, leftover[:5]
    return model.eval()
