"""Shared test helpers: fixture loading, model construction from name-seeded weights, error metrics."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from tests.golden import gen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load_shapes():
    with open(os.path.join(GOLDEN, "tiny_shapes.json")) as f:
        d = json.load(f)
    return d["spec"], d["hp"], {k: tuple(v) for k, v in d["shapes"].items()}


def load_case(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN, f"tiny_{name}.npz"))
    out = {k: z[k] for k in z.files}
    out["modalities"] = tuple(str(m) for m in out["modalities"])
    return out


def case_inputs(spec: dict, case: dict) -> dict:
    return gen.make_inputs(spec, int(case["B"]), int(case["L"]), seed=int(case["seed"]),
                           modalities=case["modalities"], pad_tail=int(case["pad_tail"]),
                           with_labels=bool(int(case["with_labels"])))


def bf16_round(sd: dict) -> dict:
    return {k: (v.to(torch.bfloat16).float() if v.is_floating_point() else v) for k, v in sd.items()}


def build_tiny_model(device="cuda", dtype=torch.bfloat16):
    """MM_LLMs (this repo) with the tiny golden config and name-seeded weights, on `device` in `dtype`."""
    from macaw_llm_b200.modeling import MM_LLMs, MM_LLMs_Config

    spec, hp, shapes = load_shapes()
    clip, whisper, llama = gen.build_configs(spec)
    cfg = MM_LLMs_Config(n_frames=spec["n_frames"], attention_heads=spec["attention_heads"], clip_config=clip,
                         whisper_config=whisper, llm_config=llama)
    model = MM_LLMs(cfg)
    weights = gen.make_weights(shapes, seed=0)
    missing, unexpected = model.load_state_dict(weights, strict=False)
    assert not unexpected and all(any(s in m for s in gen.SKIP_SUBSTR) for m in missing), (missing, unexpected)
    model = model.to(device=device, dtype=dtype).eval()
    return model, spec, hp, weights
