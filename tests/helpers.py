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


# ---------------------------------------------------------------------------------------------------- Philox (dropout masks)
def philox4x32_10(ctr: np.ndarray, key) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11) on an (n, 4) uint32 counter array with one 64-bit key (k0, k1): the
    generator of macaw-llm_b200/csrc/philox.cuh restated in numpy."""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return np.stack(c, axis=1).astype(np.uint32)


def dropout_multipliers(rows: int, cols: int, p: float, seed: int, sid: int) -> np.ndarray:
    """fp32 (rows, cols) multipliers (0 or 1/(1-p)) as csrc/philox.cuh defines them: element (r, c) = word (c & 3) of
    philox(key = (seed_lo, seed_hi), counter = (c >> 2, r, sid, 0)); kept iff word >= floor(p * 2^32)."""
    n4 = (cols + 3) // 4
    r, c4 = np.meshgrid(np.arange(rows, dtype=np.uint32), np.arange(n4, dtype=np.uint32), indexing="ij")
    ctr = np.stack([c4.ravel(), r.ravel(), np.full(r.size, sid, np.uint32), np.zeros(r.size, np.uint32)], axis=1)
    words = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(rows, n4 * 4)[:, :cols]
    thr = min(int(float(np.float32(p)) * 4294967296.0), 0xFFFFFFFF)
    return np.where(words >= np.uint32(thr), np.float32(1.0) / (np.float32(1.0) - np.float32(p)), np.float32(0.0)).astype(np.float32)
