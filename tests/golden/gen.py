"""Deterministic, name-seeded weights / inputs / tiny configurations shared by the golden-vector generator
(make_golden.py, build container only) and the tests that replay the fixtures (anywhere).

Weights are a pure function of (parameter name, shape, seed): no checkpoint has to travel with the repo.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Tuple

import torch

SKIP_SUBSTR = ("inv_freq", "position_ids")  # buffers that must keep their constructor value


def _seed(name: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:7], "little")


def make_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(_seed(name, seed))
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    is_norm_w = name.endswith("weight") and len(shape) == 1
    if name == "logit_scale" or len(shape) == 0:
        return torch.full(tuple(shape), math.log(1 / 0.07))
    if is_norm_w:
        return 1.0 + 0.1 * t
    if len(shape) == 1:
        return 0.1 * t
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    if "bias_k" in name or "bias_v" in name:
        return 0.5 * t
    if "embed_tokens" in name or "embedding" in name or "embed_positions" in name:
        return 0.5 * t
    return t / math.sqrt(fan_in)


def make_weights(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: make_tensor(k, tuple(v), seed) for k, v in shapes.items() if not any(s in k for s in SKIP_SUBSTR)}


# Tiny configurations.  "kernel" dims are chosen so the CUDA path supports them (LLaMA head_dim 128, encoder head_dim 64,
# video-long head_dim 96); "micro" is smaller still and only exercises the oracle against the reference.
TINY = dict(
    n_frames=2, attention_heads=2,
    clip=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=224,
              patch_size=14, projection_dim=192),
    whisper=dict(d_model=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=256, num_mel_bins=80,
                 max_source_positions=1500, decoder_layers=1, decoder_attention_heads=2, decoder_ffn_dim=64,
                 vocab_size=64),
    llama=dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
               rms_norm_eps=1e-6, max_position_embeddings=2048),
)


# A second shape family that only the ORACLE is pinned on (CPU): different widths / heads / frame count / down-sampler
# hyper-parameters (video-long head_dim 160 is outside the CUDA kernels' supported set, so the GPU tests skip it).
ALT = dict(
    n_frames=3, attention_heads=1,
    image_conv=(40, 24), video_conv=(50, 45), audio_conv=(300, 200),
    clip=dict(hidden_size=96, intermediate_size=160, num_hidden_layers=1, num_attention_heads=3, image_size=224,
              patch_size=14, projection_dim=160),
    whisper=dict(d_model=96, encoder_layers=1, encoder_attention_heads=3, encoder_ffn_dim=128, num_mel_bins=80,
                 max_source_positions=1500, decoder_layers=1, decoder_attention_heads=3, decoder_ffn_dim=48,
                 vocab_size=48),
    llama=dict(hidden_size=192, intermediate_size=320, num_hidden_layers=1, num_attention_heads=3, vocab_size=300,
               rms_norm_eps=1e-5, max_position_embeddings=2048),
)


def build_configs(spec: dict):
    """-> (CLIPConfig, WhisperConfig, LlamaConfig) from a spec like TINY (uses the transformers config classes only)."""
    from transformers import CLIPConfig, LlamaConfig, WhisperConfig

    c = spec["clip"]
    clip = CLIPConfig(
        text_config=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                         vocab_size=64, max_position_embeddings=16, projection_dim=c["projection_dim"]),
        vision_config=dict(hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"],
                           num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                           image_size=c["image_size"], patch_size=c["patch_size"],
                           projection_dim=c["projection_dim"], hidden_act="quick_gelu"),
        projection_dim=c["projection_dim"],
    )
    w = spec["whisper"]
    whisper = WhisperConfig(pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
                            suppress_tokens=None, begin_suppress_tokens=None, **w)
    llama = LlamaConfig(pad_token_id=0, bos_token_id=1, eos_token_id=2, **spec["llama"])
    for cfg in (clip, clip.vision_config, clip.text_config, whisper, llama):
        try:
            cfg._attn_implementation = "eager"
        except Exception:
            pass
    return clip, whisper, llama


def make_inputs(spec: dict, B: int, L: int, seed: int, modalities=("image", "audio", "video"), pad_tail: int = 0,
                with_labels: bool = True) -> dict:
    """Seeded synthetic inputs in the reference's `inputs` dict format (llm_trainer.py:366-381)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    V = spec["llama"]["vocab_size"]
    S = spec["clip"]["image_size"]
    d = dict(images=None, audios=None, videos=None)
    if "image" in modalities:
        d["images"] = torch.randn(B, 3, S, S, generator=g)
    if "audio" in modalities:
        d["audios"] = torch.randn(B, spec["whisper"]["num_mel_bins"], 2 * spec["whisper"]["max_source_positions"],
                                  generator=g)
    if "video" in modalities:
        d["videos"] = torch.randn(B, spec["n_frames"], 3, S, S, generator=g)
    ids = torch.randint(3, V - 6, (B, L), generator=g)
    ids[:, 0] = 1
    d["input_ids"] = ids
    mask = torch.ones(B, L, dtype=torch.int64)
    if pad_tail:
        mask[0, L - pad_tail:] = 0
    d["attention_mask"] = mask
    if with_labels:
        labels = ids.clone()
        labels[:, : L // 3] = -100
        labels[mask == 0] = -100
        d["labels"] = labels
    sp = [V - 6 + i for i in range(6)]  # six distinct special ids below V
    for i, name in enumerate(("image", "audio", "video")):
        d[f"{name}_starts"] = torch.full((B,), sp[2 * i], dtype=torch.int32)
        d[f"{name}_ends"] = torch.full((B,), sp[2 * i + 1], dtype=torch.int32)
    return d


# ------------------------------------------------------------------------------------------------ input-pipeline fixtures
PREPROCESS_IMAGE_SIZES = [(300, 400), (517, 389), (120, 200)]  # landscape down-scale, portrait down-scale, up-scale
PREPROCESS_AUDIO_SECONDS = [5.0, 31.0]                          # padded clip, trimmed clip


def synth_image(h: int, w: int, seed: int = 0):
    """Deterministic 8-bit RGB test image (smooth gradients + a checker + seeded noise), numpy only."""
    import numpy as np

    rng = np.random.default_rng(1000 + seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 17.0 + seed) * np.cos(y / 23.0),
                    255.0 * x / max(w - 1, 1),
                    255.0 * ((x.astype(np.int64) // 16 + y.astype(np.int64) // 16) % 2)], axis=-1)
    img += rng.normal(0.0, 12.0, size=img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_audio(seconds: float, seed: int = 0, sr: int = 16000):
    """Deterministic fp32 waveform in [-1, 1]: two chirps, an amplitude-modulated tone and low-level seeded noise."""
    import numpy as np

    rng = np.random.default_rng(2000 + seed)
    n = int(seconds * sr)
    t = np.arange(n, dtype=np.float64) / sr
    x = 0.4 * np.sin(2 * np.pi * (200.0 + 300.0 * t) * t) + 0.2 * np.sin(2 * np.pi * (4000.0 - 100.0 * t) * t)
    x += 0.15 * np.sin(2 * np.pi * 1000.0 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
    x += rng.normal(0.0, 0.003, size=n)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
