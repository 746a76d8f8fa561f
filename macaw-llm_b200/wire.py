"""Wire formats either side of the hot path: the tokenizer's special ids, the pickle dataset
schema and the `inputs` dict the reference trainer hands to `MM_LLMs.forward`.

  special ids            /root/reference/llm_trainer.py:126-133 ('<image>' .. '</video>' = 32000 .. 32005),
                         run_clm_llms.py:353 / modeling.py:958 ([PAD] = 32006; the table is resized to 32007 rows, :495)
  dataset cache schema   /root/reference/preprocess_data_supervised.py:438-451 — a pickle of a dict of equal-length
                         lists: input_ids, attention_mask (max_length 256, padded), labels, images, audios, videos
                         (each an index into the media name list, or -1 for "absent")
  label masking          /root/reference/run_clm_llms.py:353-356: labels equal to the pad id become IGNORE_INDEX
  inputs dict            /root/reference/llm_trainer.py:363-381 (`get_self_inputs`): float media tensors (zeros when
                         absent, :315, :332, :352), int64 ids / mask / labels, int32 (B,) start / end token ids

Pure host-side plumbing over torch tensors; no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import pickle
from typing import Dict, List, Optional, Sequence

import torch

SPECIAL_TOKENS = {"<image>": 32000, "</image>": 32001, "<audio>": 32002, "</audio>": 32003, "<video>": 32004,
                  "</video>": 32005}
PAD_TOKEN_ID = 32006
VOCAB_WITH_SPECIALS = 32007  # len(tokenizer) after the seven additions: model.llm.resize_token_embeddings(32007)
IGNORE_INDEX = -100
CACHE_KEYS = ("input_ids", "attention_mask", "labels", "images", "audios", "videos")


def load_cache(path: str) -> Dict[str, list]:
    """Read a `data/*.cache` pickle written by the reference's preprocess scripts and validate its schema."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    validate_cache(d)
    return d


def validate_cache(d: Dict[str, list]) -> int:
    missing = [k for k in CACHE_KEYS if k not in d]
    if missing:
        raise KeyError(f"dataset cache lacks {missing}; expected keys {CACHE_KEYS}")
    n = len(d["input_ids"])
    for k in CACHE_KEYS:
        if len(d[k]) != n:
            raise ValueError(f"dataset cache: column {k} has {len(d[k])} rows, input_ids has {n}")
    return n


def mask_pad_labels(labels: Sequence[Sequence[int]], pad_token_id: int = PAD_TOKEN_ID) -> List[List[int]]:
    """run_clm_llms.py:353-356."""
    return [[(l if l != pad_token_id else IGNORE_INDEX) for l in row] for row in labels]


def collate(cache: Dict[str, list], indices: Sequence[int]) -> Dict[str, torch.Tensor]:
    """Rows `indices` of a dataset cache as the batch the HF Trainer would deliver (int64 tensors; media columns keep the
    (B, 1) index layout `get_self_inputs` expects: `vid = vid[0]`, llm_trainer.py:311)."""
    out = {}
    for k in ("input_ids", "attention_mask"):
        out[k] = torch.tensor([cache[k][i] for i in indices], dtype=torch.int64)
    out["labels"] = torch.tensor(mask_pad_labels([cache["labels"][i] for i in indices]), dtype=torch.int64)
    for k in ("images", "audios", "videos"):
        out[k] = torch.tensor([[int(cache[k][i])] for i in indices], dtype=torch.int64)
    return out


def make_inputs(batch: Dict[str, torch.Tensor], images: Optional[torch.Tensor], audios: Optional[torch.Tensor],
                videos: Optional[torch.Tensor], *, n_frames: int = 6, image_size: int = 224, mel_frames: int = 3000,
                dtype=torch.bfloat16, inference: bool = False) -> Dict[str, object]:
    """The `inputs` dict of llm_trainer.py:363-381.  A modality passed as None is materialised as ZEROS, as the reference
    trainer does for index -1 (the model itself also accepts None per modality, modeling.py:967-969 — pass the tensors
    you want spliced)."""
    B = batch["input_ids"].shape[0]
    if images is None:
        images = torch.zeros(B, 3, image_size, image_size)
    if audios is None:
        audios = torch.zeros(B, 80, mel_frames)
    if videos is None:
        videos = torch.zeros(B, n_frames, 3, image_size, image_size)
    d = {
        "videos": videos.to(dtype), "audios": audios.to(dtype), "images": images.to(dtype),
        "input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"], "labels": batch.get("labels"),
    }
    for name in ("image", "audio", "video"):
        d[f"{name}_starts"] = torch.full((B,), SPECIAL_TOKENS[f"<{name}>"], dtype=torch.int32)
        d[f"{name}_ends"] = torch.full((B,), SPECIAL_TOKENS[f"</{name}>"], dtype=torch.int32)
    if inference:
        d["inference"] = True
    return d
