#!/usr/bin/env python
"""Entry point with the reference's model-construction surface (reference run_clm_llms.py:129-256, 462-497).

Only the part of the reference script that touches the forward hot path is kept: hyper-parameter flags →
`MM_LLMs_Config` → `MM_LLMs` → `llm.resize_token_embeddings` → encoder freezing → one forward.  Dataset loading, the
HF Trainer / DeepSpeed loop and checkpoint saving are outside the hot path.

  python run_clm_llms.py --check                 # tiny model, one synthetic image+audio+video+text forward on cuda:0
  python run_clm_llms.py --n_frames 6 --attention_heads 8 --llm_model_name_or_path DIR ...   # real configs from disk
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    # ModelArguments of the reference (run_clm_llms.py:194-251).  NB: the reference's CLI default for
    # --attention_heads is 220, which cannot construct nn.MultiheadAttention(4096, 440); the config-class default (8) is used.
    ap.add_argument("--n_frames", type=int, default=6)
    ap.add_argument("--attention_heads", type=int, default=8)
    ap.add_argument("--image_conv_kernel", type=int, default=48)
    ap.add_argument("--image_conv_stride", type=int, default=36)
    ap.add_argument("--video_conv_kernel", type=int, default=36)
    ap.add_argument("--video_conv_stride", type=int, default=30)
    ap.add_argument("--audio_conv_kernel", type=int, default=240)
    ap.add_argument("--audio_conv_stride", type=int, default=220)
    ap.add_argument("--freeze_multi_modal_encoder", type=lambda s: s.lower() == "true", default=True)
    ap.add_argument("--clip_model_name_or_path", default="trained_models/clip_model")
    ap.add_argument("--whisper_model_name_or_path", default="trained_models/whisper_model")
    ap.add_argument("--llm_model_name_or_path", default="trained_models/llama_model")
    ap.add_argument("--vocab_extra_tokens", type=int, default=7, help="[PAD] + 6 modal tokens (llm_trainer.py:126-133)")
    ap.add_argument("--check", action="store_true", help="tiny model + one synthetic forward (needs a CUDA device)")
    return ap.parse_args(argv)


def prepare_model_for_training(model):
    """run_clm_llms.py:390-393: every parameter whose name contains 'encoder' is frozen (the flag is ignored there too)."""
    for name, p in model.named_parameters():
        if "encoder" in name:
            p.requires_grad = False
    return model


def build_model(args):
    from transformers import AutoConfig, CLIPConfig, WhisperConfig

    from modeling import MM_LLMs, MM_LLMs_Config

    if args.check:
        from tests.golden import gen

        clip_config, whisper_config, llm_config = gen.build_configs(gen.TINY)
        args.n_frames, args.attention_heads = gen.TINY["n_frames"], gen.TINY["attention_heads"]
    else:
        clip_config = CLIPConfig.from_pretrained(args.clip_model_name_or_path)
        whisper_config = WhisperConfig.from_pretrained(args.whisper_model_name_or_path)
        llm_config = AutoConfig.from_pretrained(args.llm_model_name_or_path)
    model_config = MM_LLMs_Config(
        n_frames=args.n_frames, attention_heads=args.attention_heads, image_conv_kernel=args.image_conv_kernel,
        image_conv_stride=args.image_conv_stride, video_conv_kernel=args.video_conv_kernel,
        video_conv_stride=args.video_conv_stride, audio_conv_kernel=args.audio_conv_kernel,
        audio_conv_stride=args.audio_conv_stride, clip_config=clip_config, whisper_config=whisper_config,
        llm_config=llm_config)
    model = MM_LLMs(config=model_config)
    model.llm.resize_token_embeddings(llm_config.vocab_size + args.vocab_extra_tokens)
    return prepare_model_for_training(model)


def main(argv=None):
    import torch

    args = parse_args(argv)
    model = build_model(args)
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_all = sum(p.numel() for p in model.parameters())
    print(f"MM_LLMs: {n_all / 1e6:.1f} M parameters, {n_train / 1e6:.1f} M trainable (encoders frozen)")
    if not args.check:
        return
    if not torch.cuda.is_available():
        raise SystemExit("--check needs a CUDA device: the forward has no CPU path")
    from tests.golden import gen

    model = model.cuda().to(torch.bfloat16).eval()
    V = model.llm.model.embed_tokens.weight.shape[0]
    inp = gen.make_inputs(gen.TINY, B=2, L=32, seed=0)
    inp["image_starts"][:], inp["image_ends"][:] = V - 6, V - 5      # the resized table carries the modal tokens
    out = model(inp)
    torch.cuda.synchronize()
    print(f"forward ok: loss {float(out.loss):.4f}, logits {tuple(out.logits.shape)}")


if __name__ == "__main__":
    main()
