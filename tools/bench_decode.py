#!/usr/bin/env python
"""Decode-path measurement (row 8f-2): image+text prefill then greedy decoding at LLaMA-7B width.
Prints prefill ms and ms per decode step / tokens per second.  Not the benchmark of record (bench.py is)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq-len", type=int, default=256)
    ap.add_argument("--new", type=int, default=32)
    a = ap.parse_args()
    from macaw_llm_b200.modeling import MM_LLMs, MM_LLMs_Config

    (clip, whisper, llama), hyper = bench.real_configs()
    cfg = MM_LLMs_Config(clip_config=clip, whisper_config=whisper, llm_config=llama, **hyper)
    model = MM_LLMs.build_random(cfg, device="cuda", dtype=torch.bfloat16, seed=0)
    host = bench.synth_inputs(a.batch, a.seq_len, llama.vocab_size, 224, 3000, 1234)
    host["audios"] = None
    dev_in = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    eng = model.engine
    for n in (2, a.new, 1, a.new):  # warm-up, then: prefill + (new-1) steps, prefill only, again
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        toks = eng.generate(dev_in, max_new_tokens=n, eos_token_id=-1)  # eos -1: never stop early (random weights)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"[bench_decode] B={a.batch} new={n}: {dt * 1e3:.1f} ms, out {tuple(toks.shape)}")
        if n == 1:
            t_prefill = dt
        last = dt
    per_step = (last - t_prefill) / (a.new - 1)
    print(f"[bench_decode] prefill {t_prefill * 1e3:.1f} ms (T={a.seq_len + 8}); decode {per_step * 1e3:.2f} ms/step -> "
          f"{a.batch / per_step:.0f} tokens/s at B={a.batch} (weight streaming floor 13.5 GB / 6.57 TB/s = 2.05 ms)")


if __name__ == "__main__":
    main()
