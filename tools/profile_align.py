#!/usr/bin/env python
"""Run only the alignment cross-attention block at a BASELINE shape (for ncu): B x N x C modal features against the
V x E embedding table, 16 heads.  Default = cfg4 image modality on one GPU: B=32, 256 x 768 -> Lq=6 -> Nq=192, R=3072."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tokens", type=int, default=256)
    ap.add_argument("--channels", type=int, default=768)
    ap.add_argument("--kernel", type=int, default=48)
    ap.add_argument("--stride", type=int, default=36)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    from macaw_llm_b200 import ops
    from macaw_llm_b200.engine import Engine

    E, V, H = 4096, 32000, 16

    class M(torch.nn.Module):
        pass

    m = M()
    m.project_image = torch.nn.Conv1d(a.channels, a.channels, a.kernel, a.stride)
    m.transform_image_to_hidden = torch.nn.Linear(a.channels, E)
    m.image_align_attention = torch.nn.MultiheadAttention(E, H, dropout=0.1, add_bias_kv=True, add_zero_attn=True)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0, 0.02)
    m = m.cuda().to(torch.bfloat16).eval()
    table = (torch.randn(V, E) * 0.02).cuda().to(torch.bfloat16)
    feats = torch.randn(a.batch, a.tokens, a.channels).cuda().to(torch.bfloat16)
    eng = Engine(m)
    Lq = (a.tokens - a.kernel) // a.stride + 1
    prefix = torch.zeros(a.batch, Lq + 2, E, device="cuda", dtype=torch.bfloat16)
    eng.align(feats, "image", table, prefix, 1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = []
    e0.record()
    for _ in range(a.iters):
        eng.align(feats, "image", table, prefix, 1)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    Nq = a.batch * Lq
    flops = 4.0 * H * Nq * V * E
    by = {}
    for tag, fl, s, t in ops.PROFILE:
        d = by.setdefault(tag, [0.0, 0.0])
        d[0] += fl
        d[1] += s.elapsed_time(t) * 1e-3
    print(f"[profile_align] Nq={Nq} R={H * Nq}: {e0.elapsed_time(e1) / a.iters:.3f} ms per block; table contractions "
          f"{flops / 1e12:.2f} TF;", {k: f"{v[0] / v[1] / 1e12:.0f} TF/s" for k, v in by.items()})


if __name__ == "__main__":
    main()
