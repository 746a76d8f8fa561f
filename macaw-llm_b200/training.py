"""Training step of the MM_LLMs hot path on the sm_100a kernel library.

Reference: the HF Trainer drives `compute_loss -> model(**inputs)[0]` then autograd / optimizer
(/root/reference/llm_trainer.py:184-188, train.sh:14-41); every parameter whose name contains 'encoder' is frozen
(run_clm_llms.py:390-393); the loss is the shifted cross entropy of modeling.py:597-610.

What runs here (all on hand-written kernels; torch only owns memory and the autograd hook):
  forward   the vendored LLaMA decoder (modeling.py:234-299) with activations kept for the backward pass: RMSNorm kernel,
            q / k / v / o / gate / up / down GEMMs on the raw nn.Linear weights (RoPE in the q / k epilogues, residual adds
            in the o / down epilogues), tcgen05 flash attention, SwiGLU kernel, lm_head GEMM, CE kernel.
  backward  dX = dY W (MN-major B operand), dW = dY^T X (MN-major A and B operands — the activations are read as the
            forward pass stored them), attention backward composed of batched tcgen05 GEMMs around one row-wise
            softmax-backward kernel, RMSNorm / SwiGLU / RoPE / CE backward kernels, embedding-gradient scatter.
  update    fused AdamW with fp32 master weights and moments (`FusedAdamW`).
  sync      gradients live in ONE flat bf16 buffer (`GradBuffer`); `allreduce_grads` averages it across data-parallel
            ranks in contiguous buckets (NCCL on GPUs, gloo in the CPU test tier), launched per decoder layer while the
            backward pass is still running.

Differentiable set: every LLaMA parameter (decoder layers, norms, lm_head, embed_tokens) AND the alignment modules of
every modality (project_* Conv1d, transform_*_to_hidden Linear, *_align_attention in/out projections and bias_k / bias_v)
— their backward runs through the (V + 2)-key softmax of the absorbed alignment attention, including the embedding
table's gradient as the attention's keys and values (`AlignTrainer`) — AND `video_long_self_attention` (through the
Conv1d data gradient of the video down-sampler).  That is every parameter the reference trains: the encoders are frozen
as there (names containing 'encoder').  One documented omission: the attention-probability dropout of the MHAs
(p = 0.1, modeling.py:879) — the training step runs them without dropout (SURVEY.md §7 allows "implements Philox
dropout or documents the omission").  `trainable_parameters` lists what gets a gradient.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import ops

BF16 = torch.bfloat16


ALIGN_MODALITIES = ("image", "audio", "video")


def _video_long_params(model) -> List[torch.Tensor]:
    mha = model.video_long_self_attention
    return [mha.in_proj_weight, mha.in_proj_bias, mha.bias_k, mha.bias_v, mha.out_proj.weight, mha.out_proj.bias]


def _align_params(model, name: str) -> List[torch.Tensor]:
    conv = getattr(model, f"project_{name}")
    lin = getattr(model, f"transform_{name}_to_hidden")
    mha = getattr(model, f"{name}_align_attention")
    return [conv.weight, conv.bias, lin.weight, lin.bias, mha.in_proj_weight, mha.in_proj_bias, mha.bias_k, mha.bias_v,
            mha.out_proj.weight, mha.out_proj.bias]


def trainable_parameters(model) -> List[tuple]:
    """(name, parameter) pairs this training step produces gradients for."""
    pre = tuple(f"project_{n}." for n in ALIGN_MODALITIES) + tuple(f"transform_{n}_to_hidden." for n in ALIGN_MODALITIES) + \
        tuple(f"{n}_align_attention." for n in ALIGN_MODALITIES) + ("video_long_self_attention.",)
    return [(n, p) for n, p in model.named_parameters() if n.startswith("llm.") or n.startswith(pre)]


def freeze_like_reference(model) -> None:
    """run_clm_llms.py:390-393: every parameter whose name contains 'encoder' is frozen (the flag is ignored there too)."""
    for n, p in model.named_parameters():
        if "encoder" in n:
            p.requires_grad_(False)


# ---------------------------------------------------------------------------------------------------- gradient storage
class GradBuffer:
    """One flat bf16 buffer holding the gradient of every trainable parameter, in BACKWARD order (lm_head, final norm,
    layers L-1 .. 0, embed_tokens) so that the bucket of a finished layer is a contiguous slice that can be all-reduced
    while earlier layers are still being differentiated.  `p.grad` is a view into the buffer."""

    def __init__(self, model):
        llm = model.llm
        order = [llm.lm_head.weight, llm.model.norm.weight]
        self.layer_slices = []
        groups = [order]
        for layer in reversed(list(llm.model.layers)):
            ps = [layer.mlp.down_proj.weight, layer.mlp.gate_proj.weight, layer.mlp.up_proj.weight,
                  layer.post_attention_layernorm.weight, layer.self_attn.o_proj.weight, layer.self_attn.q_proj.weight,
                  layer.self_attn.k_proj.weight, layer.self_attn.v_proj.weight, layer.input_layernorm.weight]
            groups.append(ps)
        # last bucket: the embedding table (gathered rows + keys / values of the alignment attention) and the alignment modules
        self.align_params = {n: _align_params(model, n) for n in ALIGN_MODALITIES}
        self.align_params["video"] = self.align_params["video"] + _video_long_params(model)
        groups.append([llm.model.embed_tokens.weight] + [p for n in ALIGN_MODALITIES for p in self.align_params[n]])
        dev = llm.lm_head.weight.device
        total = sum(p.numel() for g in groups for p in g)
        self.flat = torch.zeros((total,), device=dev, dtype=BF16)
        self.views: Dict[int, torch.Tensor] = {}
        self.buckets = []  # (start, end) per group, in backward order
        off = 0
        for g in groups:
            start = off
            for p in g:
                n = p.numel()
                # 16-byte aligned views (GEMM epilogues store vectors): every parameter size here is a multiple of 8
                self.views[id(p)] = self.flat[off:off + n].view(p.shape)
                off += n
            self.buckets.append((start, off))
        self.params = [p for g in groups for p in g]

    def attach(self, skip=()) -> Dict[int, bool]:
        """Point every `p.grad` at its view.  Returns {id(p): fresh}: a parameter whose grad was None (after
        `zero_grad(set_to_none=True)`) is overwritten by the next backward pass, otherwise accumulated into.
        Parameters in `skip` (alignment modules of modalities absent from the batch) keep `grad = None`."""
        fresh = {}
        skip = {id(p) for p in skip}
        for p in self.params:
            if id(p) in skip:
                fresh[id(p)] = True
                continue
            v = self.views[id(p)]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
                fresh[id(p)] = True
            else:
                fresh[id(p)] = False
        return fresh

    def zero(self) -> None:
        for p in self.params:
            p.grad = None


def allreduce_grads(buf: GradBuffer, world: int, bucket: Optional[int] = None, async_op: bool = False):
    """Average the flat gradient buffer (or one bucket of it) over the data-parallel group: one collective per contiguous
    bucket, bf16 on the wire (north_star: "a single NCCL all-reduce on gradients")."""
    import torch.distributed as dist

    if world <= 1:
        return None
    if bucket is None:
        t = buf.flat
    else:
        s, e = buf.buckets[bucket]
        t = buf.flat[s:e]
    if t.is_cuda:
        from . import dist as D

        if D._NCCL_READY:  # the kernel library's own communicator (mm_nccl_allreduce), on the caller's current stream
            D.nccl_allreduce_(t, average=True)
            return None
        # torch.distributed's NCCL group: native average, bf16 on the wire, asynchronous w.r.t. the compute stream
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
    tmp = t.float()  # gloo (CPU test tier): sum in fp32, then average
    dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
    t.copy_(tmp / world)
    return None


# ---------------------------------------------------------------------------------------------------- optimizer
class FusedAdamW:
    """AdamW (decoupled weight decay) with fp32 master weights and moments; one fused kernel launch per parameter tensor.
    Mirrors torch.optim.AdamW's update rule (checked against it in tests/test_train_gpu.py)."""

    def __init__(self, params, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.state: Dict[int, tuple] = {}
        self.t = 0
        self.t_dev: Optional[torch.Tensor] = None  # device-side step counter (the kernels read it: graph-replayable)

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self, grad_scale: float = 1.0) -> None:
        self.t += 1
        with torch.no_grad():
            if self.t_dev is None:
                dev = next(p for p in self.params).device
                self.t_dev = torch.zeros((1,), device=dev, dtype=torch.int32)
            self.t_dev.add_(1)
            for p in self.params:
                if p.grad is None:
                    continue
                st = self.state.get(id(p))
                if st is None:
                    st = (p.detach().float().clone().contiguous(), torch.zeros_like(p, dtype=torch.float32),
                          torch.zeros_like(p, dtype=torch.float32))
                    self.state[id(p)] = st
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.adamw(p.data, g, st[0], st[1], st[2], lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                          eps=self.eps, weight_decay=self.weight_decay, step=self.t, grad_scale=grad_scale,
                          step_dev=self.t_dev)
                # the kernel's in-place write is invisible to autograd's version counter, which the engine's
                # derived-weight / CUDA-graph caches key on: bump it
                torch.autograd.graph.increment_version(p)


# ---------------------------------------------------------------------------------------------------- LLaMA forward / backward
class LlamaTrainer:
    """Forward-with-activations and backward of the decoder stack of one MM_LLMs model."""

    def __init__(self, model):
        self.m = model
        self.grads: Optional[GradBuffer] = None
        self.world = 1
        self.overlap_allreduce = True
        self._pending = []

    # ---- helpers
    def _w(self, p: torch.Tensor) -> torch.Tensor:
        if not p.is_cuda or p.dtype != BF16:
            raise RuntimeError("macaw_b200 training: parameters must be bf16 CUDA tensors (model.to('cuda', torch.bfloat16)); "
                               "there is no CPU / fp32 execution path")
        return p.detach()

    def _dims(self):
        cfg = self.m.llm.config
        E, H = cfg.hidden_size, cfg.num_attention_heads
        hd = E // H
        if hd != 128:
            raise NotImplementedError(f"macaw_b200 training: LLaMA head_dim {hd} unsupported (RoPE epilogue is specialised for 128)")
        return E, H, hd, cfg.intermediate_size, cfg.rms_norm_eps

    def forward(self, embeds: torch.Tensor, attention_mask: Optional[torch.Tensor], labels: torch.Tensor):
        """embeds (B, T, E) bf16 (the spliced inputs_embeds), mask (B, T) | None, labels (B, T) int64 -> (loss, ctx)."""
        ops.TAG = "train.fwd"
        if ops.ACT() != BF16:
            raise RuntimeError("macaw_b200 training: the training step computes in bf16 (fp16 gradients of 1e-7 would "
                               "underflow without loss scaling); use a bf16 model")
        E, H, hd, I, eps = self._dims()
        B, T, _ = embeds.shape
        dev = embeds.device
        eng = self.m.engine
        cos, sin = eng.rope_tables(T, hd, dev)
        rope = (cos, sin, T, E)
        scale = 1.0 / math.sqrt(hd)
        kmask = attention_mask.to(device=dev, dtype=torch.int32).contiguous() if attention_mask is not None else None
        x = embeds.reshape(B * T, E).contiguous()
        saved = []
        for l in self.m.llm.model.layers:
            sa, mlp = l.self_attn, l.mlp
            g1, g2 = self._w(l.input_layernorm.weight), self._w(l.post_attention_layernorm.weight)
            rstd1 = ops.rms_rstd(x, eps)
            xn1 = ops.rmsnorm(x, g1, eps)
            q = ops.linear(xn1, self._w(sa.q_proj.weight), epi=ops.EPI_ROPE, rope=rope)
            k = ops.linear(xn1, self._w(sa.k_proj.weight), epi=ops.EPI_ROPE, rope=rope)
            v = ops.linear(xn1, self._w(sa.v_proj.weight))
            a = ops.attention(q.view(B, T, H, hd), k.view(B, T, H, hd), v.view(B, T, H, hd), scale=scale, causal=True,
                              key_mask=kmask).view(B * T, E)
            x1 = ops.linear(a, self._w(sa.o_proj.weight), residual=x)
            rstd2 = ops.rms_rstd(x1, eps)
            xn2 = ops.rmsnorm(x1, g2, eps)
            gt = ops.linear(xn2, self._w(mlp.gate_proj.weight))
            up = ops.linear(xn2, self._w(mlp.up_proj.weight))
            h = ops.swiglu_fwd(gt, up)
            x2 = ops.linear(h, self._w(mlp.down_proj.weight), residual=x1)
            saved.append((x, rstd1, xn1, q, k, v, a, x1, rstd2, xn2, gt, up, h))
            x = x2
        llm = self.m.llm
        gn = self._w(llm.model.norm.weight)
        rstdf = ops.rms_rstd(x, eps)
        xnf = ops.rmsnorm(x, gn, eps)
        logits = ops.linear(xnf, self._w(llm.lm_head.weight)).view(B, T, -1)
        labels = labels.to(dev).to(torch.int64).contiguous()
        loss, n_valid = ops.ce_loss_with_count(logits, labels)
        ctx = dict(saved=saved, x_final=x, rstdf=rstdf, xnf=xnf, logits=logits, labels=labels, n_valid=n_valid, kmask=kmask,
                   B=B, T=T, neg_sin=None)
        return loss, ctx

    def backward(self, ctx: dict, grad_loss=1.0, skip=()) -> torch.Tensor:
        """Fills `p.grad` of every LLaMA parameter (accumulating into existing gradients) and returns d loss / d embeds."""
        ops.TAG = "train.bwd"
        E, H, hd, I, eps = self._dims()
        B, T = ctx["B"], ctx["T"]
        llm = self.m.llm
        dev = ctx["logits"].device
        if self.grads is None:
            self.grads = GradBuffer(self.m)
        gb = self.grads
        fresh = gb.attach(skip)
        self._fresh = fresh
        eng = self.m.engine
        cos, sin = eng.rope_tables(T, hd, dev)
        nsin = -sin
        scale = 1.0 / math.sqrt(hd)
        norm_acc = torch.zeros((2 * len(llm.model.layers) + 1, E), device=dev, dtype=torch.float32)

        def dw(dy, xin, p):
            ops.gemm_dw(dy, xin, gb.views[id(p)], accumulate=not fresh[id(p)])

        def norm_grad(row, p):
            g = gb.views[id(p)]
            if fresh[id(p)]:
                g.copy_(norm_acc[row])
            else:
                g.add_(norm_acc[row].to(BF16))

        # ---- CE + lm_head + final norm
        dlogits = ops.ce_bwd(ctx["logits"], ctx["labels"], ctx["n_valid"], grad_loss).view(B * T, -1)
        w_lm = self._w(llm.lm_head.weight)
        dxnf = ops.gemm_dx(dlogits, w_lm)
        dw(dlogits, ctx["xnf"], llm.lm_head.weight)
        dx = ops.rmsnorm_bwd(dxnf, ctx["x_final"], ctx["rstdf"], self._w(llm.model.norm.weight), None, norm_acc[0])
        norm_grad(0, llm.model.norm.weight)
        self._sync_bucket(0)
        del dlogits, dxnf
        ctx["logits"] = None

        # ---- decoder layers, last to first
        layers = list(llm.model.layers)
        for li in range(len(layers) - 1, -1, -1):
            l = layers[li]
            sa, mlp = l.self_attn, l.mlp
            x, rstd1, xn1, q, k, v, a, x1, rstd2, xn2, gt, up, h = ctx["saved"][li]
            ctx["saved"][li] = None
            # MLP: x2 = x1 + down(silu(gate(xn2)) * up(xn2))
            dh = ops.gemm_dx(dx, self._w(mlp.down_proj.weight))
            dw(dx, h, mlp.down_proj.weight)
            dgt, dup = ops.swiglu_bwd(dh, gt, up)
            del dh
            dxn2 = ops.gemm_dx(dgt, self._w(mlp.gate_proj.weight))
            ops.gemm_dx(dup, self._w(mlp.up_proj.weight), out=dxn2, accumulate=True)
            dw(dgt, xn2, mlp.gate_proj.weight)
            dw(dup, xn2, mlp.up_proj.weight)
            del dgt, dup
            r = 1 + 2 * li
            dx1 = ops.rmsnorm_bwd(dxn2, x1, rstd2, self._w(l.post_attention_layernorm.weight), dx, norm_acc[r])
            norm_grad(r, l.post_attention_layernorm.weight)
            # attention: x1 = x + o(attn(rope(q(xn1)), rope(k(xn1)), v(xn1)))
            da = ops.gemm_dx(dx1, self._w(sa.o_proj.weight))
            dw(dx1, a, sa.o_proj.weight)
            dq, dk, dv = ops.attention_bwd(q.view(B, T, H, hd), k.view(B, T, H, hd), v.view(B, T, H, hd),
                                           da.view(B, T, H, hd), scale=scale, causal=True, key_mask=ctx["kmask"])
            dq, dk, dv = dq.view(B * T, E), dk.view(B * T, E), dv.view(B * T, E)
            ops.rope_rows(dq, E, cos, nsin, T)  # the rotation is orthogonal: its transpose is the rotation by -theta
            ops.rope_rows(dk, E, cos, nsin, T)
            dxn1 = ops.gemm_dx(dq, self._w(sa.q_proj.weight))
            ops.gemm_dx(dk, self._w(sa.k_proj.weight), out=dxn1, accumulate=True)
            ops.gemm_dx(dv, self._w(sa.v_proj.weight), out=dxn1, accumulate=True)
            dw(dq, xn1, sa.q_proj.weight)
            dw(dk, xn1, sa.k_proj.weight)
            dw(dv, xn1, sa.v_proj.weight)
            dx = ops.rmsnorm_bwd(dxn1, x, rstd1, self._w(l.input_layernorm.weight), dx1, norm_acc[r + 1])
            norm_grad(r + 1, l.input_layernorm.weight)
            self._sync_bucket(len(layers) - li)
        return dx.view(B, T, E)

    def embed_backward(self, d_embeds: torch.Tensor, inputs: dict, n_prefix: int, prefix_ids: Optional[torch.Tensor]) -> None:
        """Gradient of the embedding table through the gathered rows of inputs_embeds: BOS / text tokens and the
        start / end tokens of every modality block (modeling.py:971-972, 979-980, 996-997, 1019-1020)."""
        gb = self.grads
        table = self.m.llm.model.embed_tokens.weight
        g = gb.views[id(table)]
        if self._fresh[id(table)]:
            g.zero_()  # the scatter adds into the buffer
        B, T, E = d_embeds.shape
        dev = d_embeds.device
        ids = inputs["input_ids"].to(dev)
        ops.embed_scatter_add(d_embeds[:, 0, :], ids[:, 0], g)
        L = ids.shape[1]
        if L > 1:
            txt = d_embeds[:, 1 + n_prefix:, :].reshape(B * (L - 1), E)
            ops.embed_scatter_add(txt, ids[:, 1:].reshape(-1), g)
        if prefix_ids is not None:  # (B, n_prefix) token id per prefix position, -1 for aligned rows
            pre = d_embeds[:, 1:1 + n_prefix, :].reshape(B * n_prefix, E)
            ops.embed_scatter_add(pre, prefix_ids.reshape(-1).to(dev), g)

    def sync_last_bucket(self) -> None:
        self._sync_bucket(len(self.grads.buckets) - 1)

    # ---- gradient all-reduce overlapped with the backward pass
    def _sync_bucket(self, i: int) -> None:
        """All-reduce bucket i (complete on the compute stream as of now) on a side stream, so the collective overlaps the
        backward pass of the layers still to be differentiated."""
        if self.world > 1 and self.overlap_allreduce:
            from . import dist as D

            if D._NCCL_READY and self.grads.flat.is_cuda:
                if getattr(self, "_comm_stream", None) is None:
                    self._comm_stream = torch.cuda.Stream(device=self.grads.flat.device)
                ev = torch.cuda.Event()
                ev.record()
                self._comm_stream.wait_event(ev)
                with torch.cuda.stream(self._comm_stream):
                    allreduce_grads(self.grads, self.world, bucket=i)
                self._comm_used = True
                return
            w = allreduce_grads(self.grads, self.world, bucket=i, async_op=True)
            if w is not None:
                self._pending.append(w)

    def finish_allreduce(self) -> None:
        if self.world > 1 and not self.overlap_allreduce:
            allreduce_grads(self.grads, self.world)
        if getattr(self, "_comm_used", False):
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._comm_used = False
        for w in self._pending:
            w.wait()
        self._pending.clear()


# ---------------------------------------------------------------------------------------------------- alignment backward
class AlignTrainer:
    """Backward of one modality's alignment block (reference: autograd of modeling.py:982-987 / 999-1008 / 1022-1026 —
    Conv1d -> Linear -> nn.MultiheadAttention(Q = modal tokens, K = V = the whole embedding table)) in the ABSORBED form the
    forward kernels execute, from the activations `Engine.align(save=...)` kept:

        out   = ctx W_o^T + b_o
        ctx_h = ctx~_h W_v[h]^T + p_sum_real b_v[h] + p_extra bias_v[h]          ctx~ = P[:, :V] . table
        P     = softmax over V + 2 keys of { q~ . table_v + rb,  extra,  0 }    q~_h = s q_h W_k[h],  rb = s q_h . b_k[h],
        q     = z W_q^T + b_q,   z = y W_t^T + b_t,   y = Conv1d(feats)         extra = s q_h . bias_k[h],  s = 1/sqrt(hd)

    The forward chain is fp16; the backward runs in bf16 (gradients of 1e-6 .. 1e-8 would underflow fp16), on bf16 copies
    of the saved activations.  The two R x V tensors of the backward (P and dS) are bf16; the table receives
    dT += P^T dctx~ + dS^T q~ (its role as values and as keys) on top of the gathered-row gradient."""

    def __init__(self, model, grads: GradBuffer):
        self.m, self.gb = model, grads

    def video_long_backward(self, sv: dict, d_feats: torch.Tensor) -> None:
        """Backward of `video_long_self_attention(x, x, x)` (modeling.py:1078; nn.MultiheadAttention with 8 heads of 96,
        add_bias_kv, add_zero_attn) given d(out) (B, N, P): out_proj, attention backward over the N + 2 keys (batched
        tcgen05 GEMMs + the row-wise softmax-backward kernel), the in-projection; its input (frozen CLIP frame features +
        the sinusoid PE) is a constant, so the chain stops at the in-projection's weight / bias and at bias_k / bias_v."""
        ops.TAG = "train.video_long_bwd"
        m, gb = self.m, self.gb
        mha = m.video_long_self_attention
        g = lambda p: gb.views[id(p)]  # noqa: E731
        B, N, P, H, hd = (sv[k] for k in ("B", "N", "P", "H", "hd"))
        dev = d_feats.device
        d_out = d_feats.reshape(B * N, P).contiguous()
        a2 = sv["a"].reshape(B * N, P)
        da = ops.gemm_dx(d_out, mha.out_proj.weight.detach())
        ops.gemm_dw(d_out, a2, g(mha.out_proj.weight), accumulate=True)
        acc_bo = ops.colsum(d_out, torch.zeros((P,), device=dev, dtype=torch.float32))
        q5 = sv["qkv"].view(B, N + 2, 3, H, hd)
        dq, dk, dv = ops.attention_bwd(q5[:, :N, 0], q5[:, :, 1], q5[:, :, 2], da.view(B, N, H, hd), scale=hd ** -0.5,
                                       causal=False, dropout=sv.get("dropout"))
        dqkv = torch.empty((B * N, 3 * P), device=dev, dtype=BF16)
        ops.add_rows(dq.view(B * N, P), None, dqkv[:, :P])
        for b in range(B):
            ops.add_rows(dk.view(B, N + 2, P)[b, :N], None, dqkv[b * N:(b + 1) * N, P:2 * P])
            ops.add_rows(dv.view(B, N + 2, P)[b, :N], None, dqkv[b * N:(b + 1) * N, 2 * P:])
        ops.gemm_dw(dqkv, sv["xp"], g(mha.in_proj_weight), accumulate=True)
        acc_bin = ops.colsum(dqkv, torch.zeros((3 * P,), device=dev, dtype=torch.float32))
        g(mha.in_proj_bias).add_(acc_bin.to(BF16))
        g(mha.out_proj.bias).add_(acc_bo.to(BF16))
        # the appended (un-projected) bias_k / bias_v rows are key / value N of every sample
        g(mha.bias_k).add_(dk.view(B, N + 2, P)[:, N].float().sum(0).to(BF16).view(1, 1, P))
        g(mha.bias_v).add_(dv.view(B, N + 2, P)[:, N].float().sum(0).to(BF16).view(1, 1, P))

    def backward(self, name: str, sv: dict, d_embeds: torch.Tensor, video_long: Optional[dict] = None) -> None:
        ops.TAG = "train.align_bwd"
        m, gb = self.m, self.gb
        conv = getattr(m, f"project_{name}")
        lin = getattr(m, f"transform_{name}_to_hidden")
        mha = getattr(m, f"{name}_align_attention")
        B, N, C, Lq, kk, ss, H, hd = (sv[k] for k in ("B", "N", "C", "Lq", "kk", "ss", "H", "hd"))
        Nq, E, R = B * Lq, H * hd, H * B * Lq
        dev = d_embeds.device
        table = m.llm.model.embed_tokens.weight.detach()
        V = table.shape[0]
        scale = 1.0 / math.sqrt(hd)
        g = lambda p: gb.views[id(p)]  # noqa: E731  (gradient views; this group is zero-filled when fresh, so always accumulate)
        f32 = lambda n: torch.zeros((n,), device=dev, dtype=torch.float32)  # noqa: E731
        w_in = mha.in_proj_weight.detach()
        b_in = mha.in_proj_bias.detach()
        gw_in, gb_in = g(mha.in_proj_weight), g(mha.in_proj_bias)
        # ---- 0. gradient of the aligned rows, gathered contiguous
        off = sv["row_off"] + 1  # position inside inputs_embeds (BOS is row 0; row_off counts inside the prefix block)
        d_out = torch.empty((Nq, E), device=dev, dtype=BF16)
        for b in range(B):
            ops.add_rows(d_embeds[b, off:off + Lq, :], None, d_out[b * Lq:(b + 1) * Lq])
        # ---- 1. out_proj
        ctx_b, ctxt_b, qt_b, q_b = (ops.cast_bf16(sv[k]) for k in ("ctx", "ctxt", "qt", "q"))
        dctx = ops.gemm_dx(d_out, mha.out_proj.weight.detach())
        ops.gemm_dw(d_out, ctx_b, g(mha.out_proj.weight), accumulate=True)
        acc_bo = ops.colsum(d_out, f32(E))
        # ---- 2. per-head value projection + the two value-side bias terms
        w_v, w_k = w_in[2 * E:], w_in[E:2 * E]
        dctxt = torch.empty((H, Nq, E), device=dev, dtype=BF16)
        ops.gemm_raw(M=Nq, N=E, K=hd, batch=H, A=dctx.data_ptr(), lda=E, a_bs=hd, B=w_v.data_ptr(), ldb=E, b_bs=hd * E,
                     b_mn_major=True, Cout=dctxt.data_ptr(), ldc=E, c_bs=Nq * E)
        gwv = gw_in[2 * E:]
        ops.gemm_raw(M=hd, N=E, K=Nq, batch=H, A=dctx.data_ptr(), lda=E, a_bs=hd, a_mn_major=True, B=ctxt_b.data_ptr(), ldb=E,
                     b_bs=Nq * E, b_mn_major=True, Cout=gwv.data_ptr(), ldc=E, c_bs=hd * E, residual=gwv.data_ptr(), ldr=E,
                     r_bs=hd * E)
        bv2 = torch.stack([b_in[2 * E:], mha.bias_v.detach().reshape(E)], 0).contiguous()
        dstat_v = torch.empty((H, Nq, 2), device=dev, dtype=torch.float32)
        ops.gemm_raw(M=Nq, N=2, K=hd, batch=H, A=dctx.data_ptr(), lda=E, a_bs=hd, B=bv2.data_ptr(), ldb=E, b_bs=hd,
                     Cout=dstat_v.data_ptr(), ldc=2, c_bs=Nq * 2, c_fp32=True)
        dpsr, dpe = dstat_v[..., 0].reshape(R).contiguous(), dstat_v[..., 1].reshape(R).contiguous()
        acc_bv = ops.head_weighted_colsum(dctx, sv["psum"], hd, f32(E))
        acc_biasv = ops.head_weighted_colsum(dctx, sv["pext"], hd, f32(E))
        # ---- 3. through the (V + 2)-key softmax; the table as values (P^T dctx~) and as keys (dS^T q~)
        Vp = sv["P"].shape[1]
        G = torch.empty((R, Vp), device=dev, dtype=torch.float32)
        ops.gemm_raw(M=R, N=V, K=E, A=dctxt.data_ptr(), lda=E, B=table.data_ptr(), ldb=E, Cout=G.data_ptr(), ldc=Vp, c_fp32=True)
        # (with attention dropout: psum / pext above are the DROPPED sums the forward used; the softmax itself needs the
        #  un-dropped p_extra, and the mask is regenerated from the same Philox stream)
        P, dS, dstats = ops.align_softmax_bwd(G, sv["P"], sv["inv_l"], dpsr, sv["pext_raw"], dpe, 1.0, V,
                                              dropout=sv.get("dropout"))
        del G
        dqt = torch.empty((H, Nq, E), device=dev, dtype=BF16)
        ops.gemm_raw(M=R, N=E, K=V, A=dS.data_ptr(), lda=Vp, B=table.data_ptr(), ldb=E, b_mn_major=True, Cout=dqt.data_ptr(), ldc=E)
        gt = g(m.llm.model.embed_tokens.weight)
        for a_, b_ in ((P, dctxt), (dS, qt_b)):
            ops.gemm_raw(M=V, N=E, K=R, A=a_.data_ptr(), lda=Vp, a_mn_major=True, B=b_.data_ptr(), ldb=E, b_mn_major=True,
                         Cout=gt.data_ptr(), ldc=E, residual=gt.data_ptr(), ldr=E)
        del P, dS
        # ---- 4. q~ = s q_h W_k[h], rb = s q_h . b_k[h], extra = s q_h . bias_k[h]
        dq = torch.empty((Nq, E), device=dev, dtype=BF16)
        ds0, ds1 = dstats[0], dstats[1]  # d rb, d extra per row (h, n)
        ds0s, ds1s = (ds0 * scale).contiguous(), (ds1 * scale).contiguous()
        ops.gemm_raw(M=Nq, N=hd, K=E, batch=H, A=dqt.data_ptr(), lda=E, a_bs=Nq * E, B=w_k.data_ptr(), ldb=E, b_bs=hd * E,
                     Cout=dq.data_ptr(), ldc=E, c_bs=hd, alpha=scale, bias=b_in[E:2 * E].data_ptr(), bias_bs=hd,
                     bias_rs=ds0s.data_ptr(), bias2=mha.bias_k.detach().reshape(E).data_ptr(), bias2_rs=ds1s.data_ptr())
        gwk = gw_in[E:2 * E]
        ops.gemm_raw(M=hd, N=E, K=Nq, batch=H, A=q_b.data_ptr(), lda=E, a_bs=hd, a_mn_major=True, B=dqt.data_ptr(), ldb=E,
                     b_bs=Nq * E, b_mn_major=True, Cout=gwk.data_ptr(), ldc=E, c_bs=hd * E, alpha=scale, residual=gwk.data_ptr(),
                     ldr=E, r_bs=hd * E)
        acc_bk = ops.head_weighted_colsum(q_b, ds0s, hd, f32(E))
        acc_biask = ops.head_weighted_colsum(q_b, ds1s, hd, f32(E))
        # ---- 5. q = z W_q^T + b_q ; 6. z = y W_t^T + b_t
        z_b, y_b = ops.cast_bf16(sv["z"]), ops.cast_bf16(sv["y"])
        dz = ops.gemm_dx(dq, w_in[:E])
        ops.gemm_dw(dq, z_b, gw_in[:E], accumulate=True)
        acc_bq = ops.colsum(dq, f32(E))
        dy = ops.gemm_dx(dz, lin.weight.detach())
        ops.gemm_dw(dz, y_b, g(lin.weight), accumulate=True)
        acc_bt = ops.colsum(dz, f32(E))
        # ---- 7. Conv1d weight: dWc[o, k*C + i] = sum_(b,l) dy[b,l,o] feats[b, l*ss + k, i]  (per-sample accumulation)
        feats = sv["feats"]
        dwc = torch.zeros((C, kk * C), device=dev, dtype=BF16)
        for b in range(B):
            ops.gemm_raw(M=C, N=kk * C, K=Lq, A=dy.data_ptr() + b * Lq * C * 2, lda=C, a_mn_major=True,
                         B=feats.data_ptr() + b * feats.stride(0) * 2, ldb=ss * C, b_mn_major=True, Cout=dwc.data_ptr(),
                         ldc=kk * C, residual=dwc.data_ptr(), ldr=kk * C)
        acc_bc = ops.colsum(dy, f32(C))
        # ---- video only: the modal features are the output of the trainable video_long_self_attention -> Conv1d data
        #      gradient (col2im over the overlapping windows), then that block's backward
        if video_long is not None:
            wc = conv.weight.detach().permute(0, 2, 1).reshape(C, kk * C).contiguous()
            dwin = ops.gemm_dx(dy, wc)
            self.video_long_backward(video_long, ops.window_gather_add(dwin, B, N, C, Lq, kk, ss))
        # ---- deposit (torch glue on small tensors: layout permutation of the conv gradient, fp32 -> bf16 bias gradients)
        g(conv.weight).add_(dwc.view(C, kk, C).permute(0, 2, 1))
        g(conv.bias).add_(acc_bc.to(BF16))
        g(lin.bias).add_(acc_bt.to(BF16))
        gb_in[:E].add_(acc_bq.to(BF16))
        gb_in[E:2 * E].add_(acc_bk.to(BF16))
        gb_in[2 * E:].add_(acc_bv.to(BF16))
        g(mha.bias_k).add_(acc_biask.to(BF16).view(1, 1, E))
        g(mha.bias_v).add_(acc_biasv.to(BF16).view(1, 1, E))
        g(mha.out_proj.bias).add_(acc_bo.to(BF16))



class _TrainFn(torch.autograd.Function):
    """Bridges the kernel-library training step into torch autograd: forward returns the loss tensor, backward runs the
    hand-written backward pass and deposits parameter gradients in `p.grad` (views of the flat gradient buffer)."""

    @staticmethod
    def forward(ctx, trainer, inputs, anchor):
        loss, st = trainer._forward_full(inputs)
        ctx.trainer, ctx.st = trainer, st
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, grad_out):
        # the upstream gradient stays on the device (read by the CE-backward kernel): no host sync, graph-capturable
        ctx.trainer._backward_full(ctx.st, grad_out.detach())
        ctx.st = None
        return None, None, None


class TrainStep:
    """`MM_LLMs.forward` in train() mode: the public entry used by MM_LLMs._forward_train."""

    def __init__(self, model):
        self.m = model
        self.llama = LlamaTrainer(model)
        self._anchor = None
        # Attention dropout of the five MHAs (nn.MultiheadAttention(dropout=0.1), modeling.py:879) is live in train() mode,
        # as in the reference.  The Philox seed lives on the device and advances by one per forward, so a captured
        # CUDA graph of the step draws a fresh mask on every replay; `last_seed` is the seed the latest forward used.
        self.attention_dropout = True
        self.dropout_base_seed = 0x5EED
        self._seed = None
        self.last_seed = None

    def set_world(self, world: int, overlap: bool = True) -> None:
        self.llama.world, self.llama.overlap_allreduce = int(world), bool(overlap)

    def __call__(self, inputs: dict):
        if inputs.get("labels") is None:
            raise ValueError("macaw_b200 training: a train()-mode forward needs labels (the loss is what is differentiated)")
        if self._anchor is None or self._anchor.device != self.m.llm.lm_head.weight.device:
            self._anchor = torch.zeros((), device=self.m.llm.lm_head.weight.device, requires_grad=True)
        return _TrainFn.apply(self, inputs, self._anchor)

    def _forward_full(self, inputs: dict):
        m = self.m
        with torch.no_grad():
            # multimodal prefix: frozen encoders -> alignment block, keeping the block's activations for its backward
            saved_align = {}
            seed = None
            if self.attention_dropout:
                dev = m.llm.lm_head.weight.device
                if self._seed is None or self._seed.device != dev:
                    self._seed = torch.tensor([int(self.dropout_base_seed)], dtype=torch.int64, device=dev)
                self._seed.add_(1)
                seed = self.last_seed = self._seed.clone()  # this step's own copy: its backward regenerates the masks from it
            embeds, mask, labels = m.engine.prepare_inputs(inputs, save=saved_align, dropout_seed=seed)
            loss, ctx = self.llama.forward(embeds, mask, labels)
            n_prefix = embeds.shape[1] - inputs["input_ids"].shape[1]
            prefix_ids = None
            if n_prefix > 0:
                B = embeds.shape[0]
                dev = embeds.device
                prefix_ids = torch.full((B, n_prefix), -1, dtype=torch.int64, device=dev)
                off = 0
                for name in ("image", "audio", "video"):
                    key = {"image": "images", "audio": "audios", "video": "videos"}[name]
                    if inputs.get(key) is None:
                        continue
                    Lq = m.engine.last_lens[name]
                    prefix_ids[:, off] = inputs[f"{name}_starts"].to(dev).long()
                    prefix_ids[:, off + 1 + Lq] = inputs[f"{name}_ends"].to(dev).long()
                    off += Lq + 2
            ctx["inputs_ids"] = dict(input_ids=inputs["input_ids"])
            ctx["n_prefix"], ctx["prefix_ids"], ctx["align"] = n_prefix, prefix_ids, saved_align
        return loss, ctx

    def _backward_full(self, ctx: dict, grad_loss) -> None:
        with torch.no_grad():
            if self.llama.grads is None:
                self.llama.grads = GradBuffer(self.m)
            gb = self.llama.grads
            present = [n for n in ALIGN_MODALITIES if n in ctx["align"]]
            skip = [p for n in ALIGN_MODALITIES if n not in present for p in gb.align_params[n]]
            d_embeds = self.llama.backward(ctx, grad_loss, skip=skip)
            fresh = self.llama._fresh
            self.llama.embed_backward(d_embeds, ctx["inputs_ids"], ctx["n_prefix"], ctx["prefix_ids"])
            at = AlignTrainer(self.m, gb)
            for n in present:
                for p in gb.align_params[n]:
                    if fresh[id(p)]:
                        gb.views[id(p)].zero_()
                at.backward(n, ctx["align"][n], d_embeds, video_long=ctx["align"].get("video_long") if n == "video" else None)
            ctx["align"] = None
            self.llama.sync_last_bucket()
