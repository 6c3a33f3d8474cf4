"""oracle/gemma2.py — CPU restatement of the Gemma-2 architecture (the next model
family; BASELINE configs #4 Tower-Plus-9B and #5 Gemma-2-9B-it are Gemma-2-shaped).
TEST INFRASTRUCTURE ONLY (see oracle/ops.py).  Parity target of the B200Q_ARCH_GEMMA2 path of
libb200q (tests/test_gemma2_gpu.py); pinned against transformers' Gemma2ForCausalLM goldens
(tests/test_oracle_golden.py).

Follows vllm/model_executor/models/gemma2.py, i.e. transformers' Gemma2 (modeling_gemma2.py):
  * embeddings scaled by sqrt(hidden) (scale cast to the weight dtype)           :349-360
  * RMSNorm with (1 + w) and ONE rounding: (x_f32 * rsqrt(var+eps) * (1+w)).to()   :49-63
  * four norms per layer: input, post-attention (on the branch output, before the residual add),
    pre-feedforward, post-feedforward                                              :315-346
  * attention: scale = query_pre_attn_scalar**-0.5, logit soft-capping cap*tanh(s/cap), causal,
    alternating sliding-window layers (key j visible iff i - W < j <= i)            :195-226, 236-257
  * GeGLU MLP: gelu_tanh(gate) * up                                               :69-82
  * tied LM head, final logit soft-capping 30*tanh(logits/30)                      :537-540
Scores, soft-cap and softmax are evaluated in fp32 with P rounded to bf16 for PV (what the fused
attention kernels vLLM uses do; HF's eager path rounds the scores to bf16 at every step, which
bounds the tolerance of the bf16 golden).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops
from .ops import _r


@dataclass
class Gemma2Dims:
    hidden: int
    n_layers: int
    n_q_heads: int
    n_kv_heads: int
    head_dim: int
    intermediate: int
    vocab: int
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    query_pre_attn_scalar: float = 256.0
    attn_softcap: Optional[float] = 50.0
    final_softcap: Optional[float] = 30.0
    sliding_window: int = 4096
    layer_types: Optional[List[str]] = None  # default: sliding, full, sliding, ...
    max_pos: int = 8192

    def layer_is_sliding(self, i: int) -> bool:
        if self.layer_types:
            return self.layer_types[i] == "sliding_attention"
        return i % 2 == 0


def gemma_rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float, mode: str) -> torch.Tensor:
    x = x.float()
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return _r(y * (1.0 + w.float()), mode)


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.gelu(x, approximate="tanh")


def attention_softcap(q, k, v, q_positions, scale, softcap, window, mode):
    """q [Tq,n_q,D], k/v [Tk,n_kv,D]; key j visible to query at position p iff j <= p and
    (window is None or j > p - window)"""
    Tq, n_q, D = q.shape
    Tk, n_kv, _ = k.shape
    g = n_q // n_kv
    kk = k.float().repeat_interleave(g, dim=1)
    vv = v.float().repeat_interleave(g, dim=1)
    s = torch.einsum("qhd,khd->hqk", q.float(), kk) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    j = torch.arange(Tk)[None, :]
    p = q_positions.long()[:, None]
    mask = j > p
    if window is not None:
        mask = mask | (j <= p - window)
    s = s.masked_fill(mask[None], float("-inf"))
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    l = e.sum(-1, keepdim=True)
    o = torch.einsum("hqk,khd->qhd", _r(e, mode), vv) / l.permute(1, 0, 2)
    return _r(o, mode)


class Gemma2Oracle:
    def __init__(self, dims: Gemma2Dims, weights: Dict[str, torch.Tensor], mode: str = "bf16"):
        self.d, self.mode = dims, mode
        self.w = {k: (v.to(torch.bfloat16).float() if mode == "bf16" else v.float()) for k, v in weights.items()}
        self.table = ops.rope_table(dims.max_pos, dims.head_dim, dims.rope_theta, None, mode)
        self.scale = dims.query_pre_attn_scalar ** -0.5

    def forward(self, ids: torch.Tensor, positions: torch.Tensor, kv=None, all_logits: bool = True):
        d, mode, w = self.d, self.mode, self.w
        T = ids.shape[0]
        if kv is None:
            kv = [[torch.zeros(0, d.n_kv_heads, d.head_dim), torch.zeros(0, d.n_kv_heads, d.head_dim)]
                  for _ in range(d.n_layers)]
        embed = w["model.embed_tokens.weight"]
        scale = torch.tensor(math.sqrt(d.hidden))
        scale = scale.to(torch.bfloat16).float() if mode == "bf16" else scale
        h = _r(embed[ids.long()] * scale, mode)
        for li in range(d.n_layers):
            p = f"model.layers.{li}."
            x = gemma_rms_norm(h, w[p + "input_layernorm.weight"], d.rms_eps, mode)
            q = ops.linear(x, w[p + "self_attn.q_proj.weight"], mode).view(T, d.n_q_heads, d.head_dim)
            k = ops.linear(x, w[p + "self_attn.k_proj.weight"], mode).view(T, d.n_kv_heads, d.head_dim)
            v = ops.linear(x, w[p + "self_attn.v_proj.weight"], mode).view(T, d.n_kv_heads, d.head_dim)
            q = ops.rope_neox(q, positions, self.table, mode)
            k = ops.rope_neox(k, positions, self.table, mode)
            kv[li][0] = torch.cat([kv[li][0], k], 0)
            kv[li][1] = torch.cat([kv[li][1], v], 0)
            window = d.sliding_window if d.layer_is_sliding(li) else None
            a = attention_softcap(q, kv[li][0], kv[li][1], positions, self.scale, d.attn_softcap, window, mode)
            a = ops.linear(a.reshape(T, -1), w[p + "self_attn.o_proj.weight"], mode)
            a = gemma_rms_norm(a, w[p + "post_attention_layernorm.weight"], d.rms_eps, mode)
            h = _r(h + a, mode)
            x = gemma_rms_norm(h, w[p + "pre_feedforward_layernorm.weight"], d.rms_eps, mode)
            g = ops.linear(x, w[p + "mlp.gate_proj.weight"], mode)
            u = ops.linear(x, w[p + "mlp.up_proj.weight"], mode)
            m = _r(_r(gelu_tanh(g), mode) * u, mode)
            m = ops.linear(m, w[p + "mlp.down_proj.weight"], mode)
            m = gemma_rms_norm(m, w[p + "post_feedforward_layernorm.weight"], d.rms_eps, mode)
            h = _r(h + m, mode)
        h = gemma_rms_norm(h, w["model.norm.weight"], d.rms_eps, mode)
        if not all_logits:
            h = h[-1:]
        logits = ops.linear(h, embed, mode)  # tied LM head
        if d.final_softcap:
            c = d.final_softcap
            logits = _r(_r(torch.tanh(_r(logits / c, mode)), mode) * c, mode)
        return logits, kv

    def greedy(self, prompt_ids: List[int], max_new_tokens: int, return_logits: bool = False):
        ids = torch.tensor(prompt_ids, dtype=torch.int64)
        logits, kv = self.forward(ids, torch.arange(len(prompt_ids)), None, all_logits=False)
        out, lg = [], []
        for _ in range(max_new_tokens):
            t = int(ops.argmax_first(logits[-1:])[0])
            out.append(t)
            if return_logits:
                lg.append(logits[-1].clone())
            if len(out) == max_new_tokens:
                break
            pos = len(prompt_ids) + len(out) - 1
            logits, kv = self.forward(torch.tensor([t]), torch.tensor([pos]), kv, all_logits=False)
        return (out, lg) if return_logits else out


def random_gemma2_weights(d: Gemma2Dims, seed: int = 1234, std: float = 0.02, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    mat = lambda r, c: (torch.randn(r, c, generator=g) * std).to(dtype)
    nw = lambda: (torch.randn(d.hidden, generator=g) * 0.1).to(dtype)  # (1 + w) norms: small random w
    w = {"model.embed_tokens.weight": mat(d.vocab, d.hidden)}
    qd, kd = d.n_q_heads * d.head_dim, d.n_kv_heads * d.head_dim
    for li in range(d.n_layers):
        p = f"model.layers.{li}."
        for n in ("input_layernorm", "post_attention_layernorm", "pre_feedforward_layernorm", "post_feedforward_layernorm"):
            w[p + n + ".weight"] = nw()
        w[p + "self_attn.q_proj.weight"] = mat(qd, d.hidden)
        w[p + "self_attn.k_proj.weight"] = mat(kd, d.hidden)
        w[p + "self_attn.v_proj.weight"] = mat(kd, d.hidden)
        w[p + "self_attn.o_proj.weight"] = mat(d.hidden, qd)
        w[p + "mlp.gate_proj.weight"] = mat(d.intermediate, d.hidden)
        w[p + "mlp.up_proj.weight"] = mat(d.intermediate, d.hidden)
        w[p + "mlp.down_proj.weight"] = mat(d.hidden, d.intermediate)
    w["model.norm.weight"] = nw()
    return w
