"""oracle/model.py — whole-model CPU restatement (Llama architecture) + greedy generation.
TEST INFRASTRUCTURE ONLY (see oracle/ops.py for the import rules and the pinning story).

Layer wiring follows vllm/model_executor/models/llama.py:117-121 (MLP), 223-231 (attention),
316-333 (decoder layer: residual handling), 395-431 (model: embed -> layers -> final norm) and
the LM head / greedy sampler of vllm/model_executor/layers/logits_processor.py:89 and
vllm/v1/sample/sampler.py:91,235-236 (logits are produced in the model dtype, i.e. rounded to
bf16, then argmax'ed in fp32).

Weights use the HF checkpoint names (model.layers.N.self_attn.q_proj.weight, ...), so the same
state dict drives transformers (golden generator), this oracle and the CUDA engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops


@dataclass
class LlamaDims:
    hidden: int
    n_layers: int
    n_q_heads: int
    n_kv_heads: int
    head_dim: int
    intermediate: int
    vocab: int
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    tie_embeddings: bool = False
    max_pos: int = 8192

    @staticmethod
    def from_hf_config(cfg: dict) -> "LlamaDims":
        hd = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
        rs = cfg.get("rope_scaling") or (cfg.get("rope_parameters") if isinstance(cfg.get("rope_parameters"), dict) and cfg["rope_parameters"].get("rope_type", "default") != "default" else None)
        theta = cfg.get("rope_theta")
        if theta is None and isinstance(cfg.get("rope_parameters"), dict):
            theta = cfg["rope_parameters"].get("rope_theta")
        return LlamaDims(
            hidden=cfg["hidden_size"], n_layers=cfg["num_hidden_layers"],
            n_q_heads=cfg["num_attention_heads"],
            n_kv_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]),
            head_dim=hd, intermediate=cfg["intermediate_size"], vocab=cfg["vocab_size"],
            rms_eps=cfg.get("rms_norm_eps", 1e-5), rope_theta=float(theta or 10000.0),
            rope_scaling=rs, tie_embeddings=bool(cfg.get("tie_word_embeddings", False)),
            max_pos=cfg.get("max_position_embeddings", 8192),
        )


class LlamaOracle:
    """Reference forward with a plain (unpaged) per-sequence KV cache."""

    def __init__(self, dims: LlamaDims, weights: Dict[str, torch.Tensor], mode: str = "bf16",
                 max_pos: Optional[int] = None):
        assert mode in ("bf16", "fp32")
        self.d, self.mode = dims, mode
        # keep weights as fp32 storage of (bf16-representable, in bf16 mode) values
        self.w = {k: (v.to(torch.bfloat16).float() if mode == "bf16" else v.float())
                  for k, v in weights.items()}
        self.table = ops.rope_table(max_pos or dims.max_pos, dims.head_dim, dims.rope_theta,
                                    dims.rope_scaling, mode)
        self.scale = dims.head_dim ** -0.5

    def _w(self, name):
        return self.w[name]

    def forward(self, ids: torch.Tensor, positions: torch.Tensor,
                kv: Optional[List[List[torch.Tensor]]] = None, all_logits: bool = True):
        """One sequence. ids/positions: [T]. kv: per-layer [k, v] tensors [Tk, n_kv, D] of the
        already-processed prefix (mutated: new keys appended).  Returns logits [T or 1, V]."""
        d, mode = self.d, self.mode
        T = ids.shape[0]
        if kv is None:
            kv = [[torch.zeros(0, d.n_kv_heads, d.head_dim), torch.zeros(0, d.n_kv_heads, d.head_dim)]
                  for _ in range(d.n_layers)]
        residual = self._w("model.embed_tokens.weight")[ids.long()]
        x = None
        for li in range(d.n_layers):
            p = f"model.layers.{li}."
            if li == 0:
                x = ops.rms_norm(residual, self._w(p + "input_layernorm.weight"), d.rms_eps, mode)
            else:
                x, residual = ops.add_rms_norm(x, residual, self._w(p + "input_layernorm.weight"),
                                               d.rms_eps, mode)
            q = ops.linear(x, self._w(p + "self_attn.q_proj.weight"), mode).view(T, d.n_q_heads, d.head_dim)
            k = ops.linear(x, self._w(p + "self_attn.k_proj.weight"), mode).view(T, d.n_kv_heads, d.head_dim)
            v = ops.linear(x, self._w(p + "self_attn.v_proj.weight"), mode).view(T, d.n_kv_heads, d.head_dim)
            q = ops.rope_neox(q, positions, self.table, mode)
            k = ops.rope_neox(k, positions, self.table, mode)
            kv[li][0] = torch.cat([kv[li][0], k], 0)
            kv[li][1] = torch.cat([kv[li][1], v], 0)
            a = ops.attention(q, kv[li][0], kv[li][1], positions, self.scale, mode)
            x = ops.linear(a.reshape(T, -1), self._w(p + "self_attn.o_proj.weight"), mode)
            x, residual = ops.add_rms_norm(x, residual, self._w(p + "post_attention_layernorm.weight"),
                                           d.rms_eps, mode)
            gu = torch.cat([ops.linear(x, self._w(p + "mlp.gate_proj.weight"), mode),
                            ops.linear(x, self._w(p + "mlp.up_proj.weight"), mode)], -1)
            x = ops.linear(ops.swiglu(gu, mode), self._w(p + "mlp.down_proj.weight"), mode)
        x, residual = ops.add_rms_norm(x, residual, self._w("model.norm.weight"), d.rms_eps, mode)
        if not all_logits:
            x = x[-1:]
        head = self._w("model.embed_tokens.weight") if d.tie_embeddings else self._w("lm_head.weight")
        return ops.linear(x, head, mode), kv

    def greedy(self, prompt_ids: List[int], max_new_tokens: int, eos_id: Optional[int] = None,
               return_logits: bool = False):
        """Greedy decode of one prompt; returns generated ids (EOS included if hit)."""
        ids = torch.tensor(prompt_ids, dtype=torch.int64)
        pos = torch.arange(len(prompt_ids))
        logits, kv = self.forward(ids, pos, None, all_logits=False)
        out, lg = [], []
        for _ in range(max_new_tokens):
            t = int(ops.argmax_first(logits[-1:])[0])
            out.append(t)
            if return_logits:
                lg.append(logits[-1].clone())
            if eos_id is not None and t == eos_id:
                break
            if len(out) == max_new_tokens:
                break
            p = len(prompt_ids) + len(out) - 1
            logits, kv = self.forward(torch.tensor([t]), torch.tensor([p]), kv, all_logits=False)
        return (out, lg) if return_logits else out


def random_llama_weights(dims: LlamaDims, seed: int = 1234, std: float = 0.02,
                         dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint in HF naming: Normal(0, std) matrices, unit norm weights
    (not vLLM's near-zero dummy init, which makes argmax tie-fragile)."""
    g = torch.Generator().manual_seed(seed)

    def mat(r, c):
        return (torch.randn(r, c, generator=g) * std).to(dtype)

    w = {"model.embed_tokens.weight": mat(dims.vocab, dims.hidden)}
    qd, kd = dims.n_q_heads * dims.head_dim, dims.n_kv_heads * dims.head_dim
    for li in range(dims.n_layers):
        p = f"model.layers.{li}."
        w[p + "input_layernorm.weight"] = torch.ones(dims.hidden, dtype=dtype)
        w[p + "post_attention_layernorm.weight"] = torch.ones(dims.hidden, dtype=dtype)
        w[p + "self_attn.q_proj.weight"] = mat(qd, dims.hidden)
        w[p + "self_attn.k_proj.weight"] = mat(kd, dims.hidden)
        w[p + "self_attn.v_proj.weight"] = mat(kd, dims.hidden)
        w[p + "self_attn.o_proj.weight"] = mat(dims.hidden, qd)
        w[p + "mlp.gate_proj.weight"] = mat(dims.intermediate, dims.hidden)
        w[p + "mlp.up_proj.weight"] = mat(dims.intermediate, dims.hidden)
        w[p + "mlp.down_proj.weight"] = mat(dims.hidden, dims.intermediate)
    w["model.norm.weight"] = torch.ones(dims.hidden, dtype=dtype)
    if not dims.tie_embeddings:
        w["lm_head.weight"] = mat(dims.vocab, dims.hidden)
    return w
