"""oracle/ops.py — CPU restatement of every op on the hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this package; the product (``llmq_b200``) never does and has no CPU fallback.

Where the arithmetic lives: the reference (iPieter/llmq) contains no numeric code; every FLOP is
executed by the un-vendored dependency ``vllm>=0.7.0`` (ref:pyproject.toml:15; installed here:
vllm 0.22.0+cu129) behind ref:llmq/workers/vllm_worker.py:105-123,183-186.  These functions
restate vLLM's *native* op definitions (the ones Inductor compiles by default on CUDA), which in
turn are the HF ``transformers`` Llama definitions:

  rms_norm / fused_add_rms_norm   vllm/ir/ops/layernorm.py:9-21, 39-58
                                  (sum rounded to the activation dtype before the variance, as
                                   vLLM's _C kernel and HF LlamaDecoderLayer do)
  rotary (neox)                   vllm/model_executor/layers/rotary_embedding/base.py:140-180,
                                  common.py:144-183 ; llama3 scaling llama3_rope.py:33-54
  SiluAndMul                      vllm/model_executor/layers/activation.py:138-141
  linear                          vllm/model_executor/layers/utils.py:92-98 (F.linear, bf16 out)
  attention                       softmax(QK^T * scale) V, causal, GQA — fp32 scores and
                                  probabilities, probabilities cast to bf16 for the PV product
                                  (what FlashInfer / TRTLLM-gen kernels do)
  greedy sampler                  vllm/v1/sample/sampler.py:91, 235-236 (argmax of the logits)

Pinning: tests/test_oracle_golden.py checks this file against golden vectors produced by
``transformers.LlamaForCausalLM`` on CPU (tests/golden/make_golden.py) and, when present, against
token ids captured from the real vLLM engine on a B200 (tests/golden/vllm_*.json).

Precision modes: ``bf16`` reproduces the rounding points of the bf16 model (weights/activations
bf16, fp32 accumulation, one rounding per op output); ``fp32`` keeps everything in fp32.
"""
from __future__ import annotations

import math

import numpy as np
import torch

BF16 = torch.bfloat16


def _r(x: torch.Tensor, mode: str) -> torch.Tensor:
    """round an fp32 tensor to the activation dtype of `mode` (and come back to fp32 storage)"""
    return x.to(BF16).float() if mode == "bf16" else x


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float, mode: str = "bf16") -> torch.Tensor:
    """x, w: fp32 tensors holding values representable in `mode`'s dtype. Returns same."""
    x = x.float()
    var = x.pow(2).mean(-1, keepdim=True)
    y = _r(x * torch.rsqrt(var + eps), mode)  # x.to(weight.dtype)
    return _r(y * w.float(), mode)


def add_rms_norm(x, residual, w, eps, mode="bf16"):
    """returns (normed, new_residual); new_residual = round(x + residual)"""
    r = _r(x.float() + residual.float(), mode)
    return rms_norm(r, w, eps, mode), r


def rope_inv_freq(head_dim: int, theta: float, scaling: dict | None = None) -> torch.Tensor:
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    if scaling and scaling.get("rope_type", scaling.get("type")) == "llama3":
        factor = scaling["factor"]
        lo, hi = scaling["low_freq_factor"], scaling["high_freq_factor"]
        orig = scaling["original_max_position_embeddings"]
        low_wl, high_wl = orig / lo, orig / hi
        wl = 2 * math.pi / inv
        inv_l = torch.where(wl > low_wl, inv / factor, inv)
        smooth = (orig / wl - lo) / (hi - lo)
        smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
        is_mid = ~(wl < high_wl) * ~(wl > low_wl)
        inv = torch.where(is_mid, smoothed, inv_l)
    return inv


def rope_table(max_pos: int, head_dim: int, theta: float, scaling: dict | None = None,
               mode: str = "bf16") -> torch.Tensor:
    """[max_pos, head_dim] = (cos[D/2] | sin[D/2]); rounded to bf16 in bf16 mode (HF casts
    cos/sin to the activation dtype before use)."""
    inv = rope_inv_freq(head_dim, theta, scaling)
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(pos, inv)  # fp32
    return _r(torch.cat([freqs.cos(), freqs.sin()], dim=-1), mode)


def rope_neox(x: torch.Tensor, positions: torch.Tensor, table: torch.Tensor, mode="bf16"):
    """x: [T, n_heads, D] fp32 storage; rotates pairs (i, i+D/2) with per-op rounding."""
    D = x.shape[-1]
    cs = table[positions.long()]  # [T, D]
    cos, sin = cs[:, None, : D // 2], cs[:, None, D // 2:]
    x1, x2 = x[..., : D // 2], x[..., D // 2:]
    o1 = _r(_r(x1 * cos, mode) - _r(x2 * sin, mode), mode)
    o2 = _r(_r(x2 * cos, mode) + _r(x1 * sin, mode), mode)
    return torch.cat([o1, o2], dim=-1)


def linear(x: torch.Tensor, w: torch.Tensor, mode="bf16") -> torch.Tensor:
    """x [T,K], w [N,K] -> [T,N]; fp32 accumulate, one rounding of the output."""
    return _r(x.float() @ w.float().t(), mode)


def swiglu(gate_up: torch.Tensor, mode="bf16") -> torch.Tensor:
    """gate_up [T, 2I] = (g | u) -> round(round(silu(g)) * u)"""
    I = gate_up.shape[-1] // 2
    g, u = gate_up[..., :I].float(), gate_up[..., I:].float()
    s = _r(torch.nn.functional.silu(g), mode)
    return _r(s * u, mode)


def attention(q, k, v, q_positions, scale, mode="bf16"):
    """Causal GQA attention of one sequence.
    q: [Tq, n_q, D]; k, v: [Tk, n_kv, D] (keys 0..Tk-1); q_positions: [Tq] absolute positions.
    Key j is visible to query i iff j <= q_positions[i]."""
    Tq, n_q, D = q.shape
    Tk, n_kv, _ = k.shape
    g = n_q // n_kv
    kk = k.float().repeat_interleave(g, dim=1)  # [Tk, n_q, D]
    vv = v.float().repeat_interleave(g, dim=1)
    s = torch.einsum("qhd,khd->hqk", q.float(), kk) * scale  # fp32
    mask = torch.arange(Tk)[None, :] > q_positions.long()[:, None]  # [Tq, Tk]
    s = s.masked_fill(mask[None], float("-inf"))
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = torch.einsum("hqk,khd->qhd", _r(p, mode), vv) / l.permute(1, 0, 2)
    return _r(o, mode)


def argmax_first(logits: torch.Tensor) -> np.ndarray:
    """argmax over the last dim, lowest index on ties (numpy semantics == torch CUDA argmax)."""
    return np.argmax(logits.float().numpy(), axis=-1).astype(np.int32)


# ---- paged-KV layout helpers (mirror of the layout in llmq_b200/csrc/attn.cu) -----------------
def kv_swizzle_index(block_size: int, D: int) -> torch.Tensor:
    """index map idx[t, d] = position inside a [block_size, D] page row where logical element
    (token t, dim d) is stored: 16-byte chunk c = d // 8 lives at chunk c ^ (t & 7)."""
    t = torch.arange(block_size)[:, None]
    d = torch.arange(D)[None, :]
    return (((d // 8) ^ (t & 7)) * 8 + (d % 8)).long()


def kv_page_read(kv_layer: torch.Tensor, block: int, which: int, head: int) -> torch.Tensor:
    """kv_layer: [num_blocks, 2, n_kv, BS, D] raw (swizzled) -> logical [BS, D] page"""
    page = kv_layer[block, which, head]
    BS, D = page.shape
    return torch.gather(page, 1, kv_swizzle_index(BS, D))


def kv_page_write(kv_layer: torch.Tensor, block: int, which: int, head: int, tok: int,
                  row: torch.Tensor) -> None:
    BS, D = kv_layer.shape[-2:]
    idx = kv_swizzle_index(BS, D)[tok]
    kv_layer[block, which, head, tok, idx] = row.to(kv_layer.dtype)
