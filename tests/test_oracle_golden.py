"""Pins the CPU oracle (oracle/) against golden vectors produced by transformers'
LlamaForCausalLM (tests/golden/make_golden.py) — CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ops as O
from oracle.model import LlamaDims, LlamaOracle, random_llama_weights

G = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(G, "hf_llama_meta.json")))


@pytest.mark.parametrize("name", sorted(META))
def test_oracle_matches_hf_golden(name):
    d = LlamaDims(**META[name]["dims"])
    w = random_llama_weights(d, seed=META[name]["weights_seed"])
    z = np.load(os.path.join(G, f"hf_llama_{name}.npz"))
    ids = torch.tensor(z["ids"])
    pos = torch.arange(len(ids))
    # fp32 mode: tight agreement with the HF fp32 model
    lg32, _ = LlamaOracle(d, w, "fp32").forward(ids, pos)
    assert np.abs(lg32.numpy() - z["logits_fp32"]).max() < 2e-5
    # bf16 mode: same rounding points as the HF bf16 model -> within 2 bf16 ulps of the logits
    lg16, _ = LlamaOracle(d, w, "bf16").forward(ids, pos)
    ref16 = z["logits_bf16"].astype(np.float32)
    assert np.abs(lg16.numpy() - ref16).max() <= 2 * 2.0 ** -7 * np.abs(ref16).max()
    # and both sit near the fp32 logits (stated fp tolerance of the bf16 path on these models)
    assert np.abs(lg16.numpy() - z["logits_fp32"]).max() < 0.05
    # greedy continuation (fp32 model) reproduced token for token
    n0 = int(z["greedy_prompt_len"])
    out = LlamaOracle(d, w, "fp32").greedy(z["ids"][:n0].tolist(), len(z["greedy_fp32"]))
    assert out == z["greedy_fp32"].tolist()


def test_incremental_decode_equals_full_forward():
    d = LlamaDims(**META["d64_gqa4"]["dims"])
    w = random_llama_weights(d, seed=5)
    o = LlamaOracle(d, w, "bf16")
    ids = torch.arange(3, 30)
    full, _ = o.forward(ids, torch.arange(27))
    part, kv = o.forward(ids[:20], torch.arange(20))
    for t in range(20, 27):
        step, kv = o.forward(ids[t:t + 1], torch.tensor([t]), kv)
        assert torch.equal(step[0], full[t])


def test_kv_swizzle_roundtrip_and_argmax_ties():
    kv = torch.zeros(3, 2, 2, 16, 128)
    row = torch.arange(128.0)
    for tok in range(16):
        O.kv_page_write(kv, 1, 0, 1, tok, row + tok)
    page = O.kv_page_read(kv, 1, 0, 1)
    assert torch.equal(page, row[None] + torch.arange(16.0)[:, None])
    sw = O.kv_swizzle_index(16, 128)
    assert sorted(sw[5].tolist()) == list(range(128))  # a permutation inside the row
    assert (sw[0] == torch.arange(128)).all()  # token 0 is stored unswizzled
    lg = torch.tensor([[0.0, 3.0, 3.0, 1.0], [2.0, 2.0, 2.0, 2.0]])
    assert O.argmax_first(lg).tolist() == [1, 0]


def test_rope_llama3_scaling_matches_vllm_formula():
    import math
    inv = O.rope_inv_freq(64, 500000.0, {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                                          "high_freq_factor": 4.0, "original_max_position_embeddings": 8192})
    base = O.rope_inv_freq(64, 500000.0, None)
    wl = 2 * math.pi / base
    # vllm/model_executor/layers/rotary_embedding/llama3_rope.py:33-54
    smooth = (8192 / wl - 1.0) / (4.0 - 1.0)
    ref = torch.where(wl < 8192 / 4.0, base, torch.where(wl > 8192 / 1.0, base / 32.0,
                                                       (1 - smooth) * base / 32.0 + smooth * base))
    assert torch.allclose(inv, ref, rtol=1e-6, atol=0)


def test_gemma2_oracle_matches_hf_golden():
    """SURVEY.md §8 f1 groundwork: the Gemma-2 restatement (soft-capping, sliding window, GeGLU,
    (1+w) norms, post-norms, scaled tied embeddings) against transformers' Gemma2ForCausalLM"""
    from oracle.gemma2 import Gemma2Dims, Gemma2Oracle, random_gemma2_weights
    meta = json.load(open(os.path.join(G, "hf_gemma2_meta.json")))
    d = Gemma2Dims(**meta["dims"])
    w = random_gemma2_weights(d, seed=meta["weights_seed"])
    z = np.load(os.path.join(G, "hf_gemma2_tiny.npz"))
    ids = torch.tensor(z["ids"])
    pos = torch.arange(len(ids))
    lg32, _ = Gemma2Oracle(d, w, "fp32").forward(ids, pos)
    assert np.abs(lg32.numpy() - z["logits_fp32"]).max() < 2e-5
    lg16, _ = Gemma2Oracle(d, w, "bf16").forward(ids, pos)
    ref16 = z["logits_bf16"].astype(np.float32)
    assert np.abs(lg16.numpy() - ref16).max() <= 4 * 2.0 ** -7 * np.abs(ref16).max()
    n0 = int(z["greedy_prompt_len"])
    assert Gemma2Oracle(d, w, "fp32").greedy(z["ids"][:n0].tolist(), len(z["greedy_fp32"])) == z["greedy_fp32"].tolist()
    # the sliding window really bites: widening it changes the logits of late positions
    d_wide = Gemma2Dims(**{**meta["dims"], "sliding_window": 4096})
    lgw, _ = Gemma2Oracle(d_wide, w, "fp32").forward(ids, pos)
    assert (lgw[:16] - lg32[:16]).abs().max() < 1e-6 and (lgw[32:] - lg32[32:]).abs().max() > 1e-4
