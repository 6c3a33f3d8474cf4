"""Generate the golden vectors that pin the CPU oracle (run in the build container, CPU only):

    python tests/golden/make_golden.py

Source of truth: `transformers.LlamaForCausalLM` (transformers 5.5.0, eager attention, CPU) — the
definition vLLM's Llama implementation is validated against, and importable here; vLLM itself
needs a GPU (its outputs, captured on a B200, are pinned separately in tests/golden/vllm_*.json).
Seeded synthetic checkpoints (oracle.model.random_llama_weights) are loaded into the HF model;
we store fp32 logits for a fixed token sequence, bf16-model logits for the same sequence and the
greedy continuation of the fp32 model.  Small on purpose (<1 MB).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.model import LlamaDims, random_llama_weights  # noqa: E402

CASES = {
    "d64_gqa4": dict(hidden=256, n_layers=2, n_q_heads=8, n_kv_heads=2, head_dim=64, intermediate=512,
                     vocab=512, max_pos=256),
    "d128_tied_llama3rope": dict(hidden=256, n_layers=2, n_q_heads=4, n_kv_heads=1, head_dim=128,
                                 intermediate=384, vocab=512, max_pos=256, tie_embeddings=True,
                                 rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0,
                                               "high_freq_factor": 4.0, "original_max_position_embeddings": 32}),
}


def hf_model(d: LlamaDims, w, dtype):
    from transformers import LlamaConfig, LlamaForCausalLM

    kw = {}
    if d.rope_scaling:
        kw["rope_scaling"] = dict(d.rope_scaling)
    cfg = LlamaConfig(hidden_size=d.hidden, num_hidden_layers=d.n_layers, num_attention_heads=d.n_q_heads,
                      num_key_value_heads=d.n_kv_heads, head_dim=d.head_dim, intermediate_size=d.intermediate,
                      vocab_size=d.vocab, rms_norm_eps=d.rms_eps, rope_theta=d.rope_theta,
                      max_position_embeddings=d.max_pos, tie_word_embeddings=d.tie_embeddings,
                      attn_implementation="eager", **kw)
    m = LlamaForCausalLM(cfg).eval()
    sd = {k: v.float() for k, v in w.items()}
    if d.tie_embeddings:
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    m.load_state_dict(sd, strict=True)
    return m.to(dtype)


GEMMA2_CASE = dict(hidden=256, n_layers=4, n_q_heads=4, n_kv_heads=2, head_dim=64, intermediate=512, vocab=512,
                   query_pre_attn_scalar=64.0, sliding_window=16, max_pos=256)


def gemma2_golden(out_dir):
    """Gemma-2 (the next architecture): HF Gemma2ForCausalLM on CPU -> goldens for
    oracle/gemma2.py; 48 tokens > the 16-token sliding window, so both layer types are exercised"""
    from transformers import Gemma2Config, Gemma2ForCausalLM

    from oracle.gemma2 import Gemma2Dims, random_gemma2_weights

    d = Gemma2Dims(**GEMMA2_CASE)
    w = random_gemma2_weights(d, seed=3)
    cfg = Gemma2Config(hidden_size=d.hidden, num_hidden_layers=d.n_layers, num_attention_heads=d.n_q_heads,
                       num_key_value_heads=d.n_kv_heads, head_dim=d.head_dim, intermediate_size=d.intermediate,
                       vocab_size=d.vocab, sliding_window=d.sliding_window,
                       query_pre_attn_scalar=int(d.query_pre_attn_scalar), max_position_embeddings=d.max_pos,
                       attn_implementation="eager")
    m = Gemma2ForCausalLM(cfg).eval()
    sd = {k: v.float() for k, v in w.items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    m.load_state_dict(sd, strict=True)
    ids = torch.tensor(np.random.default_rng(0).integers(0, d.vocab, size=48)[None])
    with torch.no_grad():
        l32 = m(ids).logits[0].numpy()
        gen = m.generate(ids[:, :20], max_new_tokens=10, do_sample=False, pad_token_id=0)[0, 20:].numpy()
        l16 = m.to(torch.bfloat16)(ids).logits[0].float().numpy()
    np.savez_compressed(os.path.join(out_dir, "hf_gemma2_tiny.npz"), ids=ids[0].numpy(),
                        logits_fp32=l32.astype(np.float32), logits_bf16=l16.astype(np.float16),
                        greedy_prompt_len=20, greedy_fp32=gen)
    return {"dims": GEMMA2_CASE, "weights_seed": 3}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    meta = {}
    for name, kw in CASES.items():
        d = LlamaDims(**kw)
        w = random_llama_weights(d, seed=11)
        rng = np.random.default_rng(3)
        ids = torch.tensor(rng.integers(0, d.vocab, size=40)[None])
        with torch.no_grad():
            m32 = hf_model(d, w, torch.float32)
            logits32 = m32(ids).logits[0].numpy()
            gen = m32.generate(ids[:, :9], max_new_tokens=12, do_sample=False, pad_token_id=0)[0, 9:].numpy()
            logits16 = hf_model(d, w, torch.bfloat16)(ids).logits[0].float().numpy()
        np.savez_compressed(os.path.join(out_dir, f"hf_llama_{name}.npz"), ids=ids[0].numpy(),
                            logits_fp32=logits32.astype(np.float32), logits_bf16=logits16.astype(np.float16),
                            greedy_prompt_len=9, greedy_fp32=gen)
        meta[name] = {"dims": kw, "weights_seed": 11, "transformers": __import__("transformers").__version__,
                      "torch": torch.__version__}
    with open(os.path.join(out_dir, "hf_llama_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(out_dir, "hf_gemma2_meta.json"), "w") as f:
        json.dump(gemma2_golden(out_dir), f, indent=1)
    print("wrote", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main()
