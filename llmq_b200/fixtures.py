"""Synthetic model directories, tokenizers and JSONL job files.

No tokenizer, config or checkpoint exists on disk in the build/bench environment and there is no
network, so both the native worker and the oracle harness run on locally synthesised assets:

  * a `tokenizers` WordLevel tokenizer whose vocabulary has exactly `vocab_size` entries
    ("w<i>" for ordinary ids, Llama-3's <|begin_of_text|> / <|end_of_text|> / <|eot_id|> at their
    usual ids when the vocabulary is large enough), whitespace pre-tokenisation, a BOS-prepending
    post-processor and a Llama-3-style chat template.  "w17 w4 w99" tokenises to exactly
    [BOS, 17, 4, 99] and decodes back to the same text, so prompt length is exact.
  * config.json / generation_config.json with the real Llama dimensions,
  * optional seeded safetensors weights (Normal(0, 0.02), unit norms) for small models,
  * JSONL job files: {"id": "job-0000001", "prompt": "<127 random words>"}  => 128 prompt ids.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional

import numpy as np

from .model import ModelSpec

CHAT_TEMPLATE = (
    "{{ bos_token }}{% for message in messages %}"
    "{{ '<|start_header_id|> ' + message['role'] + ' <|end_header_id|> ' + message['content'] + ' <|eot_id|> ' }}"
    "{% endfor %}{% if add_generation_prompt %}{{ '<|start_header_id|> assistant <|end_header_id|> ' }}{% endif %}"
)


def special_token_ids(vocab_size: int) -> dict:
    if vocab_size >= 128256:  # Llama-3 layout
        return {"<|begin_of_text|>": 128000, "<|end_of_text|>": 128001, "<|start_header_id|>": 128006,
                "<|end_header_id|>": 128007, "<|eot_id|>": 128009}
    # small test vocabularies: specials at the low ids
    return {"<|begin_of_text|>": 0, "<|end_of_text|>": 1, "<|eot_id|>": 2,
            "<|start_header_id|>": vocab_size - 2, "<|end_header_id|>": vocab_size - 1}


def build_tokenizer(vocab_size: int):
    """-> transformers.PreTrainedTokenizerFast"""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast

    sp = special_token_ids(vocab_size)
    by_id = {i: f"w{i}" for i in range(vocab_size)}
    for tok, i in sp.items():
        by_id[i] = tok
    vocab = {tok: i for i, tok in by_id.items()}
    unk = "<|end_of_text|>"
    tk = Tokenizer(models.WordLevel(vocab=vocab, unk_token=unk))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    bos = "<|begin_of_text|>"
    tk.post_processor = processors.TemplateProcessing(single=f"{bos} $A", special_tokens=[(bos, sp[bos])])
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, bos_token=bos, eos_token="<|end_of_text|>",
                                   unk_token=unk, additional_special_tokens=[t for t in sp if t not in (bos, unk)])
    fast.chat_template = CHAT_TEMPLATE
    return fast


def seeded_state_dict(spec: ModelSpec, seed: int = 1234) -> dict:
    """HF-named bf16 state dict: Normal(0, 0.02) matrices drawn in a fixed order from one seeded
    generator, unit norm weights.  The same call reproduces the checkpoint anywhere (tests do not
    need the safetensors file that was handed to vLLM)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    mat = lambda r, c: (torch.randn(r, c, generator=g) * 0.02).to(torch.bfloat16)
    gemma = spec.arch == "gemma2"
    if gemma:  # GemmaRMSNorm scales by (1 + w): small random w so that the weights matter
        ones = lambda: (torch.randn(spec.hidden, generator=g) * 0.1).to(torch.bfloat16)
        norms = ("input_layernorm", "post_attention_layernorm", "pre_feedforward_layernorm",
                 "post_feedforward_layernorm")
    else:
        ones = lambda: torch.ones(spec.hidden, dtype=torch.bfloat16)
        norms = ("input_layernorm", "post_attention_layernorm")
    sd = {"model.embed_tokens.weight": mat(spec.vocab, spec.hidden)}
    qd, kd = spec.n_q_heads * spec.head_dim, spec.n_kv_heads * spec.head_dim
    for i in range(spec.n_layers):
        p = f"model.layers.{i}."
        for n in norms:
            sd[p + n + ".weight"] = ones()
        sd[p + "self_attn.q_proj.weight"] = mat(qd, spec.hidden)
        sd[p + "self_attn.k_proj.weight"] = mat(kd, spec.hidden)
        sd[p + "self_attn.v_proj.weight"] = mat(kd, spec.hidden)
        sd[p + "self_attn.o_proj.weight"] = mat(spec.hidden, qd)
        sd[p + "mlp.gate_proj.weight"] = mat(spec.intermediate, spec.hidden)
        sd[p + "mlp.up_proj.weight"] = mat(spec.intermediate, spec.hidden)
        sd[p + "mlp.down_proj.weight"] = mat(spec.hidden, spec.intermediate)
    sd["model.norm.weight"] = ones()
    if not spec.tie_embeddings:
        sd["lm_head.weight"] = mat(spec.vocab, spec.hidden)
    return sd


def write_model_dir(path: str, spec: ModelSpec, *, seed: int = 1234, with_weights: bool = True) -> str:
    """config.json + tokenizer (+ safetensors when with_weights) in `path`."""
    os.makedirs(path, exist_ok=True)
    sp = special_token_ids(spec.vocab)
    cfg = spec.to_hf_config()
    cfg["bos_token_id"] = sp["<|begin_of_text|>"]
    cfg["eos_token_id"] = sp["<|end_of_text|>"]
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    with open(os.path.join(path, "generation_config.json"), "w") as f:
        json.dump({"bos_token_id": cfg["bos_token_id"], "eos_token_id": cfg["eos_token_id"]}, f)
    build_tokenizer(spec.vocab).save_pretrained(path)
    if with_weights:
        from safetensors.torch import save_file

        save_file(seeded_state_dict(spec, seed), os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
    return path


def random_prompt_words(rng: np.random.Generator, n_tokens: int, vocab_size: int) -> str:
    """n_tokens ordinary (non-special) word tokens rendered as text"""
    sp = set(special_token_ids(vocab_size).values())
    hi = min(vocab_size, 128000)
    ids = rng.integers(3 if vocab_size < 128256 else 0, hi, size=n_tokens)
    return " ".join(f"w{int(i)}" for i in ids if int(i) not in sp or True)


def make_jobs(n_jobs: int, vocab_size: int, prompt_tokens: int = 127, seed: int = 20260921,
              start: int = 0, stride: int = 1) -> List[dict]:
    """jobs start, start+stride, ... of the canonical seeded job stream (job i is the same no
    matter how the stream is sharded across workers)"""
    out = []
    for i in range(start, n_jobs, stride):
        rng = np.random.default_rng([seed, i])
        out.append({"id": f"job-{i:07d}", "prompt": random_prompt_words(rng, prompt_tokens, vocab_size)})
    return out


def write_jobs_jsonl(path: str, n_jobs: int, vocab_size: int, prompt_tokens: int = 127,
                     seed: int = 20260921) -> str:
    with open(path, "w") as f:
        for j in make_jobs(n_jobs, vocab_size, prompt_tokens, seed):
            f.write(json.dumps(j) + "\n")
    return path


class _NoDeviceModel:
    """stands where Engine.model is: no CUDA device, nothing to close"""
    device = None

    def __init__(self, max_model_len: int):
        self.max_model_len = max_model_len

    def close(self):
        pass


class DryRunEngine:
    """llmq_b200.model.Engine's call surface over `b200q_engine_create_dryrun`: the real C++
    scheduler, paged-KV block manager and per-step metadata packing, no model and no CUDA; every
    "sampled" token is (previous token + 1) mod vocab.  For host-logic tests and
    tools/host_path_bench.py only — it cannot generate text from a model and the worker never
    constructs it.  `step_ms` stands in for the device time of a step (sleep, GIL released)."""

    def __init__(self, vocab: int, max_num_seqs: int, max_num_batched_tokens: int, max_model_len: int,
                 num_blocks: int, eos_token_id=None, policy: int = 1, step_ms: float = 0.0):
        import ctypes as C

        from . import lib as L

        self._C, self._L = C, L
        self.lib = L.load()
        stop_ids = [] if eos_token_id is None else (
            [int(eos_token_id)] if isinstance(eos_token_id, int) else sorted({int(x) for x in eos_token_id}))
        cfg = L.EngineConfig(max_num_seqs=max_num_seqs, max_num_batched_tokens=max_num_batched_tokens,
                             max_model_len=max_model_len,
                             eos_token_id=stop_ids[0] if stop_ids else -1, policy=policy)
        h = C.c_void_p()
        L.check(self.lib.b200q_engine_create_dryrun(C.byref(cfg), vocab, 16, num_blocks, C.byref(h)))
        if len(stop_ids) > 1:
            L.check(self.lib.b200q_engine_set_stop_ids(h, (C.c_int32 * len(stop_ids))(*stop_ids), len(stop_ids)))
        self.max_model_len = max_model_len
        self.handle, self.cap = h, max_num_seqs
        self._ids = np.zeros(self.cap, np.int64)
        self._tok = np.zeros(self.cap, np.int32)
        self._flg = np.zeros(self.cap, np.int32)
        self.model = _NoDeviceModel(max_model_len)
        self.step_s = step_ms / 1e3
        self.steps = 0

    def add_request(self, req_id, prompt_ids, max_new_tokens, ignore_eos=False, temperature=0.0, seed=0):
        arr = np.ascontiguousarray(prompt_ids, dtype=np.int32)
        rc = self.lib.b200q_engine_add_request(self.handle, int(req_id), arr.ctypes.data, arr.size,
                                               int(max_new_tokens), int(ignore_eos))
        if rc == -1:
            raise ValueError(self.lib.b200q_last_error().decode())
        self._L.check(rc)

    def abort(self, req_id):
        self._L.check(self.lib.b200q_engine_abort(self.handle, int(req_id)))

    def set_async(self, on: bool):
        self._L.check(self.lib.b200q_engine_set_async(self.handle, int(on)))

    def has_work(self) -> bool:
        return bool(self.lib.b200q_engine_has_work(self.handle))

    def step(self):
        n = self._C.c_int32(0)
        self._L.check(self.lib.b200q_engine_step(self.handle, self._ids.ctypes.data, self._tok.ctypes.data,
                                                 self._flg.ctypes.data, self.cap, self._C.byref(n)))
        if self.step_s:
            import time

            time.sleep(self.step_s)
        self.steps += 1
        k = n.value
        return self._ids[:k], self._tok[:k], self._flg[:k]

    def stats(self):
        s = self._L.EngineStats()
        self._L.check(self.lib.b200q_engine_get_stats(self.handle, self._C.byref(s)))
        return s

    def close(self):
        if self.handle:
            self.lib.b200q_engine_destroy(self.handle)
            self.handle = None

