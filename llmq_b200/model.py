"""Model-side host code of the b200 worker: config parsing, weight loading / synthesis, buffer
allocation (PyTorch = allocator only) and binding into libb200q.

Replaces what ``AsyncEngineArgs(model=...)`` + ``AsyncLLMEngine.from_engine_args`` do for the
reference worker (ref:llmq/workers/vllm_worker.py:105-123): resolve the model, load bf16
weights, size the paged KV pool from ``gpu_memory_utilization``.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import lib as L

BLOCK_SIZE = 16


def _as_id_tuple(v) -> Tuple[int, ...]:
    """eos_token_id in HF configs is an int, a list of ints, or absent"""
    if v is None:
        return ()
    if isinstance(v, (list, tuple)):
        return tuple(int(x) for x in v if x is not None)
    return (int(v),)


@dataclass
class ModelSpec:
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
    max_position_embeddings: int = 8192
    eos_token_id: Optional[int] = None
    bos_token_id: Optional[int] = None
    # every id that ends generation (config.json lists + generation_config.json); eos_token_id is
    # the first of them.  vLLM stops on tokenizer.eos_token_id plus all generation_config ids.
    eos_token_ids: Tuple[int, ...] = ()
    name: str = "llama"
    # architecture: "llama" (Llama / Mistral-style decoders) or "gemma2" 
    arch: str = "llama"
    query_pre_attn_scalar: Optional[float] = None  # gemma2: attention scale = this ** -0.5
    attn_softcap: float = 0.0                      # gemma2: 50
    final_softcap: float = 0.0                     # gemma2: 30
    sliding_window: int = 0                        # gemma2: 4096 on the even layers

    @property
    def attn_scale(self) -> float:
        return float(self.query_pre_attn_scalar or self.head_dim) ** -0.5

    @property
    def embed_scale(self) -> float:
        """gemma2 multiplies the embeddings by sqrt(hidden), the factor itself rounded to bf16
        (transformers Gemma2TextScaledWordEmbedding / vllm gemma2.py normalizer)"""
        if self.arch != "gemma2":
            return 0.0
        return float(torch.tensor(self.hidden ** 0.5).to(torch.bfloat16))

    @staticmethod
    def from_hf_config(cfg: dict, name: str = "llama") -> "ModelSpec":
        arch = (cfg.get("architectures") or ["LlamaForCausalLM"])[0]
        if arch == "Gemma2ForCausalLM":
            return ModelSpec._from_gemma2_config(cfg, name)
        if arch not in ("LlamaForCausalLM", "MistralForCausalLM"):
            raise ValueError(f"unsupported architecture {arch!r}: the b200 worker implements Llama-style "
                             "and Gemma-2 decoders")
        hd = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
        rp = cfg.get("rope_parameters") if isinstance(cfg.get("rope_parameters"), dict) else None
        theta = cfg.get("rope_theta") or (rp or {}).get("rope_theta") or 10000.0
        scaling = cfg.get("rope_scaling")
        if scaling is None and rp and rp.get("rope_type", "default") not in ("default", None):
            scaling = rp
        eos_all = _as_id_tuple(cfg.get("eos_token_id"))
        eos = eos_all[0] if eos_all else None
        return ModelSpec(
            hidden=cfg["hidden_size"], n_layers=cfg["num_hidden_layers"],
            n_q_heads=cfg["num_attention_heads"],
            n_kv_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]),
            head_dim=hd, intermediate=cfg["intermediate_size"], vocab=cfg["vocab_size"],
            rms_eps=cfg.get("rms_norm_eps", 1e-5), rope_theta=float(theta), rope_scaling=scaling,
            tie_embeddings=bool(cfg.get("tie_word_embeddings", False)),
            max_position_embeddings=cfg.get("max_position_embeddings", 8192),
            eos_token_id=eos, eos_token_ids=eos_all, bos_token_id=cfg.get("bos_token_id"), name=name,
            sliding_window=int(cfg.get("sliding_window") or 0) if arch == "MistralForCausalLM" else 0,
        )

    @staticmethod
    def _from_gemma2_config(cfg: dict, name: str) -> "ModelSpec":
        if cfg.get("hidden_activation", cfg.get("hidden_act", "gelu_pytorch_tanh")) != "gelu_pytorch_tanh":
            raise ValueError("gemma2: only hidden_activation=gelu_pytorch_tanh is implemented")
        lt = cfg.get("layer_types")
        if lt and any((t == "sliding_attention") != (i % 2 == 0) for i, t in enumerate(lt)):
            raise ValueError("gemma2: only the stock layer pattern (even layers sliding) is implemented")
        rp = cfg.get("rope_parameters") if isinstance(cfg.get("rope_parameters"), dict) else None
        eos_all = _as_id_tuple(cfg.get("eos_token_id"))
        eos = eos_all[0] if eos_all else None
        return ModelSpec(
            hidden=cfg["hidden_size"], n_layers=cfg["num_hidden_layers"],
            n_q_heads=cfg["num_attention_heads"], n_kv_heads=cfg["num_key_value_heads"],
            head_dim=cfg["head_dim"], intermediate=cfg["intermediate_size"], vocab=cfg["vocab_size"],
            rms_eps=cfg.get("rms_norm_eps", 1e-6),
            rope_theta=float(cfg.get("rope_theta") or (rp or {}).get("rope_theta") or 10000.0),
            rope_scaling=None, tie_embeddings=True,
            max_position_embeddings=cfg.get("max_position_embeddings", 8192),
            eos_token_id=eos, eos_token_ids=eos_all, bos_token_id=cfg.get("bos_token_id"), name=name, arch="gemma2",
            query_pre_attn_scalar=float(cfg.get("query_pre_attn_scalar", cfg["head_dim"])),
            attn_softcap=float(cfg.get("attn_logit_softcapping") or 0.0),
            final_softcap=float(cfg.get("final_logit_softcapping") or 0.0),
            sliding_window=int(cfg.get("sliding_window") or 0))

    def to_hf_config(self) -> dict:
        if self.arch == "gemma2":
            return {
                "architectures": ["Gemma2ForCausalLM"], "model_type": "gemma2",
                "hidden_size": self.hidden, "num_hidden_layers": self.n_layers,
                "num_attention_heads": self.n_q_heads, "num_key_value_heads": self.n_kv_heads,
                "head_dim": self.head_dim, "intermediate_size": self.intermediate,
                "vocab_size": self.vocab, "rms_norm_eps": self.rms_eps, "rope_theta": self.rope_theta,
                "tie_word_embeddings": True, "max_position_embeddings": self.max_position_embeddings,
                "hidden_activation": "gelu_pytorch_tanh", "attention_bias": False,
                "query_pre_attn_scalar": int(self.query_pre_attn_scalar or self.head_dim),  # HF: int field
                "attn_logit_softcapping": self.attn_softcap or None,
                "final_logit_softcapping": self.final_softcap or None,
                "sliding_window": self.sliding_window or None, "torch_dtype": "bfloat16",
                "bos_token_id": self.bos_token_id, "eos_token_id": self.eos_token_id,
            }
        return {
            "architectures": ["LlamaForCausalLM"], "model_type": "llama",
            "hidden_size": self.hidden, "num_hidden_layers": self.n_layers,
            "num_attention_heads": self.n_q_heads, "num_key_value_heads": self.n_kv_heads,
            "head_dim": self.head_dim, "intermediate_size": self.intermediate,
            "vocab_size": self.vocab, "rms_norm_eps": self.rms_eps, "rope_theta": self.rope_theta,
            "rope_scaling": self.rope_scaling, "tie_word_embeddings": self.tie_embeddings,
            "max_position_embeddings": self.max_position_embeddings, "hidden_act": "silu",
            "attention_bias": False, "mlp_bias": False, "torch_dtype": "bfloat16",
            "bos_token_id": self.bos_token_id, "eos_token_id": self.eos_token_id,
        }

    @property
    def qkv_dim(self) -> int:
        return (self.n_q_heads + 2 * self.n_kv_heads) * self.head_dim

    def weight_bytes_per_step(self) -> int:
        """bf16 bytes streamed per decode step: all layer matrices + LM head."""
        per_layer = (self.qkv_dim * self.hidden + self.hidden * self.n_q_heads * self.head_dim
                     + 3 * self.intermediate * self.hidden)
        return 2 * (self.n_layers * per_layer + self.vocab * self.hidden)

    def kv_bytes_per_token(self) -> int:
        return 2 * self.n_layers * self.n_kv_heads * self.head_dim * 2


# Appendix E of SURVEY.md
LLAMA_3_8B = ModelSpec(hidden=4096, n_layers=32, n_q_heads=32, n_kv_heads=8, head_dim=128,
                       intermediate=14336, vocab=128256, rms_eps=1e-5, rope_theta=500000.0,
                       tie_embeddings=False, max_position_embeddings=8192, eos_token_id=128001,
                       bos_token_id=128000, name="llama-3-8b")
LLAMA_32_1B = ModelSpec(hidden=2048, n_layers=16, n_q_heads=32, n_kv_heads=8, head_dim=64,
                        intermediate=8192, vocab=128256, rms_eps=1e-5, rope_theta=500000.0,
                        rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                                      "high_freq_factor": 4.0,
                                      "original_max_position_embeddings": 8192},
                        tie_embeddings=True, max_position_embeddings=131072, eos_token_id=128009,
                        bos_token_id=128000, name="llama-3.2-1b")
# google/gemma-2-9b(-it) and Unbabel/Tower-Plus-9B (BASELINE configs #4, #5) share this shape
GEMMA_2_9B = ModelSpec(hidden=3584, n_layers=42, n_q_heads=16, n_kv_heads=8, head_dim=256,
                       intermediate=14336, vocab=256000, rms_eps=1e-6, rope_theta=10000.0,
                       tie_embeddings=True, max_position_embeddings=8192, eos_token_id=1,
                       bos_token_id=2, name="gemma-2-9b", arch="gemma2", query_pre_attn_scalar=256.0,
                       attn_softcap=50.0, final_softcap=30.0, sliding_window=4096)
GEMMA_2_2B = ModelSpec(hidden=2304, n_layers=26, n_q_heads=8, n_kv_heads=4, head_dim=256,
                       intermediate=9216, vocab=256000, rms_eps=1e-6, rope_theta=10000.0,
                       tie_embeddings=True, max_position_embeddings=8192, eos_token_id=1,
                       bos_token_id=2, name="gemma-2-2b", arch="gemma2", query_pre_attn_scalar=256.0,
                       attn_softcap=50.0, final_softcap=30.0, sliding_window=4096)
BUILTIN_SPECS = {"llama-3-8b": LLAMA_3_8B, "llama-3.2-1b": LLAMA_32_1B,
                 "gemma-2-9b": GEMMA_2_9B, "gemma-2-2b": GEMMA_2_2B}


def rope_table(max_pos: int, head_dim: int, theta: float, scaling: Optional[dict]) -> torch.Tensor:
    """bf16 [max_pos, head_dim] = (cos | sin), computed in fp32 like HF's LlamaRotaryEmbedding
    (incl. the llama3 frequency scaling) and cast to the activation dtype."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    if scaling and scaling.get("rope_type", scaling.get("type")) == "llama3":
        factor = scaling["factor"]
        lo, hi = scaling["low_freq_factor"], scaling["high_freq_factor"]
        orig = scaling["original_max_position_embeddings"]
        wl = 2 * math.pi / inv
        inv_l = torch.where(wl > orig / lo, inv / factor, inv)
        smooth = (orig / wl - lo) / (hi - lo)
        smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
        mid = ~(wl < orig / hi) * ~(wl > orig / lo)
        inv = torch.where(mid, smoothed, inv_l)
    freqs = torch.outer(torch.arange(max_pos, dtype=torch.float32), inv)
    return torch.cat([freqs.cos(), freqs.sin()], dim=-1).to(torch.bfloat16)


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor, block: int = 128) -> torch.Tensor:
    """[I,H],[I,H] -> [2I,H] with rows [256j,256j+128) = gate[128j:128j+128], next 128 = the same
    rows of up, so one 256-wide GEMM tile carries both halves of 128 SwiGLU outputs and the
    activation can be applied in the GEMM epilogue (include/b200q.h: b200q_gemm_swiglu_bf16)."""
    I, H = gate.shape
    assert I % block == 0 and up.shape == gate.shape
    return torch.stack([gate.view(I // block, block, H), up.view(I // block, block, H)], 1).reshape(2 * I, H)


def fuse_hf_weights(spec: ModelSpec, sd: Dict[str, torch.Tensor]) -> Iterable[Tuple[str, torch.Tensor]]:
    """HF checkpoint names -> engine names, with q/k/v and gate/up concatenated row-wise
    (same fusion vLLM's QKVParallelLinear / MergedColumnParallelLinear perform)."""
    yield "embed", sd["model.embed_tokens.weight"]
    yield "final_norm", sd["model.norm.weight"].reshape(1, -1)
    if not spec.tie_embeddings:
        yield "lm_head", sd["lm_head.weight"]
    for i in range(spec.n_layers):
        p = f"model.layers.{i}."
        yield f"layers.{i}.input_norm", sd[p + "input_layernorm.weight"].reshape(1, -1)
        if spec.arch == "gemma2":
            yield f"layers.{i}.post_attn_norm", sd[p + "post_attention_layernorm.weight"].reshape(1, -1)
            yield f"layers.{i}.pre_ffn_norm", sd[p + "pre_feedforward_layernorm.weight"].reshape(1, -1)
            yield f"layers.{i}.post_ffn_norm", sd[p + "post_feedforward_layernorm.weight"].reshape(1, -1)
        else:
            yield f"layers.{i}.post_norm", sd[p + "post_attention_layernorm.weight"].reshape(1, -1)
        yield f"layers.{i}.qkv", torch.cat([sd[p + "self_attn.q_proj.weight"],
                                            sd[p + "self_attn.k_proj.weight"],
                                            sd[p + "self_attn.v_proj.weight"]], 0)
        yield f"layers.{i}.o", sd[p + "self_attn.o_proj.weight"]
        yield f"layers.{i}.gate_up", interleave_gate_up(sd[p + "mlp.gate_proj.weight"],
                                                        sd[p + "mlp.up_proj.weight"])
        yield f"layers.{i}.down", sd[p + "mlp.down_proj.weight"]


def random_engine_weights(spec: ModelSpec, seed: int, device, std: float = 0.02):
    """Random-init weights generated directly on the GPU in engine layout (benchmarks: there is
    no checkpoint on disk and 16 GB of safetensors should not travel).  Normal(0, std), unit norms."""
    g = torch.Generator(device=device).manual_seed(seed)

    def mat(r, c):
        return (torch.randn(r, c, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)

    gemma = spec.arch == "gemma2"
    # unit norm scale: w = 1 for Llama's x*w, w = 0 for Gemma's x*(1+w)
    ones = lambda: torch.full((1, spec.hidden), 0.0 if gemma else 1.0, dtype=torch.bfloat16, device=device)
    yield "embed", mat(spec.vocab, spec.hidden)
    yield "final_norm", ones()
    if not spec.tie_embeddings:
        yield "lm_head", mat(spec.vocab, spec.hidden)
    for i in range(spec.n_layers):
        yield f"layers.{i}.input_norm", ones()
        if gemma:
            for n in ("post_attn_norm", "pre_ffn_norm", "post_ffn_norm"):
                yield f"layers.{i}.{n}", ones()
        else:
            yield f"layers.{i}.post_norm", ones()
        yield f"layers.{i}.qkv", mat(spec.qkv_dim, spec.hidden)
        yield f"layers.{i}.o", mat(spec.hidden, spec.n_q_heads * spec.head_dim)
        yield f"layers.{i}.gate_up", mat(2 * spec.intermediate, spec.hidden)
        yield f"layers.{i}.down", mat(spec.hidden, spec.intermediate)


def load_hf_state_dict(model_dir: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file

    files = sorted(f for f in os.listdir(model_dir) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors files in {model_dir}")
    sd: Dict[str, torch.Tensor] = {}
    for f in files:
        sd.update(load_file(os.path.join(model_dir, f)))
    return sd


def resolve_model(model_name: str) -> Tuple[ModelSpec, Optional[str]]:
    """model_name is what the user passes to `llmq worker run <model> <queue>`: a local directory
    with config.json (+ safetensors, tokenizer), or `random:<builtin>` for random-init
    benchmarks (no network in this deployment, so hub ids must already be on disk)."""
    if model_name.startswith("random:"):
        key = model_name.split(":", 1)[1]
        if key not in BUILTIN_SPECS:
            raise ValueError(f"unknown builtin spec {key!r}; have {sorted(BUILTIN_SPECS)}")
        return BUILTIN_SPECS[key], None
    cfg_path = os.path.join(model_name, "config.json")
    if not os.path.exists(cfg_path):
        raise FileNotFoundError(f"{cfg_path} not found: pass a local model directory (offline deployment)")
    with open(cfg_path) as f:
        cfg = json.load(f)
    return ModelSpec.from_hf_config(cfg, name=os.path.basename(os.path.normpath(model_name))), model_name


class NativeModel:
    """Owns the device buffers (weights, KV pool, RoPE table, workspace) and the b200q_model
    handle they are bound to."""

    def __init__(self, spec: ModelSpec, weights: Iterable[Tuple[str, torch.Tensor]], *,
                 max_tokens: int = 4096, max_seqs: int = 256, max_model_len: int = 2048,
                 num_blocks: Optional[int] = None, gpu_memory_utilization: float = 0.9,
                 device: Optional[torch.device] = None):
        L.require_device()
        self.spec = spec
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.lib = L.load()
        self.max_model_len = min(max_model_len, spec.max_position_embeddings)
        if spec.arch != "gemma2" and spec.sliding_window:
            # Mistral-style window on every layer: not implemented as a window — the context is
            # capped to it instead, so the window never bites and results stay exact
            self.max_model_len = min(self.max_model_len, spec.sliding_window)
        self.cfg = L.ModelConfig(
            hidden=spec.hidden, n_layers=spec.n_layers, n_q_heads=spec.n_q_heads,
            n_kv_heads=spec.n_kv_heads, head_dim=spec.head_dim, intermediate=spec.intermediate,
            vocab=spec.vocab, block_size=BLOCK_SIZE, max_tokens=max_tokens, max_seqs=max_seqs,
            max_pos=self.max_model_len, tie_embeddings=int(spec.tie_embeddings),
            rms_eps=spec.rms_eps, attn_scale=spec.attn_scale,
            arch=L.ARCH_GEMMA2 if spec.arch == "gemma2" else L.ARCH_LLAMA,
            sliding_window=spec.sliding_window if spec.arch == "gemma2" else 0,
            attn_softcap=spec.attn_softcap,
            final_softcap=spec.final_softcap, embed_scale=spec.embed_scale)
        h = C.c_void_p()
        L.check(self.lib.b200q_model_create(C.byref(self.cfg), C.byref(h)))
        self.handle = h
        self._keep: List[torch.Tensor] = []
        for name, t in weights:
            t = t.to(device=self.device, dtype=torch.bfloat16).contiguous()
            self._keep.append(t)
            L.check(self.lib.b200q_model_bind_weight(h, name.encode(), t.data_ptr(), t.shape[0], t.shape[1]))
        self.rope = rope_table(self.max_model_len, spec.head_dim, spec.rope_theta,
                               spec.rope_scaling).to(self.device).contiguous()
        L.check(self.lib.b200q_model_bind_rope(h, self.rope.data_ptr()))
        ws_bytes = int(self.lib.b200q_model_workspace_bytes(C.byref(self.cfg)))
        if ws_bytes <= 0:
            raise L.B200QError(ws_bytes, "workspace sizing failed")
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        L.check(self.lib.b200q_model_bind_workspace(h, self.workspace.data_ptr(), ws_bytes))
        block_bytes = spec.n_layers * 2 * spec.n_kv_heads * BLOCK_SIZE * spec.head_dim * 2
        if num_blocks is None:
            num_blocks = self.pool_blocks_for(gpu_memory_utilization, block_bytes)
        if num_blocks < 1:
            raise MemoryError(
                f"KV pool of {num_blocks} blocks: gpu_memory_utilization={gpu_memory_utilization} leaves no "
                "room for a KV cache after weights and workspace")
        if num_blocks * BLOCK_SIZE < self.max_model_len:
            # vLLM refuses to start here; the worker serves what the pool can hold instead
            self.max_model_len = int(num_blocks) * BLOCK_SIZE
        self.num_blocks = int(num_blocks)
        self.kv = torch.zeros(spec.n_layers, self.num_blocks, 2, spec.n_kv_heads, BLOCK_SIZE,
                              spec.head_dim, dtype=torch.bfloat16, device=self.device)
        L.check(self.lib.b200q_model_bind_kv(h, self.kv.data_ptr(), self.num_blocks))
        torch.cuda.synchronize(self.device)

    def pool_blocks_for(self, gpu_memory_utilization: float, block_bytes: int) -> int:
        """KV blocks that fit `gpu_memory_utilization` of the device — vLLM's meaning of the knob
        (ref:llmq/workers/vllm_worker.py:107): the worker may use that fraction of the device; what
        is left after weights + workspace becomes the paged KV pool.  Memory this process has freed
        back to torch's caching allocator (a previous engine's pool, fp32 temporaries of the weight
        synthesis) is NOT in use: it is released to the driver first, otherwise cudaMemGetInfo still
        counts it and a second engine in the same process is sized to nothing."""
        torch.cuda.synchronize(self.device)
        torch.cuda.empty_cache()
        free, total = torch.cuda.mem_get_info(self.device)
        usable = int(total * gpu_memory_utilization) - (total - free)
        return max(usable // block_bytes, 0)

    def logits_view(self, n: int) -> torch.Tensor:
        """bf16 [n, vocab] view of the logits written by the last forward (tests only)."""
        ptr = self.lib.b200q_model_logits_ptr(self.handle)
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off: off + n * self.spec.vocab * 2].view(torch.bfloat16).view(n, self.spec.vocab)

    def set_profiling(self, on: bool) -> None:
        L.check(self.lib.b200q_model_set_profiling(self.handle, int(on)))

    def collect_profile(self, reset: bool = False) -> dict:
        """per-category device time (CUDA events on the forward's stream); the stream must be idle"""
        p = L.Profile()
        L.check(self.lib.b200q_model_profile_collect(self.handle, C.byref(p), int(reset)))
        return {n: {"ms": p.ms[i], "work": p.work[i], "launches": int(p.launches[i])}
                for i, n in enumerate(L.PROF_NAMES)}

    def close(self):
        """destroy the native handle and give the device memory back (weights, KV pool, workspace):
        the next engine built in this process must be able to size its pool from it"""
        if self.handle:
            torch.cuda.synchronize(self.device)
            self.lib.b200q_model_destroy(self.handle)
            self.handle = None
            self._keep.clear()
            self.kv = self.workspace = self.rope = None
            torch.cuda.empty_cache()


class Engine:
    """Python face of b200q_engine: add requests, run steps, get (req_id, token, flags) events."""

    def __init__(self, model: NativeModel, *, max_num_seqs: int, max_num_batched_tokens: int,
                 max_model_len: Optional[int] = None, eos_token_id=None,
                 policy: Optional[int] = None):
        """eos_token_id: None, one id, or an iterable of ids — every one of them ends a request
        (Llama-3.x-Instruct lists three, gemma-2-it adds <end_of_turn> in generation_config.json)"""
        import numpy as np

        self.model = model
        self.lib = model.lib
        self.np = np
        stop_ids = [] if eos_token_id is None else (
            [int(eos_token_id)] if isinstance(eos_token_id, int) else sorted({int(x) for x in eos_token_id}))
        self.max_model_len = min(max_model_len or model.max_model_len, model.max_model_len)
        ecfg = L.EngineConfig(
            max_num_seqs=max_num_seqs, max_num_batched_tokens=max_num_batched_tokens,
            max_model_len=self.max_model_len,
            eos_token_id=stop_ids[0] if stop_ids else -1,
            policy=int(os.environ.get("B200Q_SCHED_POLICY", "1")) if policy is None else int(policy))
        h = C.c_void_p()
        L.check(self.lib.b200q_engine_create(model.handle, C.byref(ecfg), C.byref(h)))
        self.handle = h
        if len(stop_ids) > 1:
            arr = (C.c_int32 * len(stop_ids))(*stop_ids)
            L.check(self.lib.b200q_engine_set_stop_ids(h, arr, len(stop_ids)))
        self.cap = max_num_seqs
        self._ids = np.zeros(self.cap, dtype=np.int64)
        self._tok = np.zeros(self.cap, dtype=np.int32)
        self._flg = np.zeros(self.cap, dtype=np.int32)

    def add_request(self, req_id: int, prompt_ids, max_new_tokens: int, ignore_eos: bool = False,
                    temperature: float = 0.0, seed: int = 0):
        """temperature 0 = greedy; > 0 = the reference's sampling (softmax(logits/T), exponential race)
        with a per-request Philox seed"""
        arr = self.np.ascontiguousarray(prompt_ids, dtype=self.np.int32)
        rc = self.lib.b200q_engine_add_request_sampled(
            self.handle, int(req_id), arr.ctypes.data, arr.size, int(max_new_tokens), int(ignore_eos),
            float(temperature), int(seed) & 0xFFFFFFFFFFFFFFFF)
        if rc == -1:  # B200Q_EINVAL: an un-servable job, the worker drops it (ValueError)
            raise ValueError(self.lib.b200q_last_error().decode())
        L.check(rc)

    def abort(self, req_id: int):
        L.check(self.lib.b200q_engine_abort(self.handle, int(req_id)))

    def set_async(self, on: bool):
        """async stepping (default on): step() enqueues step k+1 before reading step k's ids back and
        returns step k's events; off = one synchronisation per step (events of the step just run)"""
        L.check(self.lib.b200q_engine_set_async(self.handle, int(on)))

    def has_work(self) -> bool:
        return bool(self.lib.b200q_engine_has_work(self.handle))

    def step(self):
        """returns three numpy arrays (views, valid until the next step): ids, tokens, flags"""
        n = C.c_int32(0)
        L.check(self.lib.b200q_engine_step(self.handle, self._ids.ctypes.data, self._tok.ctypes.data,
                                           self._flg.ctypes.data, self.cap, C.byref(n)))
        k = n.value
        return self._ids[:k], self._tok[:k], self._flg[:k]

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.b200q_engine_stream(self.handle))

    def stats(self) -> L.EngineStats:
        s = L.EngineStats()
        L.check(self.lib.b200q_engine_get_stats(self.handle, C.byref(s)))
        return s

    def close(self):
        if self.handle:
            self.lib.b200q_engine_destroy(self.handle)
            self.handle = None
