"""ctypes binding of libb200q.so — the only way Python reaches the CUDA code.

The library is built in-tree (``llmq_b200/libb200q.so``) by ``__graft_entry__.build()`` /
``make -C llmq_b200/csrc``.  There is NO fallback: if the shared object is missing, or no
sm_100 device is present when a compute entry point is used, the call raises.

PyTorch tensors are used purely as device buffers: every wrapper passes ``tensor.data_ptr()``
and sizes across the C ABI declared in ``include/b200q.h``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200q.so")


class B200QError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200q error {code}: {msg}")
        self.code = code


class ModelConfig(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32),
        ("n_layers", C.c_int32),
        ("n_q_heads", C.c_int32),
        ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32),
        ("intermediate", C.c_int32),
        ("vocab", C.c_int32),
        ("block_size", C.c_int32),
        ("max_tokens", C.c_int32),
        ("max_seqs", C.c_int32),
        ("max_pos", C.c_int32),
        ("tie_embeddings", C.c_int32),
        ("rms_eps", C.c_float),
        ("attn_scale", C.c_float),
        # architecture extras (all zero = Llama)
        ("arch", C.c_int32),
        ("sliding_window", C.c_int32),
        ("attn_softcap", C.c_float),
        ("final_softcap", C.c_float),
        ("embed_scale", C.c_float),
    ]


ARCH_LLAMA = 0
ARCH_GEMMA2 = 1


class Batch(C.Structure):
    _fields_ = [
        ("T", C.c_int32),
        ("n_dec", C.c_int32),
        ("n_tiles", C.c_int32),
        ("n_sample", C.c_int32),
        ("bt_stride", C.c_int32),
        ("token_ids", C.c_void_p),
        ("positions", C.c_void_p),
        ("slot_mapping", C.c_void_p),
        ("block_table", C.c_void_p),
        ("ctx_lens", C.c_void_p),
        ("tiles", C.c_void_p),
        ("sample_rows", C.c_void_p),
        ("out_ids", C.c_void_p),
        ("sample_params", C.c_void_p),
        ("sum_ctx_dec", C.c_int64),
        ("prefill_flops_per_layer", C.c_int64),
        ("prev_out_ids", C.c_void_p),
    ]


class EngineConfig(C.Structure):
    _fields_ = [
        ("max_num_seqs", C.c_int32),
        ("max_num_batched_tokens", C.c_int32),
        ("max_model_len", C.c_int32),
        ("eos_token_id", C.c_int32),
        ("policy", C.c_int32),
    ]


class EngineStats(C.Structure):
    _fields_ = [
        ("steps", C.c_int64),
        ("tokens_prefilled", C.c_int64),
        ("tokens_decoded", C.c_int64),
        ("preemptions", C.c_int64),
        ("running", C.c_int32),
        ("waiting", C.c_int32),
        ("free_blocks", C.c_int32),
        ("total_blocks", C.c_int32),
        ("last_step_tokens", C.c_int32),
        ("last_step_seqs", C.c_int32),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
    ]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 4), ("work", C.c_double * 4), ("launches", C.c_int64 * 4)]


PROF_NAMES = ("gemm", "decode_attn", "prefill_attn", "elementwise")


FLAG_FINISHED_EOS = 1
FLAG_FINISHED_LENGTH = 2
FLAG_FINISHED_ABORT = 4

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); must list every symbol include/b200q.h declares
SIGNATURES = {
    "b200q_version": (_i, []),
    "b200q_last_error": (C.c_char_p, []),
    "b200q_device_check": (_i, []),
    "b200q_launch_count": (_i64, []),
    "b200q_embed": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "b200q_embed_ex": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200q_rmsnorm": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200q_add_rmsnorm": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200q_rope_kvwrite": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200q_decode_attn": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "b200q_prefill_attn": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "b200q_gemm_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200q_gemm_swiglu_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200q_gemm_bf16_splitk": (_i, [_vp, _vp, _i, _i, _i, _vp, C.POINTER(_vp), C.POINTER(_i)]),
    "b200q_add_rmsnorm_splitk": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "b200q_rope_kvwrite_splitk": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200q_swiglu": (_i, [_vp, _vp, _i, _i, _vp]),
    "b200q_gather_rows": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "b200q_argmax_bf16": (_i, [_vp, _vp, _i, _i, _vp]),
    "b200q_sample_bf16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "b200q_embed_scaled": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200q_gemma_rmsnorm": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200q_gemma_norm_add_norm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "b200q_decode_attn_ex": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "b200q_prefill_attn_ex": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "b200q_gemm_geglu_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "b200q_softcap_bf16": (_i, [_vp, _i64, _f, _vp]),
    "b200q_model_create": (_i, [C.POINTER(ModelConfig), C.POINTER(_vp)]),
    "b200q_model_destroy": (_i, [_vp]),
    "b200q_model_bind_weight": (_i, [_vp, C.c_char_p, _vp, _i64, _i64]),
    "b200q_model_bind_kv": (_i, [_vp, _vp, _i64]),
    "b200q_model_bind_rope": (_i, [_vp, _vp]),
    "b200q_model_workspace_bytes": (_i64, [C.POINTER(ModelConfig)]),
    "b200q_model_bind_workspace": (_i, [_vp, _vp, _i64]),
    "b200q_model_forward": (_i, [_vp, C.POINTER(Batch), _vp]),
    "b200q_model_logits_ptr": (_vp, [_vp]),
    "b200q_model_set_profiling": (_i, [_vp, _i]),
    "b200q_model_profile_collect": (_i, [_vp, C.POINTER(Profile), _i]),
    "b200q_engine_stream": (_vp, [_vp]),
    "b200q_engine_create": (_i, [_vp, C.POINTER(EngineConfig), C.POINTER(_vp)]),
    "b200q_engine_create_dryrun": (_i, [C.POINTER(EngineConfig), C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "b200q_engine_destroy": (_i, [_vp]),
    "b200q_engine_add_request": (_i, [_vp, _i64, _vp, C.c_int32, C.c_int32, C.c_int32]),
    "b200q_engine_add_request_sampled": (_i, [_vp, _i64, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64]),
    "b200q_engine_set_stop_ids": (_i, [_vp, _vp, C.c_int32]),
    "b200q_engine_abort": (_i, [_vp, _i64]),
    "b200q_engine_set_async": (_i, [_vp, C.c_int32]),
    "b200q_engine_has_work": (_i, [_vp]),
    "b200q_engine_step": (_i, [_vp, _vp, _vp, _vp, C.c_int32, C.POINTER(C.c_int32)]),
    "b200q_engine_get_stats": (_i, [_vp, C.POINTER(EngineStats)]),
}
# test / tuning hooks that are not part of the reference-facing header
EXTRA_SIGNATURES = {
    "b200q_gemm_set_tile_n": (_i, [_i]),
    "b200q_gemm_set_mode": (_i, [_i]),
    "b200q_gemm_resident_pairs": (_i, []),
    "b200q_gemm_set_debug": (_i, [_i]),
    "b200q_decode_attn_set_variant": (_i, [_i]),
    "b200q_gemm_set_splitk": (_i, [_i]),
    "b200q_stream_probe": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libb200q.so (once).  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C llmq_b200/csrc`. llmq_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in {**SIGNATURES, **EXTRA_SIGNATURES}.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200q_last_error()
        raise B200QError(rc, msg.decode("utf-8", "replace") if msg else "?")


def require_device() -> None:
    check(load().b200q_device_check())


def launch_count() -> int:
    return int(load().b200q_launch_count())


def _p(t) -> int:
    """data_ptr of a CUDA tensor (contiguity is the caller's contract)."""
    if t is None:
        return 0
    assert t.is_cuda and t.is_contiguous(), "libb200q wants contiguous CUDA tensors"
    return t.data_ptr()


def _stream(stream=None) -> int:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


# ---- thin op wrappers (used by tests/bench; the worker goes through Engine) ------------------
def embed(ids, table, out, stream=None):
    check(load().b200q_embed(_p(ids), _p(table), _p(out), ids.numel(), table.shape[1], _stream(stream)))


def embed_ex(ids, prev_out, table, out, scale=0.0, stream=None):
    """ids < 0 are read through prev_out[-1 - id] (the previous step's sampled ids)"""
    check(load().b200q_embed_ex(_p(ids), _p(prev_out), _p(table), _p(out), ids.numel(), table.shape[1], scale,
                                _stream(stream)))


def rmsnorm(x, w, y, eps, stream=None):
    check(load().b200q_rmsnorm(_p(x), _p(w), _p(y), x.shape[0], x.shape[1], eps, _stream(stream)))


def add_rmsnorm(x, residual, w, eps, stream=None):
    check(load().b200q_add_rmsnorm(_p(x), _p(residual), _p(w), x.shape[0], x.shape[1], eps, _stream(stream)))


def rope_kvwrite(qkv, cos_sin, positions, slot_mapping, kv_layer, n_q, n_kv, D, block_size, stream=None):
    check(load().b200q_rope_kvwrite(_p(qkv), _p(cos_sin), _p(positions), _p(slot_mapping), _p(kv_layer),
                                    qkv.shape[0], n_q, n_kv, D, block_size, _stream(stream)))


def decode_attn(qkv, out, kv_layer, block_table, ctx_lens, n_q, n_kv, D, block_size, scale, stream=None,
                softcap: float = 0.0, window: int = 0):
    if softcap or window:
        check(load().b200q_decode_attn_ex(_p(qkv), qkv.shape[1], _p(out), _p(kv_layer), _p(block_table),
                                          block_table.shape[1], _p(ctx_lens), ctx_lens.numel(), n_q, n_kv, D,
                                          block_size, scale, softcap, window, _stream(stream)))
        return
    check(load().b200q_decode_attn(_p(qkv), qkv.shape[1], _p(out), _p(kv_layer), _p(block_table),
                                   block_table.shape[1], _p(ctx_lens), ctx_lens.numel(), n_q, n_kv, D,
                                   block_size, scale, _stream(stream)))


def prefill_attn(qkv, out, kv_layer, block_table, tiles, n_q, n_kv, D, block_size, scale, stream=None,
                 softcap: float = 0.0, window: int = 0):
    if softcap or window:
        check(load().b200q_prefill_attn_ex(_p(qkv), qkv.shape[1], _p(out), _p(kv_layer), _p(block_table),
                                           block_table.shape[1], _p(tiles), tiles.shape[0], n_q, n_kv, D,
                                           block_size, scale, softcap, window, _stream(stream)))
        return
    check(load().b200q_prefill_attn(_p(qkv), qkv.shape[1], _p(out), _p(kv_layer), _p(block_table),
                                    block_table.shape[1], _p(tiles), tiles.shape[0], n_q, n_kv, D,
                                    block_size, scale, _stream(stream)))


def embed_scaled(ids, table, out, scale, stream=None):
    check(load().b200q_embed_scaled(_p(ids), _p(table), _p(out), ids.numel(), table.shape[1], scale, _stream(stream)))


def gemma_rmsnorm(x, w, y, eps, stream=None):
    check(load().b200q_gemma_rmsnorm(_p(x), _p(w), _p(y), x.shape[0], x.shape[1], eps, _stream(stream)))


def gemma_norm_add_norm(x, residual, w_post, w_next, eps, stream=None):
    """in place: residual <- bf16(residual + gemma_norm(x, w_post)); x <- gemma_norm(residual, w_next)"""
    check(load().b200q_gemma_norm_add_norm(_p(x), _p(residual), _p(w_post), _p(w_next), x.shape[0], x.shape[1],
                                           eps, _stream(stream)))


def gemm_geglu_bf16(a, w_interleaved, c, stream=None):
    M, K = a.shape
    N = w_interleaved.shape[0]
    assert w_interleaved.shape[1] == K and tuple(c.shape) == (M, N // 2)
    check(load().b200q_gemm_geglu_bf16(_p(a), _p(w_interleaved), _p(c), M, N, K, _stream(stream)))


def softcap_bf16(logits, cap, stream=None):
    check(load().b200q_softcap_bf16(_p(logits), logits.numel(), cap, _stream(stream)))


def gemm_bf16(a, w, c, stream=None):
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and tuple(c.shape) == (M, N)
    check(load().b200q_gemm_bf16(_p(a), _p(w), _p(c), M, N, K, _stream(stream)))


def gemm_bf16_splitk(a, w, stream=None):
    """decode-sized batches: split-K GEMM without its reduce pass.  Returns (partials_ptr, splits);
    splits == 1 means nothing was launched (the library would not split this shape)"""
    M, K = a.shape
    N = w.shape[0]
    part, splits = C.c_void_p(), C.c_int(1)
    check(load().b200q_gemm_bf16_splitk(_p(a), _p(w), M, N, K, _stream(stream), C.byref(part), C.byref(splits)))
    return part.value, splits.value


def add_rmsnorm_splitk(x, residual, w, partials_ptr, splits, eps, stream=None):
    check(load().b200q_add_rmsnorm_splitk(_p(x), _p(residual), _p(w), partials_ptr, splits, x.shape[0], x.shape[1], eps,
                                          _stream(stream)))


def rope_kvwrite_splitk(qkv, partials_ptr, splits, cos_sin, positions, slot_mapping, kv_layer, n_q, n_kv, D, block_size,
                        stream=None):
    check(load().b200q_rope_kvwrite_splitk(_p(qkv), partials_ptr, splits, _p(cos_sin), _p(positions), _p(slot_mapping),
                                           _p(kv_layer), qkv.shape[0], n_q, n_kv, D, block_size, _stream(stream)))


def gemm_swiglu_bf16(a, w_interleaved, c, stream=None):
    M, K = a.shape
    N = w_interleaved.shape[0]
    assert w_interleaved.shape[1] == K and tuple(c.shape) == (M, N // 2)
    check(load().b200q_gemm_swiglu_bf16(_p(a), _p(w_interleaved), _p(c), M, N, K, _stream(stream)))


def gemm_set_tile_n(bn: int):
    check(load().b200q_gemm_set_tile_n(bn))


def gemm_set_mode(mode: int):
    check(load().b200q_gemm_set_mode(mode))


def swiglu(gate_up, out, stream=None):
    check(load().b200q_swiglu(_p(gate_up), _p(out), gate_up.shape[0], out.shape[1], _stream(stream)))


def gather_rows(x, rows, out, stream=None):
    check(load().b200q_gather_rows(_p(x), _p(rows), _p(out), rows.numel(), x.shape[1], _stream(stream)))


def sample_bf16(logits, params, ids, stream=None):
    """params: int32 [B,4] = (temperature float bits, seed lo, seed hi, position)"""
    check(load().b200q_sample_bf16(_p(logits), _p(params), _p(ids), logits.shape[0], logits.shape[1], _stream(stream)))


def argmax_bf16(logits, ids, stream=None):
    check(load().b200q_argmax_bf16(_p(logits), _p(ids), logits.shape[0], logits.shape[1], _stream(stream)))
