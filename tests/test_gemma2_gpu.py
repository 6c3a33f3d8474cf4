"""GPU parity of the Gemma-2 path (BASELINE configs #4 Tower-Plus-9B and #5
Gemma-2-9B-it are Gemma-2-shaped) through the C ABI: the new op-level entry points against
oracle/gemma2.py, the whole forward against the oracle AND against the committed transformers
Gemma2ForCausalLM goldens (tests/golden/hf_gemma2_tiny.npz), and the engine's greedy decoding
(sliding window crossed during decode, head_dim 256) against the oracle."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import ops as O
from oracle.gemma2 import (Gemma2Dims, Gemma2Oracle, attention_softcap, gelu_tanh, gemma_rms_norm,
                           random_gemma2_weights)
from tests.test_model_gpu import check_against_oracle, prompts, run_engine
from tests.test_ops_gpu import _make_cache, rnd
from tests.util import bf16_close

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- op level ------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,H", [(1, 256), (33, 2304), (130, 3584), (5, 8192)])
def test_gemma_rmsnorm(cuda, T, H):
    from llmq_b200 import lib
    x, w = rnd(T, H, seed=1, scale=3.0), rnd(H, seed=2, scale=0.3)
    y = torch.empty(T, H, dtype=BF, device=cuda)
    lib.gemma_rmsnorm(x.to(cuda), w.to(cuda), y, 1e-6)
    bf16_close(y, gemma_rms_norm(x.float(), w.float(), 1e-6, "bf16"), what="gemma_rmsnorm")


@pytest.mark.parametrize("T,H", [(1, 256), (64, 2304), (257, 3584)])
def test_gemma_norm_add_norm(cuda, T, H):
    from llmq_b200 import lib
    x, r = rnd(T, H, seed=3, scale=0.7), rnd(T, H, seed=4, scale=2.0)
    w1, w2 = rnd(H, seed=5, scale=0.3), rnd(H, seed=6, scale=0.3)
    xg, rg = x.to(cuda), r.to(cuda)
    lib.gemma_norm_add_norm(xg, rg, w1.to(cuda), w2.to(cuda), 1e-6)
    a = gemma_rms_norm(x.float(), w1.float(), 1e-6, "bf16")
    r_ref = (r.float() + a).to(BF).float()
    y_ref = gemma_rms_norm(r_ref, w2.float(), 1e-6, "bf16")
    # the first norm can flip the rounding of a handful of elements (reduction order: ~1e-5 of them);
    # such a flip is one bf16 ulp of `a` (|a| up to ~4), which enters the residual and the second norm
    # 1:1 — hence the absolute term; max_mismatch_frac keeps the check tight (<= 0.1 % differ at all)
    ulp_a = 2.0 ** -7 * a.abs().max().item()
    bf16_close(rg, r_ref, atol=ulp_a, max_mismatch_frac=0.001, what="norm_add_norm residual")
    bf16_close(xg, y_ref, ulps=2.0, atol=2 * ulp_a, what="norm_add_norm output")


def test_embed_scaled_and_softcap(cuda):
    from llmq_b200 import lib
    V, H, T = 512, 256, 37
    table = rnd(V, H, seed=7)
    ids = torch.randint(0, V, (T,), generator=torch.Generator().manual_seed(8), dtype=torch.int32)
    scale = float(torch.tensor(math.sqrt(H)).to(BF))
    out = torch.empty(T, H, dtype=BF, device=cuda)
    lib.embed_scaled(ids.to(cuda), table.to(cuda), out, scale)
    assert torch.equal(out.cpu().float(), (table[ids.long()].float() * scale).to(BF).float())
    lg = rnd(9, 1024, seed=9, scale=25.0)
    lg[0, :8] = torch.tensor([0.0, -0.0, 1e-3, -1e-3, 200.0, -200.0, 30.0, -30.0]).to(BF)
    dev = lg.to(cuda)
    lib.softcap_bf16(dev, 30.0)
    x = lg.float()
    ref = ((x / 30.0).to(BF).float().tanh().to(BF).float() * 30.0).to(BF).float()
    bf16_close(dev, ref, ulps=1.0, max_mismatch_frac=0.002, what="final logit softcap")


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,I,K", [(5, 128, 64), (130, 512, 256), (300, 9216, 2304)])
def test_gemm_geglu_fused(cuda, M, I, K, mode):
    """gate_up GEMM with GeGLU (tanh approximation) in the epilogue == GEMM -> bf16 -> oracle GeGLU"""
    from llmq_b200 import lib
    from llmq_b200.model import interleave_gate_up
    a, g, u = rnd(M, K, seed=30), rnd(I, K, seed=31, scale=0.05), rnd(I, K, seed=32, scale=0.05)
    out = torch.full((M, I), float("nan"), dtype=BF, device=cuda)
    lib.gemm_set_mode(mode)
    try:
        lib.gemm_geglu_bf16(a.to(cuda), interleave_gate_up(g, u).to(cuda), out)
        torch.cuda.synchronize()
    finally:
        lib.gemm_set_mode(0)
    ad = a.to(cuda).float()
    gg = (ad @ g.to(cuda).float().t()).to(BF).float().cpu()
    uu = (ad @ u.to(cuda).float().t()).to(BF).float().cpu()
    ref = (gelu_tanh(gg).to(BF).float() * uu).to(BF).float()
    bf16_close(out, ref, ulps=8.0, atol=2e-5 * math.sqrt(K), max_mismatch_frac=0.02, what=f"gemm_geglu mode={mode}")


ATTN_CASES = [  # D, n_q, n_kv, softcap, window
    (64, 8, 2, 50.0, 0), (128, 16, 8, 50.0, 40), (256, 16, 8, 50.0, 0), (256, 8, 4, 50.0, 100),
    (256, 4, 1, 0.0, 0), (128, 32, 8, 0.0, 33), (64, 32, 8, 20.0, 16),
]


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("D,n_q,n_kv,cap,window", ATTN_CASES)
def test_decode_attn_softcap_window(cuda, D, n_q, n_kv, cap, window, variant):
    from llmq_b200 import lib
    BS = 16
    ctxs = [1, 15, 16, 17, 33, 100, 129, 192, 255, 700, 64, 48, 40, 41, 600]
    kv, bt, ks, vs = _make_cache(ctxs, n_kv, D, BS, seed=41)
    B = len(ctxs)
    qkv = rnd(B, (n_q + 2 * n_kv) * D, seed=42, scale=2.0)
    out = torch.full((B, n_q * D), float("nan"), dtype=BF, device=cuda)
    scale = 6.0 / math.sqrt(D)  # score std ~ 12, tails beyond the cap: the soft-capping visibly bends them
    Lh = lib.load()
    lib.check(Lh.b200q_decode_attn_set_variant(variant))
    try:
        lib.decode_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda),
                        torch.tensor(ctxs, dtype=torch.int32, device=cuda), n_q, n_kv, D, BS, scale,
                        softcap=cap, window=window)
        torch.cuda.synchronize()
    finally:
        lib.check(Lh.b200q_decode_attn_set_variant(0))
    got = out.float().cpu().view(B, n_q, D)
    for i, c in enumerate(ctxs):
        q = qkv[i].float().view(n_q + 2 * n_kv, D)[:n_q][None]
        ref = attention_softcap(q, ks[i].float(), vs[i].float(), torch.tensor([c - 1]), scale,
                                cap or None, window or None, "bf16")[0]
        bf16_close(got[i], ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5,
                   what=f"decode v{variant} D={D} cap={cap} window={window} ctx={c}")


@pytest.mark.parametrize("D,n_q,n_kv,cap,window", ATTN_CASES)
def test_prefill_attn_softcap_window(cuda, D, n_q, n_kv, cap, window):
    from llmq_b200 import lib
    BS = 16
    chunks = [(0, 128), (0, 5), (16, 16), (100, 37), (3, 1 + 16 * 3), (250, 70)]
    ctxs = [a + n for a, n in chunks]
    kv, bt, ks, vs = _make_cache(ctxs, n_kv, D, BS, seed=43)
    T = sum(n for _, n in chunks)
    qkv = rnd(T, (n_q + 2 * n_kv) * D, seed=44, scale=2.0)
    tiles, row = [], 0
    for i, (a, n) in enumerate(chunks):
        for j in range(0, n, 16):
            tiles.append([i, row + j, min(16, n - j), a + j])
        row += n
    tiles_t = torch.tensor(tiles, dtype=torch.int32)
    out = torch.full((T, n_q * D), float("nan"), dtype=BF, device=cuda)
    scale = 6.0 / math.sqrt(D)  # score std ~ 12, tails beyond the cap: the soft-capping visibly bends them
    lib.prefill_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda), tiles_t.to(cuda), n_q, n_kv, D, BS, scale,
                     softcap=cap, window=window)
    torch.cuda.synchronize()
    got = out.float().cpu().view(T, n_q, D)
    row = 0
    for i, (a, n) in enumerate(chunks):
        q = qkv[row: row + n].float().view(n, n_q + 2 * n_kv, D)[:, :n_q]
        ref = attention_softcap(q, ks[i].float(), vs[i].float(), torch.arange(a, a + n), scale,
                                cap or None, window or None, "bf16")
        bf16_close(got[row: row + n], ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5,
                   what=f"prefill D={D} cap={cap} window={window} chunk {i} (ctx {a}+{n})")
        row += n


# ---- model level ---------------------------------------------------------------------------------
def golden_dims():
    meta = json.load(open(os.path.join(G, "hf_gemma2_meta.json")))
    return Gemma2Dims(**meta["dims"]), meta["weights_seed"]


GEMMA_TINY = {
    "hf_golden_d64": None,  # filled from tests/golden/hf_gemma2_meta.json
    "d256": Gemma2Dims(hidden=512, n_layers=3, n_q_heads=4, n_kv_heads=2, head_dim=256, intermediate=1024,
                       vocab=1024, query_pre_attn_scalar=256.0, sliding_window=32, max_pos=512),
    "d128_g4": Gemma2Dims(hidden=256, n_layers=2, n_q_heads=8, n_kv_heads=2, head_dim=128, intermediate=512,
                          vocab=512, query_pre_attn_scalar=100.0, sliding_window=24, max_pos=512),
}


def build_gemma(name, **kw):
    from llmq_b200.model import ModelSpec, NativeModel, fuse_hf_weights
    if name == "hf_golden_d64":
        d, seed = golden_dims()
    else:
        d, seed = GEMMA_TINY[name], 5
    w = random_gemma2_weights(d, seed=seed)
    spec = ModelSpec(hidden=d.hidden, n_layers=d.n_layers, n_q_heads=d.n_q_heads, n_kv_heads=d.n_kv_heads,
                     head_dim=d.head_dim, intermediate=d.intermediate, vocab=d.vocab, rms_eps=d.rms_eps,
                     rope_theta=d.rope_theta, tie_embeddings=True, max_position_embeddings=d.max_pos,
                     arch="gemma2", query_pre_attn_scalar=d.query_pre_attn_scalar,
                     attn_softcap=d.attn_softcap or 0.0, final_softcap=d.final_softcap or 0.0,
                     sliding_window=d.sliding_window)
    kw.setdefault("max_tokens", 512)
    kw.setdefault("max_seqs", 64)
    kw.setdefault("max_model_len", d.max_pos)
    kw.setdefault("num_blocks", 256)
    model = NativeModel(spec, fuse_hf_weights(spec, w), **kw)
    return model, Gemma2Oracle(d, w, "bf16"), d


def forward_all_rows(model, ids, cuda):
    """one prefill of len(ids) tokens with every row sampled -> bf16 logits [n, V] and argmax ids"""
    from llmq_b200 import lib as L
    n = len(ids)
    BS = 16
    nb = (n + BS - 1) // BS
    blocks = [7, 3, 11, 5][:nb]
    meta = {
        "tok": torch.tensor(ids, dtype=torch.int32), "pos": torch.arange(n, dtype=torch.int32),
        "slot": torch.tensor([blocks[p // BS] * BS + p % BS for p in range(n)], dtype=torch.int32),
        "bt": torch.tensor([blocks + [0] * (8 - nb)], dtype=torch.int32),
        "ctx": torch.zeros(1, dtype=torch.int32),
        "tiles": torch.tensor([[0, j, min(16, n - j), j] for j in range(0, n, 16)], dtype=torch.int32),
        "rows": torch.arange(n, dtype=torch.int32),
    }
    dev = {k: v.to(cuda).contiguous() for k, v in meta.items()}
    out = torch.zeros(n, dtype=torch.int32, device=cuda)
    b = L.Batch(T=n, n_dec=0, n_tiles=dev["tiles"].shape[0], n_sample=n, bt_stride=8,
                token_ids=dev["tok"].data_ptr(), positions=dev["pos"].data_ptr(),
                slot_mapping=dev["slot"].data_ptr(), block_table=dev["bt"].data_ptr(),
                ctx_lens=dev["ctx"].data_ptr(), tiles=dev["tiles"].data_ptr(),
                sample_rows=dev["rows"].data_ptr(), out_ids=out.data_ptr())
    L.check(model.lib.b200q_model_forward(model.handle, C.byref(b), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return model.logits_view(n).float().cpu(), out.cpu().numpy()


def test_gemma2_forward_matches_hf_golden(cuda):
    """the native forward against logits computed by transformers' Gemma2ForCausalLM (fp32 and bf16
    runs, committed under tests/golden/) on the same seeded weights: 48 tokens > the 16-token
    sliding window, so both layer types are exercised"""
    model, oracle, d = build_gemma("hf_golden_d64")
    z = np.load(os.path.join(G, "hf_gemma2_tiny.npz"))
    got, picked = forward_all_rows(model, z["ids"].tolist(), cuda)
    ref32 = torch.tensor(z["logits_fp32"])
    ref16 = torch.tensor(z["logits_bf16"].astype(np.float32))
    ours_vs_fp32 = (got - ref32).abs().max().item()
    hf_bf16_vs_fp32 = (ref16 - ref32).abs().max().item()
    # stated tolerance: our bf16 path is at most 1.5x as far from the fp32 truth as HF's own bf16 run
    # (plus 2 bf16 ulps of the largest logit)
    assert ours_vs_fp32 <= 1.5 * hf_bf16_vs_fp32 + 2 * 2.0 ** -8 * ref32.abs().max().item(), (ours_vs_fp32, hf_bf16_vs_fp32)
    top2 = ref32.topk(2, -1).values
    safe = ((top2[:, 0] - top2[:, 1]) > 4 * hf_bf16_vs_fp32).numpy()
    assert safe.sum() >= 10
    assert np.array_equal(picked[safe], O.argmax_first(ref32)[safe])
    model.close()


@pytest.mark.parametrize("name", list(GEMMA_TINY))
def test_gemma2_forward_logits_vs_oracle(cuda, name):
    model, oracle, d = build_gemma(name)
    n = 61
    ids = prompts(d.vocab, [n], seed=3)[0]
    got, picked = forward_all_rows(model, ids, cuda)
    ref, _ = oracle.forward(torch.tensor(ids), torch.arange(n))
    diff = (got - ref).abs()
    # stated tolerance: logits (soft-capped to +-30, here |l| < ~3) within 0.1 absolute, mean < 0.015
    assert diff.max().item() < 0.1 and diff.mean().item() < 0.015, (diff.max().item(), diff.mean().item())
    top2 = ref.topk(2, -1).values
    safe = ((top2[:, 0] - top2[:, 1]) > 0.2).numpy()
    assert np.array_equal(picked[safe], O.argmax_first(ref)[safe])
    assert np.array_equal(picked, O.argmax_first(got)), "argmax kernel vs its own (soft-capped) logits"
    model.close()


@pytest.mark.parametrize("name", list(GEMMA_TINY))
def test_gemma2_engine_greedy_matches_oracle(cuda, name):
    """continuous batching over ragged prompts; prompts and generations cross the sliding window,
    so windowed decode and windowed chunked prefill are both on the path"""
    model, oracle, d = build_gemma(name)
    reqs = prompts(d.vocab, [1, 5, 16, 17, 40, 130, 64, 33, 2, 100], seed=6)
    outs, st = run_engine(model, reqs, max_new=12)
    exact = check_against_oracle(oracle, reqs, outs, 12, margin_tol=0.16)
    assert exact >= len(reqs) - 3, f"only {exact}/{len(reqs)} requests matched the oracle exactly"
    chunked, _ = run_engine(model, reqs, max_new=12, max_num_batched_tokens=24, max_num_seqs=8, policy=0)
    check_against_oracle(oracle, reqs, chunked, 12, margin_tol=0.16)
    model.close()
