"""GPU parity of every op-level C-ABI entry point against the CPU oracle (oracle/ops.py).
Run under gpurun:  python -m pytest tests -m gpu -x -q"""
import math

import numpy as np
import pytest
import torch

from oracle import ops as O
from tests.util import bf16_close

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF)


@pytest.mark.parametrize("T,H", [(1, 256), (7, 2048), (33, 4096), (300, 4096), (5, 8192)])
def test_rmsnorm(cuda, T, H):
    from llmq_b200 import lib
    x, w = rnd(T, H, seed=1, scale=3.0), rnd(H, seed=2)
    y = torch.empty(T, H, dtype=BF, device=cuda)
    lib.rmsnorm(x.to(cuda), w.to(cuda), y, 1e-5)
    bf16_close(y, O.rms_norm(x.float(), w.float(), 1e-5), what="rmsnorm")


@pytest.mark.parametrize("T,H", [(1, 256), (64, 2048), (257, 4096)])
def test_add_rmsnorm(cuda, T, H):
    from llmq_b200 import lib
    x, r, w = rnd(T, H, seed=3), rnd(T, H, seed=4, scale=2.0), rnd(H, seed=5)
    xg, rg = x.to(cuda), r.to(cuda)
    lib.add_rmsnorm(xg, rg, w.to(cuda), 1e-5)
    y_ref, r_ref = O.add_rms_norm(x.float(), r.float(), w.float(), 1e-5)
    assert torch.equal(rg.float().cpu(), r_ref), "residual sum must be bit-exact"
    bf16_close(xg, y_ref, what="add_rmsnorm")


def test_embed_gather_argmax(cuda):
    from llmq_b200 import lib
    V, H, T = 1000, 512, 77
    table = rnd(V, H, seed=6)
    ids = torch.randint(0, V, (T,), dtype=torch.int32)
    out = torch.empty(T, H, dtype=BF, device=cuda)
    lib.embed(ids.to(cuda), table.to(cuda), out)
    assert torch.equal(out.cpu(), table[ids.long()])
    rows = torch.tensor([5, 0, 76, 5], dtype=torch.int32)
    sel = torch.empty(4, H, dtype=BF, device=cuda)
    lib.gather_rows(out, rows.to(cuda), sel)
    assert torch.equal(sel.cpu(), table[ids.long()][rows.long()])
    # argmax incl. exact ties (lowest index wins) and a vocab that is not a multiple of 8
    for Vv in (128256, 1003):
        logits = rnd(9, Vv, seed=7)
        logits[0, 17] = logits[0, 900] = 50.0
        logits[1, Vv - 1] = 60.0
        logits[2, 0] = 60.0
        logits[3] = 1.0
        got = torch.empty(9, dtype=torch.int32, device=cuda)
        lib.argmax_bf16(logits.to(cuda), got)
        assert np.array_equal(got.cpu().numpy(), O.argmax_first(logits.float())), Vv


@pytest.mark.parametrize("T,I", [(3, 512), (130, 14336)])
def test_swiglu(cuda, T, I):
    from llmq_b200 import lib
    gu = rnd(T, 2 * I, seed=8, scale=2.0)
    out = torch.empty(T, I, dtype=BF, device=cuda)
    lib.swiglu(gu.to(cuda), out)
    bf16_close(out, O.swiglu(gu.float()), what="swiglu", max_mismatch_frac=0.03)


@pytest.mark.parametrize("D,n_q,n_kv", [(128, 32, 8), (64, 32, 8), (128, 4, 1), (256, 16, 8), (256, 8, 4)])
def test_rope_kvwrite(cuda, D, n_q, n_kv):
    from llmq_b200 import lib
    T, BS, NB, max_pos = 37, 16, 12, 512
    qkv = rnd(T, (n_q + 2 * n_kv) * D, seed=9)
    table = O.rope_table(max_pos, D, 500000.0, None)
    pos = torch.randint(0, max_pos, (T,), dtype=torch.int32)
    slots = torch.randperm(NB * BS)[:T].to(torch.int32)
    slots[3] = -1  # skipped write
    kv = torch.zeros(NB, 2, n_kv, BS, D, dtype=BF, device=cuda)
    qg = qkv.to(cuda)
    lib.rope_kvwrite(qg, table.to(BF).to(cuda), pos.to(cuda), slots.to(cuda), kv, n_q, n_kv, D, BS)
    x = qkv.float().view(T, n_q + 2 * n_kv, D)
    q_ref = O.rope_neox(x[:, :n_q], pos, table)
    k_ref = O.rope_neox(x[:, n_q:n_q + n_kv], pos, table)
    v_ref = x[:, n_q + n_kv:]
    got = qg.float().cpu().view(T, n_q + 2 * n_kv, D)
    assert torch.equal(got[:, :n_q], q_ref), "rotated q must be bit-exact"
    kvc = kv.float().cpu()
    expect = torch.zeros_like(kvc)
    for t in range(T):
        s = int(slots[t])
        if s < 0:
            continue
        for h in range(n_kv):
            O.kv_page_write(expect, s // BS, 0, h, s % BS, k_ref[t, h])
            O.kv_page_write(expect, s // BS, 1, h, s % BS, v_ref[t, h])
    assert torch.equal(kvc, expect), "paged KV contents (swizzled layout) must be bit-exact"


# Attention tolerance: P is rounded to bf16 before PV in both the kernels and the oracle, but the
# kernels round exp(s - running_max) page by page while the oracle rounds exp(s - global_max); each
# term carries 2^-9 relative error, so outputs (|o| ~ 0.1..1 for unit-variance V) agree to ~5e-3
# absolute / 2 bf16 ulps, whichever is larger.
def _make_cache(seqs_ctx, n_kv, D, BS, seed):
    """random logical K/V per sequence -> swizzled paged cache + block table"""
    g = torch.Generator().manual_seed(seed)
    n_blocks_needed = sum((c + BS - 1) // BS for c in seqs_ctx)
    NB = n_blocks_needed + 5
    perm = torch.randperm(NB, generator=g).tolist()
    kv = torch.zeros(NB, 2, n_kv, BS, D, dtype=BF)
    # poison unused space with NaN bit patterns: the kernels must never let them through
    kv.view(torch.int16)[:] = 0x7FC0
    max_blocks = max((c + BS - 1) // BS for c in seqs_ctx)
    bt = torch.zeros(len(seqs_ctx), (max_blocks + 7) // 8 * 8, dtype=torch.int32)
    ks, vs = [], []
    sw = O.kv_swizzle_index(BS, D)
    for i, c in enumerate(seqs_ctx):
        k = (torch.randn(c, n_kv, D, generator=g)).to(BF)
        v = (torch.randn(c, n_kv, D, generator=g)).to(BF)
        ks.append(k), vs.append(v)
        for b in range((c + BS - 1) // BS):
            blk = perm.pop()
            bt[i, b] = blk
            n = min(BS, c - b * BS)
            for h in range(n_kv):
                for which, src in ((0, k), (1, v)):
                    page = src[b * BS: b * BS + n, h]  # [n, D]
                    row = kv[blk, which, h]
                    row[:n] = torch.zeros(n, D, dtype=BF).scatter(1, sw[:n], page)
    return kv, bt, ks, vs


@pytest.mark.parametrize("D,n_q,n_kv", [(128, 32, 8), (64, 32, 8), (128, 8, 1)])
def test_decode_attn(cuda, D, n_q, n_kv):
    from llmq_b200 import lib
    BS = 16
    ctxs = [1, 15, 16, 17, 100, 129, 192, 255, 700, 64, 33]
    kv, bt, ks, vs = _make_cache(ctxs, n_kv, D, BS, seed=11)
    B = len(ctxs)
    qkv = rnd(B, (n_q + 2 * n_kv) * D, seed=12)
    out = torch.empty(B, n_q * D, dtype=BF, device=cuda)
    scale = 1.0 / math.sqrt(D)
    lib.decode_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda),
                    torch.tensor(ctxs, dtype=torch.int32, device=cuda), n_q, n_kv, D, BS, scale)
    got = out.float().cpu().view(B, n_q, D)
    for i, c in enumerate(ctxs):
        q = qkv[i].float().view(n_q + 2 * n_kv, D)[:n_q][None]
        ref = O.attention(q, ks[i].float(), vs[i].float(), torch.tensor([c - 1]), scale)[0]
        bf16_close(got[i], ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5, what=f"decode ctx={c}")


@pytest.mark.parametrize("D,n_q,n_kv", [(128, 32, 8), (64, 32, 8)])
def test_prefill_attn(cuda, D, n_q, n_kv):
    from llmq_b200 import lib
    BS = 16
    # (context already in the cache before this chunk, chunk length)
    chunks = [(0, 128), (0, 5), (16, 16), (100, 37), (3, 1 + 16 * 3)]
    ctxs = [a + n for a, n in chunks]
    kv, bt, ks, vs = _make_cache(ctxs, n_kv, D, BS, seed=13)
    T = sum(n for _, n in chunks)
    qkv = rnd(T, (n_q + 2 * n_kv) * D, seed=14)
    tiles, row = [], 0
    for i, (a, n) in enumerate(chunks):
        for j in range(0, n, 16):
            tiles.append([i, row + j, min(16, n - j), a + j])
        row += n
    tiles_t = torch.tensor(tiles, dtype=torch.int32)
    out = torch.zeros(T, n_q * D, dtype=BF, device=cuda)
    scale = 1.0 / math.sqrt(D)
    lib.prefill_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda), tiles_t.to(cuda), n_q, n_kv, D, BS, scale)
    got = out.float().cpu().view(T, n_q, D)
    row = 0
    for i, (a, n) in enumerate(chunks):
        q = qkv[row: row + n].float().view(n, n_q + 2 * n_kv, D)[:, :n_q]
        ref = O.attention(q, ks[i].float(), vs[i].float(), torch.arange(a, a + n), scale)
        bf16_close(got[row: row + n], ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5,
                   what=f"prefill chunk {i} (ctx {a}+{n})")
        row += n


GEMM_SHAPES = [(128, 256, 64), (1, 64, 64), (5, 128, 128), (130, 512, 1024), (256, 6144, 4096),
               (77, 4096, 14336), (300, 1024, 4096), (16, 128256, 2048)]


@pytest.mark.parametrize("bn", [0, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm(cuda, M, N, K, bn):
    from llmq_b200 import lib
    if bn and N % bn:
        pytest.skip("tile does not divide N")
    a, w = rnd(M, K, seed=20), rnd(N, K, seed=21, scale=0.05)
    c = torch.full((M, N), float("nan"), dtype=BF, device=cuda)
    lib.gemm_set_tile_n(bn)
    try:
        lib.gemm_bf16(a.to(cuda), w.to(cuda), c)
        torch.cuda.synchronize()
    finally:
        lib.gemm_set_tile_n(0)
    ref = (a.to(cuda).float() @ w.to(cuda).float().t()).to(BF)  # fp32-accumulate torch reference
    # atol: fp32 accumulation-order noise, ~1e-6 x sqrt(K) x |a||w| (matters only next to zero)
    atol = 2e-5 * math.sqrt(K)
    bf16_close(c, ref.cpu(), ulps=1.0, atol=atol, max_mismatch_frac=0.02, what=f"gemm {M}x{N}x{K} bn={bn}")
    if M * N * K <= 130 * 512 * 1024:
        bf16_close(c, O.linear(a.float(), w.float()), ulps=1.0, atol=atol, what="gemm vs oracle")


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("M,N,K", [(129, 256, 64), (256, 512, 1024), (1000, 6144, 4096), (700, 4096, 14336)])
def test_gemm_2cta(cuda, M, N, K, bn):
    """cta_group::2 kernel (CTA pair, UMMA M=256, B tile split across the pair)"""
    from llmq_b200 import lib
    a, w = rnd(M, K, seed=22), rnd(N, K, seed=23, scale=0.05)
    c = torch.full((M, N), float("nan"), dtype=BF, device=cuda)
    lib.gemm_set_mode(2)
    lib.gemm_set_tile_n(bn)
    try:
        lib.gemm_bf16(a.to(cuda), w.to(cuda), c)
        torch.cuda.synchronize()
    finally:
        lib.gemm_set_mode(0)
        lib.gemm_set_tile_n(0)
    ref = (a.to(cuda).float() @ w.to(cuda).float().t()).to(BF)
    bf16_close(c, ref.cpu(), ulps=1.0, atol=2e-5 * math.sqrt(K), max_mismatch_frac=0.02, what=f"gemm2 {M}x{N}x{K} bn={bn}")


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,I,K", [(5, 128, 64), (130, 512, 256), (300, 14336, 4096)])
def test_gemm_swiglu_fused(cuda, M, I, K, mode):
    """gate_up GEMM with the SwiGLU applied in the epilogue == unfused GEMM -> bf16 -> oracle swiglu"""
    from llmq_b200 import lib
    from llmq_b200.model import interleave_gate_up
    a, g, u = rnd(M, K, seed=30), rnd(I, K, seed=31, scale=0.05), rnd(I, K, seed=32, scale=0.05)
    w = interleave_gate_up(g, u)
    out = torch.full((M, I), float("nan"), dtype=BF, device=cuda)
    lib.gemm_set_mode(mode)
    try:
        lib.gemm_swiglu_bf16(a.to(cuda), w.to(cuda), out)
        torch.cuda.synchronize()
    finally:
        lib.gemm_set_mode(0)
    ad = a.to(cuda).float()
    gu = torch.cat([(ad @ g.to(cuda).float().t()).to(BF), (ad @ u.to(cuda).float().t()).to(BF)], 1).cpu()
    ref = O.swiglu(gu.float())
    # a 1-ulp flip of g or u (accumulation order) moves the product by about one ulp as well
    bf16_close(out, ref, ulps=8.0, atol=2e-5 * math.sqrt(K), max_mismatch_frac=0.02, what=f"gemm_swiglu mode={mode}")


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("D,n_q,n_kv", [(128, 32, 8), (64, 32, 8), (128, 8, 1)])
def test_decode_attn_streaming_variant(cuda, D, n_q, n_kv, variant):
    """persistent warp-per-(sequence, kv-head) kernel; variant 3 forces many items per warp so the
    page ring runs across item borders (ragged contexts, partial pages, >32-page items)"""
    from llmq_b200 import lib
    BS = 16
    g = np.random.default_rng(4)
    ctxs = g.integers(1, 90, size=40).tolist() + [1, 16, 17, 255, 129, 600, 513, 32, 48, 1]
    kv, bt, ks, vs = _make_cache(ctxs, n_kv, D, BS, seed=17)
    B = len(ctxs)
    qkv = rnd(B, (n_q + 2 * n_kv) * D, seed=18)
    out = torch.full((B, n_q * D), float("nan"), dtype=BF, device=cuda)
    scale = 1.0 / math.sqrt(D)
    L = lib.load()
    lib.check(L.b200q_decode_attn_set_variant(variant))
    try:
        lib.decode_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda),
                        torch.tensor(ctxs, dtype=torch.int32, device=cuda), n_q, n_kv, D, BS, scale)
        torch.cuda.synchronize()
    finally:
        lib.check(L.b200q_decode_attn_set_variant(0))
    got = out.float().cpu().view(B, n_q, D)
    for i, c in enumerate(ctxs):
        q = qkv[i].float().view(n_q + 2 * n_kv, D)[:n_q][None]
        ref = O.attention(q, ks[i].float(), vs[i].float(), torch.tensor([c - 1]), scale)[0]
        bf16_close(got[i], ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5, what=f"decode-stream v{variant} ctx={c}")


@pytest.mark.parametrize("splits", [0, 2, 4, 7, 8])
@pytest.mark.parametrize("M,N,K", [(1, 4096, 4096), (64, 4096, 14336), (130, 6144, 4096), (256, 4096, 14336)])
def test_gemm_splitk(cuda, M, N, K, splits):
    """split-K for decode-sized M: fp32 partial tiles + fixed-order reduce == the unsplit result up to
    fp32 summation order, and bit-identical from run to run"""
    from llmq_b200 import lib
    if splits and (K // 64) % splits:
        pytest.skip("split does not divide the k-blocks")
    a, w = rnd(M, K, seed=40), rnd(N, K, seed=41, scale=0.05)
    ad, wd = a.to(cuda), w.to(cuda)
    c1 = torch.full((M, N), float("nan"), dtype=BF, device=cuda)
    c2 = torch.full((M, N), float("nan"), dtype=BF, device=cuda)
    L = lib.load()
    lib.check(L.b200q_gemm_set_splitk(splits))
    try:
        lib.gemm_bf16(ad, wd, c1)
        lib.gemm_bf16(ad, wd, c2)
        torch.cuda.synchronize()
    finally:
        lib.check(L.b200q_gemm_set_splitk(0))
    assert torch.equal(c1, c2), "split-K must be deterministic"
    ref = (ad.float() @ wd.float().t()).to(BF)
    bf16_close(c1, ref.cpu(), ulps=1.0, atol=2e-5 * math.sqrt(K), max_mismatch_frac=0.02, what=f"splitk={splits} {M}x{N}x{K}")


@pytest.mark.parametrize("variant", [1, 3])
def test_decode_attn_long_context(cuda, variant):
    """contexts beyond 32 pages per warp (block-id register window reloads) up to 5000 tokens"""
    from llmq_b200 import lib
    D, n_q, n_kv, BS = 128, 8, 2, 16
    ctxs = [5000, 2049, 513, 4097, 33]
    kv, bt, ks, vs = _make_cache(ctxs, n_kv, D, BS, seed=51)
    B = len(ctxs)
    qkv = rnd(B, (n_q + 2 * n_kv) * D, seed=52)
    out = torch.full((B, n_q * D), float("nan"), dtype=BF, device=cuda)
    scale = 1.0 / math.sqrt(D)
    L = lib.load()
    lib.check(L.b200q_decode_attn_set_variant(variant))
    try:
        lib.decode_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda),
                        torch.tensor(ctxs, dtype=torch.int32, device=cuda), n_q, n_kv, D, BS, scale)
        torch.cuda.synchronize()
    finally:
        lib.check(L.b200q_decode_attn_set_variant(0))
    got = out.float().cpu().view(B, n_q, D)
    for i, c in enumerate(ctxs):
        q = qkv[i].float().view(n_q + 2 * n_kv, D)[:n_q][None]
        ref = O.attention(q, ks[i].float(), vs[i].float(), torch.tensor([c - 1]), scale)[0]
        bf16_close(got[i], ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5, what=f"decode long ctx={c} v{variant}")


def test_prefill_attn_long_context_chunk(cuda):
    """a 40-token chunk appended to a 3000-token context (chunked prefill deep into a sequence)"""
    from llmq_b200 import lib
    D, n_q, n_kv, BS = 128, 8, 2, 16
    a, n = 3000, 40
    kv, bt, ks, vs = _make_cache([a + n], n_kv, D, BS, seed=53)
    qkv = rnd(n, (n_q + 2 * n_kv) * D, seed=54)
    tiles = torch.tensor([[0, j, min(16, n - j), a + j] for j in range(0, n, 16)], dtype=torch.int32)
    out = torch.zeros(n, n_q * D, dtype=BF, device=cuda)
    scale = 1.0 / math.sqrt(D)
    lib.prefill_attn(qkv.to(cuda), out, kv.to(cuda), bt.to(cuda), tiles.to(cuda), n_q, n_kv, D, BS, scale)
    q = qkv.float().view(n, n_q + 2 * n_kv, D)[:, :n_q]
    ref = O.attention(q, ks[0].float(), vs[0].float(), torch.arange(a, a + n), scale)
    bf16_close(out.float().cpu().view(n, n_q, D), ref, ulps=2.0, atol=5e-3, max_mismatch_frac=0.5, what="prefill long ctx")


@pytest.mark.parametrize("M", [1, 16, 120, 256])
def test_splitk_partials_fused_into_norm_and_rope_are_bit_identical(cuda, M):
    """decode-sized batches: b200q_gemm_bf16_splitk leaves the fp32 partials of a split-K projection
    in the library's scratch and the consumer (add+RMSNorm / RoPE+KV write) reduces them on the fly.
    The fused chain must be BIT-identical to GEMM (with its own reduce pass) followed by the plain op:
    same summation order over the splits, same rounding of the sum to bf16."""
    from llmq_b200 import lib
    H, n_q, n_kv, D, BS = 4096, 32, 8, 128, 16
    QKV = (n_q + 2 * n_kv) * D
    a = rnd(M, H, seed=21).to(cuda)
    # o-projection shape [H, H] -> add + RMSNorm
    w_o = rnd(H, H, seed=22, scale=0.02).to(cuda)
    res0, wn = rnd(M, H, seed=23, scale=2.0).to(cuda), rnd(H, seed=24).to(cuda)
    c = torch.empty(M, H, dtype=BF, device=cuda)
    lib.gemm_bf16(a, w_o, c)
    r_ref = res0.clone()
    lib.add_rmsnorm(c, r_ref, wn, 1e-5)   # c <- normed, r_ref <- residual + gemm
    part, splits = lib.gemm_bf16_splitk(a, w_o)
    assert splits > 1, "a [<=256, 4096] x [4096, 4096] projection is expected to split K"
    x, r = torch.empty(M, H, dtype=BF, device=cuda), res0.clone()
    lib.add_rmsnorm_splitk(x, r, wn, part, splits, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(r, r_ref) and torch.equal(x, c)
    # a shape the library does not split: nothing is launched and the caller is told so
    assert lib.gemm_bf16_splitk(rnd(512, 256, seed=26).to(cuda), rnd(256, 256, seed=27).to(cuda))[1] == 1
    # qkv projection [QKV, H] -> RoPE + paged KV write
    w_qkv = rnd(QKV, H, seed=25, scale=0.02).to(cuda)
    table = O.rope_table(512, D, 500000.0, None).to(BF).to(cuda)
    pos = torch.randint(0, 512, (M,), dtype=torch.int32).to(cuda)
    slots = torch.randperm(40 * BS)[:M].to(torch.int32)
    if M > 3:
        slots[3] = -1
    slots = slots.to(cuda)
    qkv_ref = torch.empty(M, QKV, dtype=BF, device=cuda)
    lib.gemm_bf16(a, w_qkv, qkv_ref)
    kv_ref = torch.zeros(40, 2, n_kv, BS, D, dtype=BF, device=cuda)
    lib.rope_kvwrite(qkv_ref, table, pos, slots, kv_ref, n_q, n_kv, D, BS)
    part, splits = lib.gemm_bf16_splitk(a, w_qkv)
    if M > 128:  # two m-tiles x 48 n-tiles already occupy most SMs: the library does not split this one
        assert splits == 1
        return
    assert splits > 1
    qkv = torch.zeros(M, QKV, dtype=BF, device=cuda)
    kv = torch.zeros(40, 2, n_kv, BS, D, dtype=BF, device=cuda)
    lib.rope_kvwrite_splitk(qkv, part, splits, table, pos, slots, kv, n_q, n_kv, D, BS)
    torch.cuda.synchronize()
    assert torch.equal(qkv[:, :n_q * D], qkv_ref[:, :n_q * D]), "rotated q"
    assert torch.equal(kv, kv_ref), "paged K/V"



def test_model_forward_is_identical_with_and_without_splitk_fusion(cuda, monkeypatch):
    """whole-engine check of the same: greedy + sampled tokens with B200Q_FUSE_SPLITK on / off"""
    from llmq_b200.model import BUILTIN_SPECS, Engine
    from llmq_b200.service import build_service
    import dataclasses
    outs = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("B200Q_FUSE_SPLITK", fuse)
        svc = build_service("random:llama-3.2-1b", max_num_seqs=64, max_model_len=256, gpu_memory_utilization=0.9,
                            max_num_batched_tokens=512, seed=3, num_blocks=1024)
        eng = svc.engine
        g = np.random.default_rng(2)
        for i in range(24):
            eng.add_request(i, g.integers(0, 128000, size=int(g.integers(3, 90))).tolist(), 12, ignore_eos=True,
                            temperature=0.7 if i % 4 == 0 else 0.0, seed=i)
        res = {i: [] for i in range(24)}
        while eng.has_work():
            ids, toks, _ = eng.step()
            for i, t in zip(ids.tolist(), toks.tolist()):
                res[i].append(t)
        outs[fuse] = res
        eng.close()
        svc.engine.model.close()
    assert outs["1"] == outs["0"]
