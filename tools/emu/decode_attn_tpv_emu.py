"""CPU emulation of one warp of the decode-attention kernel with the TRANSPOSED PV formulation
("worked-out designs", item 1), lane by lane with the fragment layouts of
warp_mma_emu.py, against a plain fp32 softmax-attention reference.  Mirrors the CUDA code
statement by statement (qk_page as in attn.cu; pv_page_t, alpha/l hand-over by shuffles,
movmatrix epilogue are the new parts)."""
import math
import sys

import numpy as np

from warp_mma_emu import bf16, ldmatrix_x4, mma_16816, movmatrix_trans, shfl

BS = 16


def swizzled_page(page):  # page [n_tokens<=16][D] logical -> stored [16][D] with chunk c at c ^ (t & 7)
    D = page.shape[1]
    out = np.full((BS, D), np.nan, dtype=np.float32)  # stale rows poisoned
    for t in range(page.shape[0]):
        for c in range(D // 8):
            out[t, ((c ^ (t & 7)) * 8): ((c ^ (t & 7)) * 8) + 8] = page[t, c * 8: c * 8 + 8]
    return out


def run_item(q, K, V, scale, softcap=0.0, window=0):
    """q [G][D]; K, V [ctx][D]  (bf16-representable).  Returns out [G][D] as the warp would write it."""
    G, D = q.shape
    ctx = K.shape[0]
    lanes = np.arange(32)
    r, cq = lanes // 4, (lanes % 4) * 2
    log2e = 1.4426950408889634
    if softcap > 0:
        c0, c1 = scale / softcap, softcap * log2e
    else:
        c0, c1 = scale * log2e, 0.0
    # Q fragments: qa[ks][0] = q[r][ks*16 + cq, +1], qa[ks][2] = q[r][ks*16 + 8 + cq, +1] (rows >= G are zero)
    qa0 = np.zeros((D // 16, 32, 2), np.float32)
    qa2 = np.zeros((D // 16, 32, 2), np.float32)
    for l in range(32):
        if r[l] < G:
            for ks in range(D // 16):
                qa0[ks, l] = q[r[l], ks * 16 + cq[l]: ks * 16 + cq[l] + 2]
                qa2[ks, l] = q[r[l], ks * 16 + 8 + cq[l]: ks * 16 + 8 + cq[l] + 2]
    zero = np.zeros((32, 2), np.float32)
    o = np.zeros((D // 16, 32, 4), np.float32)  # O^T tiles: [dim = r (+8)][head = cq, cq+1]
    m = np.full(32, -np.inf, np.float32)
    lsum = np.zeros(32, np.float32)
    lo = ctx - window if (window > 0 and ctx > window) else 0
    n_pages = (ctx + BS - 1) // BS
    for p in range(lo // BS, n_pages):
        n_valid = min(BS, ctx - p * BS)
        t_lo = lo - p * BS
        ks_page = swizzled_page(K[p * BS: p * BS + n_valid])
        vs_page = swizzled_page(V[p * BS: p * BS + n_valid])
        vs_page[n_valid:] = 0.0  # zero_tail_rows
        # ---- qk_page<false>: S[nt] (16 x 8 tokens), rows = heads ----
        s = np.zeros((2, 32, 4), np.float32)
        for nt in range(2):
            for j in range(D // 32):
                def rows_of(l, nt=nt, j=j):
                    token = nt * 8 + (l & 7)
                    chunk = 4 * j + (l >> 3)
                    pc = chunk ^ (token & 7)
                    return np.nan_to_num(ks_page[token, pc * 8: pc * 8 + 8], nan=1e30)  # stale K rows: garbage
                b = ldmatrix_x4(rows_of, trans=False)
                s[nt] = mma_16816(s[nt], [qa0[2 * j], zero, qa2[2 * j], zero], [b[0], b[1]])
                s[nt] = mma_16816(s[nt], [qa0[2 * j + 1], zero, qa2[2 * j + 1], zero], [b[2], b[3]])
        # ---- mask / scale / online softmax (rows = heads: lane holds head r, tokens cq, cq+1) ----
        mx = np.full(32, -np.inf, np.float32)
        for nt in range(2):
            for l in range(32):
                for e in range(2):
                    t0 = nt * 8 + cq[l] + e
                    ok = t0 < n_valid and t0 >= t_lo
                    v = s[nt, l, e]
                    v = (math.tanh(v * c0) * c1 if softcap > 0 else v * c0) if ok else -np.inf
                    s[nt, l, e] = v
                mx[l] = max(mx[l], s[nt, l, 0], s[nt, l, 1])
        mx = np.maximum(mx, shfl(mx, lanes ^ 1))
        mx = np.maximum(mx, shfl(mx, lanes ^ 2))
        m_new = np.maximum(m, mx)
        alpha = np.exp2(m - m_new)
        psum = np.zeros(32, np.float32)
        for nt in range(2):
            s[nt, :, 0] = np.exp2(s[nt, :, 0] - m_new)
            s[nt, :, 1] = np.exp2(s[nt, :, 1] - m_new)
            psum += s[nt, :, 0] + s[nt, :, 1]
        lsum = lsum * alpha + psum
        m = m_new
        # ---- NEW: rescale O^T by the per-head alpha of the lane's two columns (heads cq, cq+1) ----
        a_h0 = shfl(alpha, cq * 4)        # lane 4*h holds head h
        a_h1 = shfl(alpha, (cq + 1) * 4)
        o[:, :, 0] *= a_h0
        o[:, :, 1] *= a_h1
        o[:, :, 2] *= a_h0
        o[:, :, 3] *= a_h1
        # ---- NEW: pv_page_t.  B = P^T fragments are the S accumulators themselves ----
        pb0 = bf16(s[0, :, 0:2])  # tokens 0-7:  B[k = cq, cq+1][n = r]
        pb1 = bf16(s[1, :, 0:2])  # tokens 8-15
        for tile in range(D // 16):
            def rows_of(l, tile=tile):
                i = l >> 3                      # matrix index
                token = (i >> 1) * 8 + (l & 7)  # a0,a1: tokens 0-7; a2,a3: tokens 8-15
                chunk = 2 * tile + (i & 1)      # a0,a2: dims 0-7 of the tile; a1,a3: dims 8-15
                pc = chunk ^ (token & 7)
                return vs_page[token, pc * 8: pc * 8 + 8]
            a = ldmatrix_x4(rows_of, trans=True)
            o[tile] = mma_16816(o[tile], [a[0], a[1], a[2], a[3]], [pb0, pb1])
    # ---- epilogue: 1/l per head, movmatrix back to [head][dim], 4-byte stores ----
    lsum = lsum + shfl(lsum, lanes ^ 1)
    lsum = lsum + shfl(lsum, lanes ^ 2)
    inv = np.where(lsum > 0, 1.0 / np.maximum(lsum, 1e-38), 0.0).astype(np.float32)
    i_h0, i_h1 = shfl(inv, cq * 4), shfl(inv, (cq + 1) * 4)
    out = np.zeros((8, D), np.float32)
    for tile in range(D // 16):
        lo_half = bf16(np.stack([o[tile, :, 0] * i_h0, o[tile, :, 1] * i_h1], 1))   # dims tile*16 + r
        hi_half = bf16(np.stack([o[tile, :, 2] * i_h0, o[tile, :, 3] * i_h1], 1))   # dims tile*16 + 8 + r
        t_lo_half, t_hi_half = movmatrix_trans(lo_half), movmatrix_trans(hi_half)
        for l in range(32):  # lane now holds head r, dims tile*16 (+8) + cq, cq+1
            if r[l] < G:
                out[r[l], tile * 16 + cq[l]: tile * 16 + cq[l] + 2] = t_lo_half[l]
                out[r[l], tile * 16 + 8 + cq[l]: tile * 16 + 8 + cq[l] + 2] = t_hi_half[l]
    return out[:G]


def reference(q, K, V, scale, softcap=0.0, window=0):
    ctx = K.shape[0]
    s = (q.astype(np.float64) @ K.astype(np.float64).T) * scale
    if softcap > 0:
        s = softcap * np.tanh(s / softcap)
    j = np.arange(ctx)
    mask = j > ctx - 1
    if window:
        mask = mask | (j <= ctx - 1 - window)
    s[:, mask] = -np.inf
    e = np.exp(s - s.max(-1, keepdims=True))
    return (bf16(e.astype(np.float32)).astype(np.float64) @ V.astype(np.float64)) / e.sum(-1, keepdims=True)


def main():
    rng = np.random.default_rng(0)
    worst = 0.0
    for D in (64, 128, 256):
        for G in (1, 2, 4, 8):
            for ctx, cap, win in ((1, 0, 0), (15, 0, 0), (16, 50.0, 0), (17, 0, 0), (40, 50.0, 24), (100, 0, 33), (77, 20.0, 0)):
                q = bf16(rng.standard_normal((G, D)) * 2)
                K = bf16(rng.standard_normal((ctx, D)))
                V = bf16(rng.standard_normal((ctx, D)))
                scale = 6.0 / math.sqrt(D)
                got = run_item(q, K, V, scale, cap, win)
                ref = reference(q, K, V, scale, cap, win)
                err = np.abs(got - ref).max()
                worst = max(worst, err)
                assert err < 3e-2, (D, G, ctx, cap, win, err)
    print("transposed-PV decode attention emulation matches the reference; worst abs err %.4f" % worst)


if __name__ == "__main__":
    sys.exit(main())
