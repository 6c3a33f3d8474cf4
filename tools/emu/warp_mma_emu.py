"""Lane-accurate numpy model of the warp-level pieces the attention kernels use (PTX ISA fragment
layouts): mma.sync.m16n8k16 (bf16 in, f32 accumulate), ldmatrix.x4 (.trans), movmatrix.m8n8.trans.
Used to check the per-lane index math of a kernel variant on the CPU before it meets a GPU.
Registers are modelled per lane: a "b32 register holding two bf16" is a float32 pair (lo, hi)."""
import numpy as np

LANES = np.arange(32)


def bf16(x):
    """round-to-nearest-even to bf16, kept as float32"""
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def ldmatrix_x4(smem_rows_of, trans):
    """smem_rows_of(lane) -> the 8 bf16 values (a 16-byte row) at the address lane supplies.
    Returns regs[i][lane] = (lo, hi) for the four 8x8 matrices i = 0..3."""
    rows = [np.asarray(smem_rows_of(l), dtype=np.float32) for l in range(32)]
    regs = np.zeros((4, 32, 2), dtype=np.float32)
    for i in range(4):
        M = np.stack([rows[8 * i + r] for r in range(8)])  # M[row][col], rows supplied by lanes 8i..8i+7
        if trans:
            M = M.T
        for l in range(32):
            regs[i, l] = M[l // 4, (l % 4) * 2: (l % 4) * 2 + 2]
    return regs


def mma_16816(c, a, b):
    """c: [32][4] f32 accumulators; a: [4][32][2] (a0..a3); b: [2][32][2] (b0, b1).  Returns new c."""
    A = np.zeros((16, 16), dtype=np.float32)
    B = np.zeros((16, 8), dtype=np.float32)
    C = np.zeros((16, 8), dtype=np.float32)
    for l in range(32):
        g, t = l // 4, (l % 4) * 2
        A[g, t: t + 2] = a[0][l]
        A[g + 8, t: t + 2] = a[1][l]
        A[g, t + 8: t + 10] = a[2][l]
        A[g + 8, t + 8: t + 10] = a[3][l]
        B[t: t + 2, g] = b[0][l]
        B[t + 8: t + 10, g] = b[1][l]
        C[g, t: t + 2] = c[l][0:2]
        C[g + 8, t: t + 2] = c[l][2:4]
    D = C + A.astype(np.float64) @ B.astype(np.float64)
    out = np.zeros((32, 4), dtype=np.float32)
    for l in range(32):
        g, t = l // 4, (l % 4) * 2
        out[l][0:2] = D[g, t: t + 2]
        out[l][2:4] = D[g + 8, t: t + 2]
    return out


def movmatrix_trans(reg):
    """reg: [32][2] = an 8x8 b16 matrix in fragment layout (row = lane/4, cols 2(lane%4),+1).
    Returns the transposed matrix in the same fragment layout."""
    M = np.zeros((8, 8), dtype=np.float32)
    for l in range(32):
        M[l // 4, (l % 4) * 2: (l % 4) * 2 + 2] = reg[l]
    M = M.T
    out = np.zeros((32, 2), dtype=np.float32)
    for l in range(32):
        out[l] = M[l // 4, (l % 4) * 2: (l % 4) * 2 + 2]
    return out


def shfl(vals, src_lane_of):
    """vals: [32]; src_lane_of: [32] source lane per lane"""
    return np.asarray(vals)[np.asarray(src_lane_of)]
