// gemm_tcgen05.cu — C[M,N] = A[M,K] . W[N,K]^T (bf16 in, fp32 accumulate in TMEM, bf16 out)
// for the QKV / O / gate_up / down / LM-head projections (SURVEY.md §2.1 K3,K8,K9,K11,K12),
// the only dense contractions of the hot path.  Hand-written for sm_100a:
//
//   * persistent, warp-specialised CTA (192 threads): warp 0 = TMA producer (UTMALDG, 128B
//     swizzle), warp 1 = tcgen05.mma issuer (one lane; also owns the TMEM allocation),
//     warps 2-5 = epilogue (tcgen05.ld -> bf16 -> global).
//   * tile 128 x BN x 64 (BN in {64,128,256}), smem ring of 4..8 stages (<= 192 KB), operands
//     K-major in shared memory exactly as TMA lays them down (SWIZZLE_128B), described to the
//     tensor core by UMMA shared-memory descriptors; UMMA shape M=128, N=BN, K=16.
//   * two TMEM accumulator stages (2 x BN fp32 columns): the MMA warp runs tile i+1 while the
//     epilogue warps drain tile i.
//   * tile order is m-fastest so concurrently running CTAs share one weight tile through L2
//     (weights stream from HBM exactly once; activations are L2-resident).
//
// Roofline: decode-sized M is HBM-bound on the weight stream (N*K*2 bytes), M >= ~256 is
// tensor-bound (2*M*N*K flops against the measured bf16 peak).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace b200q {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BUDGET = 196608;  // ring bytes

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = GEMM_SMEM_BUDGET / STAGE_BYTES > 8 ? 8 : GEMM_SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: powers of two >= 32
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4 * 4096 /*epilogue*/;
};

// ---- tcgen05 PTX wrappers ------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)
//  layout_type[61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);        // start address, 16 B units
  d |= (uint64_t)1 << 16;                              // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

// instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(GEMM_BM >> 4) << 24);
}

// Drain one warp's 32 x BN slice of the fp32 accumulator tile (TMEM lanes 32q..32q+31, one row per
// thread) to global memory.  Each thread converts 64 columns of its row to bf16 (128 B), the warp
// transposes them through a 4 KB XOR-swizzled shared-memory patch and stores them back as
// 4 rows x 128 contiguous bytes per instruction — full 128 B lines instead of 32 scattered 16 B
// pieces, which matters because the L2 request rate is shared with the TMA operand stream.
// kSwiGLU (K9+K10 fused): the weight rows were interleaved at load time so that the tile's
// columns are [gate (BN/2) | up (BN/2)] of the SAME BN/2 output columns; the epilogue applies
// out = bf16(bf16(silu(bf16(g))) * bf16(u)) — identical rounding points to the unfused path
// (GEMM output rounded to bf16, then oracle/ops.py::swiglu) — and writes BN/2 columns, so the
// [T, 2I] intermediate never touches HBM.
constexpr int EPI_SMEM_PER_WARP = 32 * 128;  // 32 rows x 64 bf16

// kAct: 0 = plain GEMM, 1 = SwiGLU (Llama), 2 = GeGLU with the tanh approximation (Gemma-2):
// out = bf16(bf16(gelu_tanh(bf16(g))) * bf16(u)), gelu_tanh evaluated in fp32 like torch's kernel:
// 0.5 * g * (1 + tanh(sqrt(2/pi) * (g + 0.044715 g^3)))
constexpr int ACT_NONE = 0, ACT_SWIGLU = 1, ACT_GEGLU = 2;

template <int BN, int kAct>
__device__ __forceinline__ void epilogue_rows(uint32_t taddr, bf16* ctile, long long ldc,
                                              int row0, int M, uint8_t* patch, int lane) {
  constexpr bool kSwiGLU = kAct != ACT_NONE;  // gated epilogue: tile = [gate (BN/2) | up (BN/2)]
  constexpr int OUT_BN = kSwiGLU ? BN / 2 : BN;
  auto act = [](uint32_t gb, uint32_t ub) -> float {
    const float g = round_bf16(__uint_as_float(gb)), u = round_bf16(__uint_as_float(ub));
    if constexpr (kAct == ACT_GEGLU) {
      const float inner = 0.7978845608028654f * (g + 0.044715f * (g * g * g));
      return round_bf16(0.5f * g * (1.f + tanhf(inner))) * u;
    } else {
      return round_bf16(g / (1.f + __expf(-g))) * u;
    }
  };
  const uint32_t patch_s = smem_u32(patch);
#pragma unroll 1
  for (int c0 = 0; c0 < OUT_BN; c0 += 64) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // two 32-column TMEM loads -> 64 output columns of this row
      uint32_t v[32];
      uint4 o[4];
      if constexpr (!kSwiGLU) {
        tmem_ld32(taddr + (uint32_t)(c0 + 32 * h), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j].x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
          o[j].y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
          o[j].z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
          o[j].w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
        }
      } else {
        uint32_t u[32];
        tmem_ld32(taddr + (uint32_t)(c0 + 32 * h), v);
        tmem_ld32(taddr + (uint32_t)(BN / 2 + c0 + 32 * h), u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j].x = pack_bf16x2(act(v[8 * j + 0], u[8 * j + 0]), act(v[8 * j + 1], u[8 * j + 1]));
          o[j].y = pack_bf16x2(act(v[8 * j + 2], u[8 * j + 2]), act(v[8 * j + 3], u[8 * j + 3]));
          o[j].z = pack_bf16x2(act(v[8 * j + 4], u[8 * j + 4]), act(v[8 * j + 5], u[8 * j + 5]));
          o[j].w = pack_bf16x2(act(v[8 * j + 6], u[8 * j + 6]), act(v[8 * j + 7], u[8 * j + 7]));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // chunk (4h + j) of this thread's 128 B row
        const uint32_t a = patch_s + lane * 128 + (((4 * h + j) ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o[j].x), "r"(o[j].y),
                     "r"(o[j].z), "r"(o[j].w)
                     : "memory");
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // 4 rows x 128 B per instruction
      const int rr = i * 4 + (lane >> 3), c = lane & 7;
      uint4 o;
      const uint32_t a = patch_s + rr * 128 + ((c ^ (rr & 7)) << 4);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                   : "r"(a));
      if (row0 + rr < M) st_v4(ctile + (long long)rr * ldc + c0 + c * 8, o);
    }
    __syncwarp();
  }
}

template <int BN, int kAct = ACT_NONE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_b, bf16* __restrict__ C, int M,
                     int N, int K, int dbg, int splits, float* __restrict__ ws, int a_bytes) {
  constexpr bool kSwiGLU = kAct != ACT_NONE;
  // a_bytes: bytes one A k-block really carries.  For M < 128 the tensor map's box holds only the
  // rows that exist (rounded up to the 8-row swizzle atom), not 128: the other rows of the UMMA's A
  // tile are stale shared memory whose products land in output rows >= M, which are never stored.
  // A decode-sized GEMM re-reads the whole activation matrix in EVERY CTA — with 128 x 128 tiles as
  // many L2->SM bytes as the weights themselves (ncu: profiles/r2_ncu_small_m.csv) — so rows that do
  // not exist should not be paid for.
  // splits > 1 (split-K for decode-sized M, where a projection has too few output tiles to put
  // every SM on the weight stream): tile space = m x split x n, each CTA accumulates K/splits of
  // the reduction and writes an fp32 partial tile to ws[split][M][N]; splitk_reduce_kernel sums the
  // partials in a fixed order (deterministic) and rounds to bf16.
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + 256;  // 4 warps x 4 KB transpose patches

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = N / BN;
  const int num_tiles = m_tiles * n_tiles * splits;
  const int k_blocks = (K / GEMM_BK) / splits;  // k-blocks per tile (per split)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full + 0, 1);
    mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty + 0, 4);
    mbar_init(tmem_empty + 1, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % m_tiles, rest = tile / m_tiles;
        const int n_blk = rest / splits, kb0 = (rest % splits) * k_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty + stage, phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (dbg & 1) {  // timing experiment: MMA on stale smem, no operand traffic
            mbar_arrive(full + stage);
          } else {
            mbar_expect_tx(full + stage, (uint32_t)(a_bytes + Cfg::B_BYTES));
            tma_load_2d(sa, &tmap_a, (kb0 + kb) * GEMM_BK, m_blk * GEMM_BM, full + stage);
            tma_load_2d(sb, &tmap_b, (kb0 + kb) * GEMM_BK, n_blk * BN, full + stage);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(tmem_empty + acc, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * 32);
            const uint64_t bd = make_smem_desc(sb + k * 32);
            umma_bf16(tmem_d, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + stage);  // smem stage free once these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tmem_full + acc);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % m_tiles, rest = tile / m_tiles;
      const int n_blk = rest / splits, split = rest % splits;
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const int row0 = m_blk * GEMM_BM + quarter * 32;
      constexpr int OUT_BN = kSwiGLU ? BN / 2 : BN;
      const long long ldc = kSwiGLU ? N / 2 : N;
      bf16* ctile = C + (long long)row0 * ldc + (long long)n_blk * OUT_BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      if (splits > 1) {
        // fp32 partial of this K-slice: ws[split][row][col]; one full 128 B line per thread per load
        const int row = row0 + lane;
        float* wrow = ws + ((long long)split * M + row) * N + (long long)n_blk * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (row < M) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_v4(wrow + c0 + 4 * j, make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
        }
      } else if (!(dbg & 2)) {  // dbg&2: skip the drain (timing experiment)
        epilogue_rows<BN, kAct>(taddr, ctile, ldc, row0, M, epi_smem + (warp - 2) * EPI_SMEM_PER_WARP, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// out[m][n] = bf16(sum_s ws[s][m][n]), partials added in split order (deterministic)
__global__ void __launch_bounds__(256)
    splitk_reduce_kernel(const float4* __restrict__ ws, uint4* __restrict__ out, long long mn8,
                         int splits) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < mn8;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = ws[2 * i], b = ws[2 * i + 1];
    for (int s = 1; s < splits; ++s) {
      const float4 c = ws[2 * (i + (long long)s * mn8)], d = ws[2 * (i + (long long)s * mn8) + 1];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    out[i] = o;
  }
}

// ==================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x BN tile.  Each CTA stages
// its own 128 rows of A and HALF of the B tile (BN/2 weight rows), the leader issues
// tcgen05.mma.cta_group::2 (UMMA M=256) which reads both CTAs' shared memory, and each CTA drains
// its own 128 x BN accumulator from its TMEM.  Per-SM operand ingest per k-block drops from
// 16 KB + BN*128 B to 16 KB + BN*64 B for the same MMA work — the 1-CTA kernel is ingest-bound
// (profiles/r1_microbench.md), this one moves the bound back to the tensor pipe.
// ==================================================================================================
template <int BN>
struct Gemm2Cfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;        // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;       // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = GEMM_SMEM_BUDGET / STAGE_BYTES > 8 ? 8 : GEMM_SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + 1024 + 256 + 4 * 4096;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(
                   bar_cluster_addr),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, int c0, int c1,
                                                uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_2sm(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int kAct = ACT_NONE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_b, bf16* __restrict__ C, int M,
                      int N, int K, int dbg) {
  constexpr bool kSwiGLU = kAct != ACT_NONE;
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + 256;  // 4 warps x 4 KB transpose patches

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = (M + 255) / 256;
  const int n_tiles = N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / GEMM_BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 2);   // (leader's copy is the live one) one arrive per CTA's producer
      mbar_init(empty + s, 1);  // multicast tcgen05.commit from the leader
    }
    mbar_init(tmem_full + 0, 1);
    mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty + 0, 8);  // 4 epilogue warps x 2 CTAs, all arriving at the leader
    mbar_init(tmem_empty + 1, 8);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before anybody signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of B, completion on the LEADER =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty + stage, phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const uint32_t lfull = mapa_shared(full + stage, 0);
          if (dbg & 1) {  // timing experiment: MMA on stale smem, no operand traffic
            mbar_arrive_cluster(lfull);
          } else {
            mbar_expect_tx_cluster(lfull, Cfg::STAGE_BYTES);
            tma_load_2d_2sm(sa, &tmap_a, kb * GEMM_BK, m_blk * 256 + (int)rank * 128, lfull);
            tma_load_2d_2sm(sb, &tmap_b, kb * GEMM_BK, n_blk * BN + (int)rank * (BN / 2), lfull);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_2sm(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(tmem_empty + acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              umma_bf16_2sm(tmem_d, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(empty + stage);
            if (kb == k_blocks - 1) umma_commit_2sm(tmem_full + acc);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 rows of the 256-row tile =====
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const int row0 = m_blk * 256 + (int)rank * 128 + quarter * 32;
      constexpr int OUT_BN = kSwiGLU ? BN / 2 : BN;
      const long long ldc = kSwiGLU ? N / 2 : N;
      bf16* ctile = C + (long long)row0 * ldc + (long long)n_blk * OUT_BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      if (!(dbg & 2))  // dbg&2: skip the drain (timing experiment)
        epilogue_rows<BN, kAct>(taddr, ctile, ldc, row0, M, epi_smem + (warp - 2) * EPI_SMEM_PER_WARP, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tmem_empty + acc, 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- host side: tensor maps + launch ---------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  long long rows, cols;
  int box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<long long>()(k.rows * 1315423911LL + k.cols * 2654435761LL + k.box_rows) +
         0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};

// bf16 row-major [rows, cols] -> tiled map with box {64 cols, box_rows rows}, 128B swizzle
static int get_tmap(const void* ptr, long long rows, long long cols, int box_rows,
                    CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, rows, cols, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return B200Q_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old / no driver)");
    return B200Q_ECUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%lld,%lld] box_rows=%d ptr=%p", (int)r,
              rows, cols, box_rows, ptr);
    return B200Q_ECUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return B200Q_OK;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int g_gemm_debug = 0;  // timing experiments only (wrong results): 1 = no TMA loads, 2 = no epilogue

static float* g_splitk_ws = nullptr;
static size_t g_splitk_ws_bytes = 0;

template <int BN, int kAct = ACT_NONE>
static int launch_gemm(const void* A, const void* W, void* C, int M, int N, int K,
                       cudaStream_t st, int splits = 1, const float** partials_out = nullptr) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  const int a_rows = M < GEMM_BM ? (M + 7) & ~7 : GEMM_BM;  // one m-tile: load only the rows that exist
  int rc = get_tmap(A, M, K, a_rows, &ta);
  if (rc) return rc;
  rc = get_tmap(W, N, K, BN, &tb);
  if (rc) return rc;
  auto kern = gemm_bf16_kernel<BN, kAct>;
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_TOTAL));
    attr_set = true;
  }
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * (N / BN) * splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  float* ws = nullptr;
  if (splits > 1) {
    // fixed-size scratch (8 splits x 256 rows x 8192 cols fp32 = 64 MiB), allocated once and never
    // moved: CUDA graphs captured by the engine bake this pointer into their kernel nodes
    const size_t need = (size_t)splits * M * N * sizeof(float);
    const size_t cap = (size_t)8 * 256 * 8192 * sizeof(float);
    B200Q_CHECK_ARG(need <= cap, "split-K scratch too small for M=%d N=%d splits=%d", M, N, splits);
    if (!g_splitk_ws) {
      B200Q_CUDA(cudaMalloc(&g_splitk_ws, cap));
      g_splitk_ws_bytes = cap;
    }
    ws = g_splitk_ws;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_TOTAL, st>>>(ta, tb, (bf16*)C, M, N, K, g_gemm_debug, splits, ws,
                                                    a_rows * GEMM_BK * 2);
  B200Q_LAUNCH_CHECK();
  if (partials_out) *partials_out = ws;  // the caller fuses the reduction into the consumer of C
  if (splits > 1 && !partials_out) {
    const long long mn8 = (long long)M * N / 8;
    const int rgrid = (int)((mn8 + 255) / 256 < 148 * 8 ? (mn8 + 255) / 256 : 148 * 8);
    splitk_reduce_kernel<<<rgrid, 256, 0, st>>>((const float4*)ws, (uint4*)C, mn8, splits);
    B200Q_LAUNCH_CHECK();
  }
  return B200Q_OK;
}

static int g_gemm2_pairs = 0;  // co-resident CTA pairs reported by the occupancy query

template <int BN, int kAct = ACT_NONE>
static int launch_gemm2(const void* A, const void* W, void* C, int M, int N, int K,
                        cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  CUtensorMap ta, tb;
  int rc = get_tmap(A, M, K, GEMM_BM, &ta);
  if (rc) return rc;
  rc = get_tmap(W, N, K, BN / 2, &tb);
  if (rc) return rc;
  auto kern = gemm2_bf16_kernel<BN, kAct>;
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_TOTAL));
    attr_set = true;
  }
  const int tiles = ((M + 255) / 256) * (N / BN);
  // the tile loop is static round-robin over CO-RESIDENT CTA pairs: size the grid by what the
  // hardware can really keep resident at once (GPC/TPC pairing can make this < SMs/2)
  static int pairs = 0;
  if (pairs == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() & ~1);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_TOTAL;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / 2;
    }
    pairs = n < num_sms() / 2 ? n : num_sms() / 2;
    g_gemm2_pairs = pairs;
  }
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_TOTAL, st>>>(ta, tb, (bf16*)C, M, N, K, g_gemm_debug);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

// Measured on B200 (profiles/r1_gemm_2cta.md): at M >= 256 the CTA-pair kernel with a 256 x 256
// tile is 0-8 % faster per unit of work than the 1-CTA 128 x 256 kernel (half the B traffic per
// SM).  Use it when its coarser tiles do not cost an extra round of the persistent loop:
// rounds(pair tiles over 74 pairs) <= rounds(1-CTA tiles over 148 SMs); a round costs the same
// wall time in both kernels (every SM works one 128 x 256 tile per round).
static bool prefer_2cta(int M, int N) {
  if (M < 256 || N % 256) return false;
  const long long sms = num_sms(), pairs = sms / 2;
  const long long t2 = (long long)((M + 255) / 256) * (N / 256);
  const long long t1 = (long long)((M + GEMM_BM - 1) / GEMM_BM) * (N / 256);
  if (t2 < pairs) return false;  // not even one full round of pairs: finer 1-CTA tiles / split-K fill the chip better
  return (t2 + pairs - 1) / pairs <= (t1 + sms - 1) / sms;
}

// ---- weight-stream probe (tools/gpu_probe.py stream_probe; not used by the product path) --------
// How fast can ONE CTA per SM pull a weight matrix through a shared-memory ring, as a function of the
// access pattern?  mode 0: the GEMM's own pattern — 2-D TMA boxes of `rows` weight rows x 64 columns
// (128 B per row, rows K*2 bytes apart); mode 1: the same bytes as 1-D bulk copies of contiguous
// rows*128-byte chunks (what a pre-tiled weight layout would allow).  No math: the consumer only
// releases the stage.  Answers whether the decode-sized GEMMs' ~35 GB/s per SM is the pattern's limit.
__global__ void __launch_bounds__(64)
    stream_probe_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ W, int N, int K,
                        int rows, int stages, int mode, unsigned long long* __restrict__ sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = rows * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int n_tiles = N / rows, kb_total = K / 64;
  const long long chunks = (long long)n_tiles * kb_total;   // one chunk = one stage fill
  if (threadIdx.x == 0) {          // producer
    int stage = 0;
    uint32_t phase = 0;
    for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
      mbar_wait(empty + stage, phase ^ 1u);
      mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
      if (mode == 0) {
        // consecutive chunks of a CTA walk K first (like a GEMM tile's k-loop)
        const long long tile = c / kb_total, kb = c % kb_total;
        tma_load_2d(smem + stage * stage_bytes, &tmap, (int)kb * 64, (int)tile * rows, full + stage);
      } else {
        tma_bulk_g2s(smem + stage * stage_bytes, W + c * stage_bytes, (uint32_t)stage_bytes, full + stage);
      }
      if (++stage == stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (threadIdx.x == 32) {  // consumer
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long acc = 0;
    for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
      mbar_wait(full + stage, phase);
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + stage * stage_bytes);
      mbar_arrive(empty + stage);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
  }
}

int g_gemm_force_bn = 0;  // test hook: 0 = heuristic
int g_gemm_splitk = 0;    // 0 = auto, 1 = never, n > 1 = force n splits where legal (tests)

static bool batch_invariant_env() {
  static const bool v = [] {
    const char* e = getenv("B200Q_BATCH_INVARIANT");
    return e && e[0] == '1';
  }();
  return v;
}

// split-K factor for a decode-sized projection (1 = do not split): when it has too few 128-wide
// output tiles to put every SM on the weight stream, split the reduction so that tiles x splits ~ #SMs
static int choose_splits(int M, int N, int K, bool invariant, int forced_bn) {
  const int want = invariant ? 1 : g_gemm_splitk;  // 0 = auto, 1 = off, >1 = forced (tests)
  if (want == 1 || forced_bn != 0 || M > 256 || N % 128 != 0 || N > 8192) return 1;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * (N / 128);
  const int kb = K / GEMM_BK;
  int splits = want > 1 ? want : 1;
  if (want == 0 && tiles * 2 <= num_sms()) {
    for (int sN = 8; sN >= 2; --sN)
      if (kb % sN == 0 && kb / sN >= 8 && tiles * sN <= num_sms() + num_sms() / 8) {
        splits = sN;
        break;
      }
  }
  return (splits > 1 && splits <= 8 && kb % splits == 0) ? splits : 1;
}
int g_gemm_mode = 0;      // 0 = auto, 1 = force 1-CTA kernels, 2 = force the 2-CTA kernel

}  // namespace b200q

using namespace b200q;

// shared body of the fused gate_up GEMMs (SwiGLU / GeGLU)
template <int kAct>
static int gemm_gated(const char* what, const void* A, const void* W, void* C, int M, int N, int K,
                      void* stream) {
  B200Q_CHECK_ARG(M >= 0 && N > 0 && K > 0 && K % GEMM_BK == 0 && N % 256 == 0,
                  "%s: unsupported shape M=%d N=%d K=%d (need K%%64==0, N%%256==0)", what, M, N, K);
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                  "%s: operands must be 16-byte aligned", what);
  if (M == 0) return B200Q_OK;
  cudaStream_t st = as_stream(stream);
  if ((g_gemm_mode == 2 && M > GEMM_BM) || (g_gemm_mode == 0 && prefer_2cta(M, N)))
    return launch_gemm2<256, kAct>(A, W, C, M, N, K, st);
  return launch_gemm<256, kAct>(A, W, C, M, N, K, st);
}

extern "C" {

// measurement hook (tools/gpu_probe.py stream_probe): stream W [N,K] bf16 through per-SM rings
// mode 0 = 2-D TMA boxes (rows x 64 cols), mode 1 = 1-D bulk copies of rows*128 contiguous bytes;
// chunk order, mode 0: CTA b takes (tile, k-block) pairs b, b+grid, ... with k fastest
int b200q_stream_probe(const void* W, int N, int K, int rows, int stages, int mode, int grid, void* stream) {
  B200Q_CHECK_ARG(W && N > 0 && K > 0 && K % 64 == 0 && rows > 0 && rows <= 256 && N % rows == 0 && stages >= 1 &&
                      stages * rows * 128 <= 200 * 1024 && grid > 0 && (mode == 0 || mode == 1),
                  "stream_probe: bad arguments");
  static unsigned long long* sink = nullptr;
  if (!sink) B200Q_CUDA(cudaMalloc(&sink, 8));
  CUtensorMap tm;
  int rc = get_tmap(W, N, K, rows, &tm);
  if (rc) return rc;
  const int smem_bytes = stages * rows * 128 + 1024 + 2 * stages * 8 + 64;
  static int attr_max = 0;
  if (smem_bytes > attr_max) {
    B200Q_CUDA(cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_max = smem_bytes;
  }
  stream_probe_kernel<<<grid, 64, smem_bytes, as_stream(stream)>>>(tm, (const uint8_t*)W, N, K, rows, stages, mode, sink);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

// test/tuning hook (not part of the reference-facing surface): force the N tile (0 = auto)
int b200q_gemm_set_tile_n(int bn) {
  B200Q_CHECK_ARG(bn == 0 || bn == 64 || bn == 128 || bn == 256, "gemm tile N must be 0/64/128/256");
  g_gemm_force_bn = bn;
  bump_tuning_epoch();
  return B200Q_OK;
}

// test/tuning hook: 0 = auto, 1 = 1-CTA kernels only, 2 = 2-CTA (cta_group::2) kernel whenever legal
// diagnostics: co-resident CTA pairs the 2-CTA kernel's grid is sized to (0 until first use)
int b200q_gemm_resident_pairs(void) { return g_gemm2_pairs; }

// timing experiments only — results are garbage: bit0 skips the operand loads, bit1 the epilogue
int b200q_gemm_set_debug(int flags) {
  g_gemm_debug = flags & 3;
  bump_tuning_epoch();
  return B200Q_OK;
}

// test/tuning hook: 0 = auto split-K, 1 = off, n > 1 = force n splits (decode-sized M only)
int b200q_gemm_set_splitk(int n) {
  B200Q_CHECK_ARG(n >= 0 && n <= 16, "split-K factor must be in [0,16]");
  g_gemm_splitk = n;
  bump_tuning_epoch();
  return B200Q_OK;
}

int b200q_gemm_set_mode(int mode) {
  B200Q_CHECK_ARG(mode >= 0 && mode <= 2, "gemm mode must be 0/1/2");
  g_gemm_mode = mode;
  bump_tuning_epoch();
  return B200Q_OK;
}

int b200q_gemm_bf16(const void* A, const void* W, void* C, int M, int N, int K, void* stream) {
  B200Q_CHECK_ARG(M >= 0 && N > 0 && K > 0 && K % GEMM_BK == 0 && N % 64 == 0,
                  "gemm: unsupported shape M=%d N=%d K=%d (need K%%64==0, N%%64==0)", M, N, K);
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                  "gemm: operands must be 16-byte aligned");
  if (M == 0) return B200Q_OK;
  cudaStream_t st = as_stream(stream);
  int bn = g_gemm_force_bn;
  if (g_gemm_mode == 2 && M > GEMM_BM) {
    if ((bn == 0 || bn == 256) && N % 256 == 0) return launch_gemm2<256>(A, W, C, M, N, K, st);
    if ((bn == 0 || bn == 128) && N % 128 == 0) return launch_gemm2<128>(A, W, C, M, N, K, st);
  }
  if (g_gemm_mode == 0 && bn == 0 && prefer_2cta(M, N)) return launch_gemm2<256>(A, W, C, M, N, K, st);
  {
    // split-K for decode-sized batches: when a projection has too few 128-wide output tiles to
    // put every SM on the weight stream, split the reduction so that tiles x splits ~ #SMs
    // B200Q_BATCH_INVARIANT=1 (read once): never split K, so every output element is reduced in
    // one sequential pass and a request's tokens cannot depend on its batch-mates (like
    // VLLM_BATCH_INVARIANT); costs decode-sized GEMMs up to 2x
    const int splits = choose_splits(M, N, K, batch_invariant_env(), bn);
    if (splits > 1) return launch_gemm<128>(A, W, C, M, N, K, st, splits);
  }
  if (bn == 0) {
    // Every CTA walks ceil(tiles / SMs) tiles; measured on B200 (profiles/r1_microbench.md) a
    // 128-wide tile costs ~0.85x and a 64-wide tile ~0.8x the time of a 256-wide one (the kernel
    // is bound by per-SM operand ingest, not by the MMA), so pick the candidate with the least
    // (tiles per CTA) x (relative tile time); ties go to the wider tile.
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    const int sms = num_sms();
    const int cand[3] = {256, 128, 64};
    const double cost[3] = {1.0, 0.85, 0.8};
    double best = 1e30;
    for (int i = 0; i < 3; ++i) {
      if (N % cand[i]) continue;
      const long long tiles = (long long)m_tiles * (N / cand[i]);
      const double t = (double)((tiles + sms - 1) / sms) * cost[i];
      if (t < best - 1e-9) {
        best = t;
        bn = cand[i];
      }
    }
  }
  if (N % bn != 0) bn = (N % 128 == 0) ? 128 : 64;
  if (bn == 256) return launch_gemm<256>(A, W, C, M, N, K, st);
  if (bn == 128) return launch_gemm<128>(A, W, C, M, N, K, st);
  return launch_gemm<64>(A, W, C, M, N, K, st);
}

// The split-K half of b200q_gemm_bf16 on its own: when the dispatcher would split the reduction
// (decode-sized M, few output tiles) the fp32 partial tiles are left in the library's scratch
// ([splits][M][N], *partials_out) and NO reduce pass runs — the caller folds the fixed-order sum and
// the bf16 rounding into the kernel that consumes the projection (b200q_add_rmsnorm_splitk,
// b200q_rope_kvwrite_splitk), saving a launch and a round trip per projection.  *splits_out = 1
// means nothing was launched: call b200q_gemm_bf16 instead.  The scratch is overwritten by the next
// split-K GEMM on the stream.
int b200q_gemm_bf16_splitk(const void* A, const void* W, int M, int N, int K, void* stream,
                           const float** partials_out, int* splits_out) {
  B200Q_CHECK_ARG(partials_out && splits_out, "gemm_splitk: null output argument");
  *partials_out = nullptr;
  *splits_out = 1;
  B200Q_CHECK_ARG(M >= 0 && N > 0 && K > 0 && K % GEMM_BK == 0 && N % 64 == 0,
                  "gemm: unsupported shape M=%d N=%d K=%d (need K%%64==0, N%%64==0)", M, N, K);
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                  "gemm: operands must be 16-byte aligned");
  if (M == 0 || g_gemm_mode == 2 || g_gemm_force_bn != 0 || (g_gemm_mode == 0 && prefer_2cta(M, N)))
    return B200Q_OK;
  const int splits = choose_splits(M, N, K, batch_invariant_env(), 0);
  if (splits <= 1) return B200Q_OK;
  int rc = launch_gemm<128>(A, W, nullptr, M, N, K, as_stream(stream), splits, partials_out);
  if (rc) return rc;
  *splits_out = splits;
  return B200Q_OK;
}

// K9+K10 fused: out[M, N/2] = swiglu(A . W^T) where W [N, K] holds gate/up rows interleaved in
// blocks of 128 (rows [256j, 256j+128) = gate rows [128j, 128j+128), next 128 = the matching up
// rows).  Same rounding points as b200q_gemm_bf16 followed by b200q_swiglu.
int b200q_gemm_swiglu_bf16(const void* A, const void* W, void* C, int M, int N, int K,
                           void* stream) {
  return gemm_gated<ACT_SWIGLU>("gemm_swiglu", A, W, C, M, N, K, stream);
}

// Gemma-2's GeGLU (gelu_tanh(gate) * up) in the same fused epilogue, same weight interleaving
int b200q_gemm_geglu_bf16(const void* A, const void* W, void* C, int M, int N, int K, void* stream) {
  return gemm_gated<ACT_GEGLU>("gemm_geglu", A, W, C, M, N, K, stream);
}

}  // extern "C"
