// attn.cu — paged-KV attention for the b200 worker.
//
// KV cache layout (per layer):  [num_blocks][2 (K,V)][n_kv][BS tokens][D] bf16, i.e. one
// contiguous "page" of BS*D*2 bytes per (block, K|V, kv-head).  Inside a token row the 16-byte
// chunk c lives at chunk position c ^ (token_in_block & 7).  The swizzle is applied by the
// RoPE+KV-write kernel (elementwise.cu), so a page can be pulled into shared memory with ONE
// 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) of exactly the valid rows and then be read
// with ldmatrix without bank conflicts — the block-table gather is done by the TMA engine, no
// thread touches KV bytes in flight.
//
// Decode  (K7): one CTA per (sequence, kv-head).  4 warps split the sequence's pages
//   round-robin; every warp owns a private 3-stage ring (K page + V page per stage) fed by its
//   own bulk copies, keeps an online-softmax state (m, l, O[G x D]) for the G query heads of
//   the GQA group in registers, and the 4 partial states are merged through shared memory.
//   QK^T and PV run on mma.sync m16n8k16 (rows 0..G-1 of the 16-row tile are the G heads);
//   the kernel is HBM-bound: algorithmic bytes = ctx * D * 2 (K,V) * 2 B per (seq, kv-head).
// Prefill (K6): one CTA per (q-tile of <=16 consecutive prompt tokens, kv-head).  G consumer
//   warps (one per query head of the group) share a 4-stage page ring filled by a producer
//   warp; causal masking by absolute position; K/V are read back from the paged cache, so
//   chunked prefill over an existing context needs no special case.
// Gemma-2 extras, compiled as separate instantiations so the Llama kernels keep
// their code: kCap = attention-logit soft-capping (score <- cap * tanh(score * scale / cap)), and a
// runtime sliding window (query at position p sees keys p - window < j <= p; pages that lie entirely
// before the window are never fetched).  D = 256 (Gemma-2 head size) uses 2-stage rings.
#include "common.cuh"

namespace b200q {

constexpr int DEC_WARPS = 4;
constexpr int PF_STAGES = 4;
// decode ring depth per warp: 3 pages of K+V in flight (2 at D = 256, where a page is 8 KB)
template <int D>
struct DecCfg {
  static constexpr int STAGES = D > 128 ? 2 : 3;
};

// score -> log2-domain logit.  plain: s * (scale * log2 e).  capped: tanh(s * scale / cap) * (cap * log2 e)
template <bool kCap>
__device__ __forceinline__ float score_xform(float s, float c0, float c1) {
  if (kCap) return tanhf(s * c0) * c1;
  return s * c0;
}

template <int D, int BS>
struct Geo {
  static constexpr int CPR = D / 8;            // 16 B chunks per token row
  static constexpr int ROW_BYTES = D * 2;
  static constexpr int PAGE_BYTES = BS * D * 2;
  static constexpr int NT = BS / 8;            // 8-token n-tiles per page
  static constexpr int KS = D / 16;            // k-steps of QK^T
  static constexpr int ND = D / 8;             // 8-dim n-tiles of O
};

// S[nt] (16 x 8 tokens) = Q(16 x D) . K_page^T for all n-tiles of one page.
template <int D, int BS, bool kFullRows>
__device__ __forceinline__ void qk_page(uint32_t k_s, const uint32_t (&qa)[D / 16][4],
                                        float (&s)[BS / 8][4], int lane) {
  using G_ = Geo<D, BS>;
#pragma unroll
  for (int nt = 0; nt < G_::NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const int token = nt * 8 + (lane & 7);
    const uint32_t row_addr = k_s + token * G_::ROW_BYTES;
#pragma unroll
    for (int j = 0; j < D / 32; ++j) {
      const int chunk = 4 * j + (lane >> 3);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(row_addr + ((chunk ^ (token & 7)) << 4), b0, b1, b2, b3);
      if (kFullRows) {
        mma_bf16_16816(s[nt], qa[2 * j][0], qa[2 * j][1], qa[2 * j][2], qa[2 * j][3], b0, b1);
        mma_bf16_16816(s[nt], qa[2 * j + 1][0], qa[2 * j + 1][1], qa[2 * j + 1][2],
                       qa[2 * j + 1][3], b2, b3);
      } else {
        mma_bf16_16816(s[nt], qa[2 * j][0], 0u, qa[2 * j][2], 0u, b0, b1);
        mma_bf16_16816(s[nt], qa[2 * j + 1][0], 0u, qa[2 * j + 1][2], 0u, b2, b3);
      }
    }
  }
}

// O(16 x D) += P(16 x BS) . V_page(BS x D); P given as fp32 accumulator fragments.
template <int D, int BS, bool kFullRows>
__device__ __forceinline__ void pv_page(uint32_t v_s, const float (&p)[BS / 8][4],
                                        float (&o)[D / 8][4], int lane) {
  using G_ = Geo<D, BS>;
#pragma unroll
  for (int kk = 0; kk < BS / 16; ++kk) {
    const uint32_t a0 = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    const uint32_t a2 = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    uint32_t a1 = 0u, a3 = 0u;
    if (kFullRows) {
      a1 = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
      a3 = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    }
    const int token = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
    const uint32_t row_addr = v_s + token * G_::ROW_BYTES;
#pragma unroll
    for (int nd2 = 0; nd2 < D / 16; ++nd2) {
      const int chunk = 2 * nd2 + (lane >> 4);
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(row_addr + ((chunk ^ (token & 7)) << 4), b0, b1, b2, b3);
      mma_bf16_16816(o[2 * nd2], a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(o[2 * nd2 + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// Which decode formulation a head size uses.  Measured on B200 (bench, Llama-3-8B, D = 128): the
// row-major PV form streams 6.2 TB/s, the transposed form 5.8 TB/s (two extra shuffles sit on the
// per-page dependency chain); at D = 256 the row-major form needs 128 accumulator registers, spills
// so the transposed form (64 accumulator registers, no spills) is used there: 6.3 TB/s on Gemma-2-9B.
template <int D>
constexpr bool kTransposedPV = D >= 256;

// Decode: O^T(D x 8 heads) += V^T(D x BS tokens) . P^T(BS x 8 heads).  With the G <= 8 heads on
// the N side the accumulator has no unused rows (D/16 tiles x 4 registers instead of D/8 x 4) and
// PV needs D/16 MMAs per page instead of D/8.  The S accumulators of qk_page<false> (rows = heads,
// cols = tokens) are exactly the B fragments of P^T: b0 = tokens 0-7, b1 = tokens 8-15.  A = V^T
// comes from the swizzled V page with ldmatrix.trans: matrix i of the x4 load covers tokens
// 8*(i>>1).. and the 16-byte chunk 2*tile + (i&1).  (lane-accurate check: tools/emu/)
template <int D, int BS>
__device__ __forceinline__ void pv_page_t(uint32_t v_s, const float (&p)[BS / 8][4],
                                          float (&o)[D / 16][4], int lane) {
  static_assert(BS == 16, "one 16-token page = one k-step of the PV MMA");
  using G_ = Geo<D, BS>;
  const uint32_t b0 = pack_bf16x2(p[0][0], p[0][1]);
  const uint32_t b1 = pack_bf16x2(p[1][0], p[1][1]);
  const int mi = lane >> 3;
  const int token = (mi >> 1) * 8 + (lane & 7);
  const uint32_t row_addr = v_s + token * G_::ROW_BYTES;
#pragma unroll
  for (int tile = 0; tile < D / 16; ++tile) {
    const int chunk = 2 * tile + (mi & 1);
    uint32_t a0, a1, a2, a3;
    ldsm_x4_t(row_addr + ((chunk ^ (token & 7)) << 4), a0, a1, a2, a3);
    mma_bf16_16816(o[tile], a0, a1, a2, a3, b0, b1);
  }
}

// O^T accumulators hold heads 2(lane&3), +1 in their columns; a per-head value (alpha, 1/l) lives
// in the lanes of row `head` of the S layout, i.e. lanes 4*head .. 4*head+3
__device__ __forceinline__ void per_head_pair(float v, int lane, float& h0, float& h1) {
  const int cq = (lane & 3) * 2;
  h0 = __shfl_sync(0xffffffffu, v, cq * 4);
  h1 = __shfl_sync(0xffffffffu, v, cq * 4 + 4);
}

template <int D>
__device__ __forceinline__ void scale_ot(float (&o)[D / 16][4], float h0, float h1) {
#pragma unroll
  for (int i = 0; i < D / 16; ++i) {
    o[i][0] *= h0;
    o[i][1] *= h1;
    o[i][2] *= h0;
    o[i][3] *= h1;
  }
}

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// zero rows [n_valid, BS) of a V page in shared memory (whole warp), so that masked
// probabilities (exactly 0) never meet stale NaN/Inf bit patterns in the PV mma.
template <int D, int BS>
__device__ __forceinline__ void zero_tail_rows(uint8_t* v_page, int n_valid, int lane) {
  using G_ = Geo<D, BS>;
  const int n = (BS - n_valid) * G_::CPR;
  uint4* p = reinterpret_cast<uint4*>(v_page + n_valid * G_::ROW_BYTES);
  for (int i = lane; i < n; i += 32) p[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------
// K7 decode
// ------------------------------------------------------------------------------------------
template <int D, int BS>
struct DecSmem {
  static constexpr int STAGES = DecCfg<D>::STAGES;
  static constexpr int STAGE_BYTES = 2 * Geo<D, BS>::PAGE_BYTES;
  static constexpr int RING_BYTES = DEC_WARPS * STAGES * STAGE_BYTES;
  static constexpr int MERGE_FLOATS = DEC_WARPS * 8 * (D + 2);
  static constexpr int TOTAL = RING_BYTES + MERGE_FLOATS * 4 + DEC_WARPS * STAGES * 8;
};

template <int D, int BS, bool kCap>
__global__ void __launch_bounds__(DEC_WARPS * 32, 2)
    decode_attn_kernel(const bf16* __restrict__ q, int q_stride, bf16* __restrict__ out,
                       const uint8_t* __restrict__ kv_layer,
                       const int32_t* __restrict__ block_table, int bt_stride,
                       const int32_t* __restrict__ ctx_lens, int n_q, int n_kv, int G,
                       float c0, float c1, int window) {
  using G_ = Geo<D, BS>;
  using S_ = DecSmem<D, BS>;
  constexpr int DEC_STAGES = S_::STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  const int seq = blockIdx.x, kvh = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + warp * DEC_STAGES * S_::STAGE_BYTES;
  float* merge = reinterpret_cast<float*>(smem + S_::RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_::RING_BYTES + S_::MERGE_FLOATS * 4) +
                   warp * DEC_STAGES;

  const int ctx = __ldg(ctx_lens + seq);
  // the query sits at position ctx-1 and sees keys [lo, ctx); pages before lo's page are skipped
  const int lo = (window > 0 && ctx > window) ? ctx - window : 0;
  const int page0 = lo / BS;
  const int n_pages = (ctx + BS - 1) / BS - page0;
  const int my_n = n_pages > warp ? (n_pages - warp + DEC_WARPS - 1) / DEC_WARPS : 0;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < DEC_STAGES; ++s) mbar_init(bars + s, 1);
    mbar_fence_init();
  }
  __syncwarp();

  const int32_t* bt = block_table + (long long)seq * bt_stride + page0;
  const long long page_stride = (long long)G_::PAGE_BYTES;  // bytes per (block,kv,head) page
  // page address: ((blk*2 + kv) * n_kv + kvh) * PAGE_BYTES
  int ids_base = 0;
  int ids = (lane < my_n) ? __ldg(bt + warp + DEC_WARPS * lane) : 0;

  auto issue = [&](int k) {  // whole warp calls; k = warp-local page index
    if (k - ids_base >= 32) {
      ids_base += 32;
      ids = (ids_base + lane < my_n) ? __ldg(bt + warp + DEC_WARPS * (ids_base + lane)) : 0;
    }
    const int blk = __shfl_sync(0xffffffffu, ids, k - ids_base);
    const int p = page0 + warp + DEC_WARPS * k;
    const int n_valid = min(BS, ctx - p * BS);
    const int st = k % DEC_STAGES;
    uint8_t* ks = ring + st * S_::STAGE_BYTES;
    uint8_t* vs = ks + G_::PAGE_BYTES;
    if (n_valid < BS) {
      zero_tail_rows<D, BS>(vs, n_valid, lane);
      __syncwarp();
    }
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)n_valid * G_::ROW_BYTES;
      const uint8_t* kp = kv_layer + (((long long)blk * 2 + 0) * n_kv + kvh) * page_stride;
      const uint8_t* vp = kv_layer + (((long long)blk * 2 + 1) * n_kv + kvh) * page_stride;
      mbar_expect_tx(bars + st, 2 * bytes);
      tma_bulk_g2s(ks, kp, bytes, bars + st);
      tma_bulk_g2s(vs, vp, bytes, bars + st);
    }
  };

  const int n_pro = my_n < DEC_STAGES ? my_n : DEC_STAGES;
  for (int k = 0; k < n_pro; ++k) issue(k);

  // Q fragments: rows 0..G-1 of the 16-row tile are the G query heads of this kv head
  uint32_t qa[D / 16][4];
  {
    const int r = lane >> 2, cq = (lane & 3) * 2;
    const bf16* qrow = q + (long long)seq * q_stride + (long long)(kvh * G + r) * D;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      qa[ks][0] = r < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + cq) : 0u;
      qa[ks][2] = r < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + cq) : 0u;
      qa[ks][1] = 0u;
      qa[ks][3] = 0u;
    }
  }

  // TPV (D = 256): O^T tiles [dim = 16 i + lane/4 (+8)][head = 2(lane&3), +1], D/16 x 4 registers;
  // else O tiles [head = lane/4 (rows 8-15 unused)][dim = 8 i + 2(lane&3), +1], D/8 x 4 registers
  constexpr bool TPV = kTransposedPV<D>;
  constexpr int NO = TPV ? D / 16 : D / 8;
  float o[NO][4];
#pragma unroll
  for (int i = 0; i < NO; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m = -INFINITY, l = 0.f;

  for (int k = 0; k < my_n; ++k) {
    const int st = k % DEC_STAGES;
    const uint32_t phase = (uint32_t)(k / DEC_STAGES) & 1u;
    mbar_wait(bars + st, phase);
    const int p = page0 + warp + DEC_WARPS * k;
    const int n_valid = min(BS, ctx - p * BS);
    const int t_lo = lo - p * BS;  // first visible token of this page (<= 0 except on the window's first page)
    const uint32_t k_s = smem_u32(ring + st * S_::STAGE_BYTES);
    const uint32_t v_s = k_s + G_::PAGE_BYTES;

    float s[BS / 8][4];
    qk_page<D, BS, false>(k_s, qa, s, lane);

    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < G_::NT; ++nt) {
      const int t0 = nt * 8 + (lane & 3) * 2;
      s[nt][0] = (t0 < n_valid && t0 >= t_lo) ? score_xform<kCap>(s[nt][0], c0, c1) : -INFINITY;
      s[nt][1] = (t0 + 1 < n_valid && t0 + 1 >= t_lo) ? score_xform<kCap>(s[nt][1], c0, c1) : -INFINITY;
      mx = fmaxf(mx, fmaxf(s[nt][0], s[nt][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m, mx);  // finite: every page has >= 1 valid token
    const float alpha = exp2f(m - m_new);
    float psum = 0.f;
#pragma unroll
    for (int nt = 0; nt < G_::NT; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m_new);
      s[nt][1] = exp2f(s[nt][1] - m_new);
      s[nt][2] = 0.f;
      s[nt][3] = 0.f;
      psum += s[nt][0] + s[nt][1];
    }
    l = l * alpha + psum;
    m = m_new;
    if constexpr (TPV) {
      float a0, a1;
      per_head_pair(alpha, lane, a0, a1);
      scale_ot<D>(o, a0, a1);
      pv_page_t<D, BS>(v_s, s, o, lane);
    } else {
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        o[i][0] *= alpha;
        o[i][1] *= alpha;
      }
      pv_page<D, BS, false>(v_s, s, o, lane);
    }
    __syncwarp();
    if (k + DEC_STAGES < my_n) issue(k + DEC_STAGES);
  }

  // ---- merge the 4 per-warp partial states ----
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  {
    const int r = lane >> 2;
    if constexpr (TPV) {
      // partial state of this warp: O^T columns are heads 2(lane&3), +1; m and l sit in the S layout
      const int h0 = (lane & 3) * 2;
      float* m0 = merge + (warp * 8 + h0) * (D + 2);
      float* m1 = m0 + (D + 2);
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        if (h0 < G) {
          m0[i * 16 + r] = o[i][0];
          m0[i * 16 + 8 + r] = o[i][2];
        }
        if (h0 + 1 < G) {
          m1[i * 16 + r] = o[i][1];
          m1[i * 16 + 8 + r] = o[i][3];
        }
      }
    } else if (r < G) {
      float* mo = merge + (warp * 8 + r) * (D + 2);
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        mo[i * 8 + (lane & 3) * 2] = o[i][0];
        mo[i * 8 + (lane & 3) * 2 + 1] = o[i][1];
      }
    }
    if (r < G && (lane & 3) == 0) {
      float* mo = merge + (warp * 8 + r) * (D + 2);
      mo[D] = m;
      mo[D + 1] = l;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * D; idx += DEC_WARPS * 32) {
    const int r = idx / D, d = idx % D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) M = fmaxf(M, merge[(w * 8 + r) * (D + 2) + D]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float* mo = merge + (w * 8 + r) * (D + 2);
      const float wgt = (mo[D] == -INFINITY) ? 0.f : exp2f(mo[D] - M);
      num += wgt * mo[d];
      den += wgt * mo[D + 1];
    }
    const float v = den > 0.f ? num / den : 0.f;
    out[(long long)seq * n_q * D + (long long)(kvh * G + r) * D + d] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// K7 decode, streaming variant for large batches.  Every warp of a persistent grid owns whole
// (sequence, kv-head) items and walks them back to back: the 3-stage TMA page ring never
// drains between items (the pages of the next item are already in flight while the current
// one finishes), there is no cross-warp merge and no CTA turnover, and the next item's Q
// fragments / context length are prefetched one item ahead.  With ~12 pages per item
// (128-in/128-out) the v1 kernel above spends as long in its prologue (ctx -> block ids -> first
// TMA, ~1-2 us of dependent latency) as in streaming; this one pays that once per warp.
// Used when there are enough items to give every resident warp several of them; v1 keeps the
// small-batch / long-context cases (it splits one item across 4 warps).
// ------------------------------------------------------------------------------------------
template <int D, int BS>
struct Dec2Smem {
  static constexpr int STAGES = DecCfg<D>::STAGES;
  static constexpr int STAGE_BYTES = 2 * Geo<D, BS>::PAGE_BYTES;
  static constexpr int RING_BYTES = DEC_WARPS * STAGES * STAGE_BYTES;
  static constexpr int TOTAL = RING_BYTES + DEC_WARPS * STAGES * 8;
};

template <int D, int BS, bool kCap>
__global__ void __launch_bounds__(DEC_WARPS * 32, 2)
    decode_attn_stream_kernel(const bf16* __restrict__ q, int q_stride, bf16* __restrict__ out,
                              const uint8_t* __restrict__ kv_layer,
                              const int32_t* __restrict__ block_table, int bt_stride,
                              const int32_t* __restrict__ ctx_lens, int n_seqs, int n_q, int n_kv,
                              int G, float c0, float c1, int window) {
  using G_ = Geo<D, BS>;
  using S_ = Dec2Smem<D, BS>;
  constexpr int DEC_STAGES = S_::STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + warp * DEC_STAGES * S_::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_::RING_BYTES) + warp * DEC_STAGES;
  const int total = n_seqs * n_kv;
  const int n_warps = gridDim.x * DEC_WARPS;
  const int first = blockIdx.x * DEC_WARPS + warp;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < DEC_STAGES; ++s) mbar_init(bars + s, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (first >= total) return;

  const long long page_stride = (long long)G_::PAGE_BYTES;
  const int r = lane >> 2, cq = (lane & 3) * 2;

  // ---- issue cursor: runs up to DEC_STAGES pages ahead of the consumer, across item borders ----
  // with a sliding window an item's pages run from first_page(ctx) to the last one
  auto first_page = [&](int ctx) { return (window > 0 && ctx > window) ? (ctx - window) / BS : 0; };
  int i_item = first, i_ctx = max(__ldg(ctx_lens + first / n_kv), 1);
  int i_page = first_page(i_ctx);
  int i_npages = (i_ctx + BS - 1) / BS;
  int i_ids_base = i_page;
  int i_ids = (i_ids_base + lane < i_npages)
                  ? __ldg(block_table + (long long)(first / n_kv) * bt_stride + i_ids_base + lane)
                  : 0;
  int issued = 0;
  bool i_done = false;
  auto issue_next = [&]() {  // whole warp
    if (i_done) return;
    if (i_page - i_ids_base >= 32) {
      i_ids_base += 32;
      i_ids = (i_ids_base + lane < i_npages)
                  ? __ldg(block_table + (long long)(i_item / n_kv) * bt_stride + i_ids_base + lane)
                  : 0;
    }
    const int blk = __shfl_sync(0xffffffffu, i_ids, i_page - i_ids_base);
    const int kvh = i_item % n_kv;
    const int n_valid = min(BS, i_ctx - i_page * BS);
    const int st = issued % DEC_STAGES;
    uint8_t* ks = ring + st * S_::STAGE_BYTES;
    uint8_t* vs = ks + G_::PAGE_BYTES;
    if (n_valid < BS) {
      zero_tail_rows<D, BS>(vs, n_valid, lane);
      __syncwarp();
    }
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)n_valid * G_::ROW_BYTES;
      const uint8_t* kp = kv_layer + (((long long)blk * 2 + 0) * n_kv + kvh) * page_stride;
      const uint8_t* vp = kv_layer + (((long long)blk * 2 + 1) * n_kv + kvh) * page_stride;
      mbar_expect_tx(bars + st, 2 * bytes);
      tma_bulk_g2s(ks, kp, bytes, bars + st);
      tma_bulk_g2s(vs, vp, bytes, bars + st);
    }
    ++issued;
    if (++i_page == i_npages) {  // move to this warp's next item
      i_item += n_warps;
      if (i_item >= total) {
        i_done = true;
        return;
      }
      const int seq = i_item / n_kv;
      i_ctx = max(__ldg(ctx_lens + seq), 1);
      i_npages = (i_ctx + BS - 1) / BS;
      i_page = first_page(i_ctx);
      i_ids_base = i_page;
      i_ids = (i_ids_base + lane < i_npages)
                  ? __ldg(block_table + (long long)seq * bt_stride + i_ids_base + lane)
                  : 0;
    }
  };
  for (int k = 0; k < DEC_STAGES; ++k) issue_next();

  // ---- consumer ----
  auto load_q = [&](int item, uint32_t (&dst)[D / 16][2]) {
    const int seq = item / n_kv, kvh = item % n_kv;
    const bf16* qrow = q + (long long)seq * q_stride + (long long)(kvh * G + r) * D;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      dst[ks][0] = r < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + cq) : 0u;
      dst[ks][1] = r < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + cq) : 0u;
    }
  };
  uint32_t qn[D / 16][2];  // prefetched Q of the NEXT item
  int n_ctx = max(__ldg(ctx_lens + first / n_kv), 1);
  load_q(first, qn);
  int consumed = 0;
  for (int item = first; item < total; item += n_warps) {
    uint32_t qa[D / 16][4];
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      qa[ks][0] = qn[ks][0];
      qa[ks][2] = qn[ks][1];
      qa[ks][1] = 0u;
      qa[ks][3] = 0u;
    }
    const int ctx = n_ctx;
    const int n_pages = (ctx + BS - 1) / BS;
    const int lo = (window > 0 && ctx > window) ? ctx - window : 0;
    if (item + n_warps < total) {  // prefetch the next item's Q and context length
      n_ctx = max(__ldg(ctx_lens + (item + n_warps) / n_kv), 1);
      load_q(item + n_warps, qn);
    }
    // TPV (D = 256): O^T tiles [dim = 16 i + lane/4 (+8)][head = 2(lane&3), +1], D/16 x 4 registers;
    // else O tiles [head = lane/4 (rows 8-15 unused)][dim = 8 i + 2(lane&3), +1], D/8 x 4 registers
    constexpr bool TPV = kTransposedPV<D>;
    constexpr int NO = TPV ? D / 16 : D / 8;
    float o[NO][4];
#pragma unroll
    for (int i = 0; i < NO; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int p = lo / BS; p < n_pages; ++p) {
      const int st = consumed % DEC_STAGES;
      mbar_wait(bars + st, (uint32_t)(consumed / DEC_STAGES) & 1u);
      const int n_valid = min(BS, ctx - p * BS);
      const int t_lo = lo - p * BS;
      const uint32_t k_s = smem_u32(ring + st * S_::STAGE_BYTES);
      const uint32_t v_s = k_s + G_::PAGE_BYTES;
      float s[BS / 8][4];
      qk_page<D, BS, false>(k_s, qa, s, lane);
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < G_::NT; ++nt) {
        const int t0 = nt * 8 + cq;
        s[nt][0] = (t0 < n_valid && t0 >= t_lo) ? score_xform<kCap>(s[nt][0], c0, c1) : -INFINITY;
        s[nt][1] = (t0 + 1 < n_valid && t0 + 1 >= t_lo) ? score_xform<kCap>(s[nt][1], c0, c1) : -INFINITY;
        mx = fmaxf(mx, fmaxf(s[nt][0], s[nt][1]));
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m, mx);
      const float alpha = exp2f(m - m_new);
      float psum = 0.f;
#pragma unroll
      for (int nt = 0; nt < G_::NT; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - m_new);
        s[nt][1] = exp2f(s[nt][1] - m_new);
        s[nt][2] = 0.f;
        s[nt][3] = 0.f;
        psum += s[nt][0] + s[nt][1];
      }
      l = l * alpha + psum;
      m = m_new;
      if constexpr (TPV) {
        float a0, a1;
        per_head_pair(alpha, lane, a0, a1);
        scale_ot<D>(o, a0, a1);
        pv_page_t<D, BS>(v_s, s, o, lane);
      } else {
#pragma unroll
        for (int i = 0; i < NO; ++i) {
          o[i][0] *= alpha;
          o[i][1] *= alpha;
        }
        pv_page<D, BS, false>(v_s, s, o, lane);
      }
      __syncwarp();
      ++consumed;
      issue_next();  // refill the stage that was just drained
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int seq = item / n_kv, kvh = item % n_kv;
    bf16* orow = out + (long long)seq * n_q * D + (long long)(kvh * G + r) * D;
    if constexpr (TPV) {
      float i0, i1;
      per_head_pair(inv, lane, i0, i1);
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        // transpose the bf16 tile halves back to [head = lane/4][dims 2(lane&3), +1] for 4-byte stores
        const uint32_t d0 = movmatrix_trans(pack_bf16x2(o[i][0] * i0, o[i][1] * i1));
        const uint32_t d1 = movmatrix_trans(pack_bf16x2(o[i][2] * i0, o[i][3] * i1));
        if (r < G) {
          *reinterpret_cast<uint32_t*>(orow + i * 16 + cq) = d0;
          *reinterpret_cast<uint32_t*>(orow + i * 16 + 8 + cq) = d1;
        }
      }
    } else if (r < G) {
#pragma unroll
      for (int i = 0; i < NO; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8 + cq) = pack_bf16x2(o[i][0] * inv, o[i][1] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K6 prefill
// ------------------------------------------------------------------------------------------
template <int D, int BS>
struct PfSmem {
  static constexpr int STAGE_BYTES = 2 * Geo<D, BS>::PAGE_BYTES;
  static constexpr int RING_BYTES = PF_STAGES * STAGE_BYTES;
  static constexpr int TOTAL = RING_BYTES + 2 * PF_STAGES * 8;
};

// consumer warps + 1 producer: G <= 8 (G <= 4 at D = 256, whose O accumulators need the registers)
template <int D>
struct PfCfg {
  static constexpr int MAX_G = D > 128 ? 4 : 8;
};

template <int D, int BS, bool kCap>
__global__ void __launch_bounds__((PfCfg<D>::MAX_G + 1) * 32)
    prefill_attn_kernel(const bf16* __restrict__ q, int q_stride, bf16* __restrict__ out,
                        const uint8_t* __restrict__ kv_layer,
                        const int32_t* __restrict__ block_table, int bt_stride,
                        const int4* __restrict__ tiles, int n_q, int n_kv, int G,
                        float c0, float c1, int window) {
  using G_ = Geo<D, BS>;
  using S_ = PfSmem<D, BS>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S_::RING_BYTES);
  uint64_t* empty = full + PF_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvh = blockIdx.y;
  const int4 tile = __ldg(tiles + blockIdx.x);
  const int bt_row = tile.x, row0 = tile.y, n_rows = tile.z, pos0 = tile.w;
  const int total_kv = pos0 + n_rows;  // keys 0 .. pos0+n_rows-1 are visible to this tile
  const int n_pages = (total_kv + BS - 1) / BS;
  // sliding window: row i (position pos0+i) sees keys j > pos0+i-window; the tile starts at the
  // page holding row 0's first visible key
  const int page0 = (window > 0 && pos0 + 1 > window) ? (pos0 + 1 - window) / BS : 0;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < PF_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, G);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == G) {
    // ===== producer warp: TMA page gather through the block table =====
    const int32_t* bt = block_table + (long long)bt_row * bt_stride;
    for (int p = page0; p < n_pages; ++p) {
      const int it = p - page0;
      const int st = it % PF_STAGES;
      if (it >= PF_STAGES) mbar_wait(empty + st, (uint32_t)((it / PF_STAGES) - 1) & 1u);
      const int n_valid = min(BS, total_kv - p * BS);
      uint8_t* ks = smem + st * S_::STAGE_BYTES;
      uint8_t* vs = ks + G_::PAGE_BYTES;
      if (n_valid < BS) {
        zero_tail_rows<D, BS>(vs, n_valid, lane);
        __syncwarp();
      }
      if (lane == 0) {
        const int blk = __ldg(bt + p);
        const uint32_t bytes = (uint32_t)n_valid * G_::ROW_BYTES;
        const uint8_t* kp =
            kv_layer + (((long long)blk * 2 + 0) * n_kv + kvh) * (long long)G_::PAGE_BYTES;
        const uint8_t* vp =
            kv_layer + (((long long)blk * 2 + 1) * n_kv + kvh) * (long long)G_::PAGE_BYTES;
        mbar_expect_tx(full + st, 2 * bytes);
        tma_bulk_g2s(ks, kp, bytes, full + st);
        tma_bulk_g2s(vs, vp, bytes, full + st);
      }
      __syncwarp();
    }
    return;
  }
  if (warp > G) return;

  // ===== consumer warp `warp` = query head kvh*G + warp; 16 rows = 16 tokens of the tile =====
  const int head = kvh * G + warp;
  const int r = lane >> 2, cq = (lane & 3) * 2;
  uint32_t qa[D / 16][4];
  {
    const bf16* q0 = q + (long long)(row0 + r) * q_stride + (long long)head * D;
    const bf16* q1 = q + (long long)(row0 + r + 8) * q_stride + (long long)head * D;
    const bool v0 = r < n_rows, v1 = r + 8 < n_rows;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      qa[ks][0] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + cq) : 0u;
      qa[ks][1] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + cq) : 0u;
      qa[ks][2] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8 + cq) : 0u;
      qa[ks][3] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8 + cq) : 0u;
    }
  }
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  // last visible key index per row (causal), clamped to the keys that exist; first visible key
  // under the sliding window
  const int lim0 = min(pos0 + r, total_kv - 1), lim1 = min(pos0 + r + 8, total_kv - 1);
  const int lo0 = window > 0 ? pos0 + r + 1 - window : 0, lo1 = window > 0 ? lo0 + 8 : 0;

  for (int p = page0; p < n_pages; ++p) {
    const int it = p - page0;
    const int st = it % PF_STAGES;
    mbar_wait(full + st, (uint32_t)(it / PF_STAGES) & 1u);
    const uint32_t k_s = smem_u32(smem + st * S_::STAGE_BYTES);
    const uint32_t v_s = k_s + G_::PAGE_BYTES;
    float s[BS / 8][4];
    qk_page<D, BS, true>(k_s, qa, s, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < G_::NT; ++nt) {
      const int j = p * BS + nt * 8 + cq;
      s[nt][0] = (j <= lim0 && j >= lo0) ? score_xform<kCap>(s[nt][0], c0, c1) : -INFINITY;
      s[nt][1] = (j + 1 <= lim0 && j + 1 >= lo0) ? score_xform<kCap>(s[nt][1], c0, c1) : -INFINITY;
      s[nt][2] = (j <= lim1 && j >= lo1) ? score_xform<kCap>(s[nt][2], c0, c1) : -INFINITY;
      s[nt][3] = (j + 1 <= lim1 && j + 1 >= lo1) ? score_xform<kCap>(s[nt][3], c0, c1) : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    // without a window key 0 is visible to every row and the maxima are finite after page 0; with
    // one, a row may not have met a visible key yet (m = -inf): subtract 0 instead of -inf then, so
    // that alpha and the probabilities come out as exact zeros, not NaN
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float b0 = mn0 == -INFINITY ? 0.f : mn0, b1 = mn1 == -INFINITY ? 0.f : mn1;
    const float a0 = exp2f(m0 - b0), a1 = exp2f(m1 - b1);
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < G_::NT; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - b0);
      s[nt][1] = exp2f(s[nt][1] - b0);
      s[nt][2] = exp2f(s[nt][2] - b1);
      s[nt][3] = exp2f(s[nt][3] - b1);
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * a0 + ps0;
    l1 = l1 * a1 + ps1;
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= a0;
      o[i][1] *= a0;
      o[i][2] *= a1;
      o[i][3] *= a1;
    }
    pv_page<D, BS, true>(v_s, s, o, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + st);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  bf16* o0 = out + (long long)(row0 + r) * n_q * D + (long long)head * D;
  bf16* o1 = out + (long long)(row0 + r + 8) * n_q * D + (long long)head * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    if (r < n_rows)
      *reinterpret_cast<uint32_t*>(o0 + i * 8 + cq) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
    if (r + 8 < n_rows)
      *reinterpret_cast<uint32_t*>(o1 + i * 8 + cq) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
  }
}

// c0/c1 of score_xform for a given scale / soft cap
struct ScoreConsts {
  float c0, c1;
};
static inline ScoreConsts score_consts(float scale, float softcap) {
  const float log2e = 1.4426950408889634f;
  if (softcap > 0.f) return {scale / softcap, softcap * log2e};
  return {scale * log2e, 0.f};
}

template <int D, int BS, bool kCap>
static int launch_decode(const void* q, int q_stride, void* out, const void* kv,
                         const int32_t* bt, int bt_stride, const int32_t* ctx, int n_seqs, int n_q,
                         int n_kv, float scale, float softcap, int window, cudaStream_t st) {
  using S_ = DecSmem<D, BS>;
  auto kern = decode_attn_kernel<D, BS, kCap>;
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S_::TOTAL));
    attr_set = true;
  }
  const ScoreConsts sc = score_consts(scale, softcap);
  dim3 grid(n_seqs, n_kv);
  kern<<<grid, DEC_WARPS * 32, S_::TOTAL, st>>>((const bf16*)q, q_stride, (bf16*)out,
                                                (const uint8_t*)kv, bt, bt_stride, ctx, n_q, n_kv,
                                                n_q / n_kv, sc.c0, sc.c1, window);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int g_decode_variant = 0;  // test hook: 0 = auto, 1 = split kernel (v1), 2 = streaming kernel

static int attn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int D, int BS, bool kCap>
static int launch_decode_stream(const void* q, int q_stride, void* out, const void* kv,
                                const int32_t* bt, int bt_stride, const int32_t* ctx, int n_seqs,
                                int n_q, int n_kv, float scale, float softcap, int window,
                                cudaStream_t st) {
  using S_ = Dec2Smem<D, BS>;
  auto kern = decode_attn_stream_kernel<D, BS, kCap>;
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S_::TOTAL));
    attr_set = true;
  }
  const int total = n_seqs * n_kv;
  // 2 resident CTAs per SM (96 KB of page ring each); variant 3 (tests) squeezes everything
  // through 4 CTAs so that every warp walks many items
  const int max_ctas = g_decode_variant == 3 ? 4 : 2 * attn_num_sms();
  const int want = (total + DEC_WARPS - 1) / DEC_WARPS;
  const int grid = want < max_ctas ? want : max_ctas;
  const ScoreConsts sc = score_consts(scale, softcap);
  kern<<<grid, DEC_WARPS * 32, S_::TOTAL, st>>>((const bf16*)q, q_stride, (bf16*)out,
                                                (const uint8_t*)kv, bt, bt_stride, ctx, n_seqs, n_q,
                                                n_kv, n_q / n_kv, sc.c0, sc.c1, window);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

template <int D, int BS, bool kCap>
static int launch_prefill(const void* q, int q_stride, void* out, const void* kv,
                          const int32_t* bt, int bt_stride, const int32_t* tiles, int n_tiles,
                          int n_q, int n_kv, float scale, float softcap, int window,
                          cudaStream_t st) {
  using S_ = PfSmem<D, BS>;
  auto kern = prefill_attn_kernel<D, BS, kCap>;
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S_::TOTAL));
    attr_set = true;
  }
  const int G = n_q / n_kv;
  B200Q_CHECK_ARG(G <= PfCfg<D>::MAX_G, "prefill_attn: GQA group %d too large for D=%d (max %d)", G, D,
                  PfCfg<D>::MAX_G);
  const ScoreConsts sc = score_consts(scale, softcap);
  dim3 grid(n_tiles, n_kv);
  kern<<<grid, (G + 1) * 32, S_::TOTAL, st>>>((const bf16*)q, q_stride, (bf16*)out,
                                              (const uint8_t*)kv, bt, bt_stride,
                                              (const int4*)tiles, n_q, n_kv, G, sc.c0, sc.c1, window);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

// (D, soft-cap) dispatch: the capped variants exist for every D so that small test models can use
// them; the uncapped D = 64 / 128 instantiations are the Llama kernels.
#define B200Q_ATTN_DISPATCH(FN, ...)                                     \
  do {                                                                   \
    const bool cap = softcap > 0.f;                                      \
    if (D == 64) return cap ? FN<64, 16, true>(__VA_ARGS__) : FN<64, 16, false>(__VA_ARGS__);    \
    if (D == 128) return cap ? FN<128, 16, true>(__VA_ARGS__) : FN<128, 16, false>(__VA_ARGS__); \
    return cap ? FN<256, 16, true>(__VA_ARGS__) : FN<256, 16, false>(__VA_ARGS__);               \
  } while (0)

}  // namespace b200q

using namespace b200q;

extern "C" {

// test/tuning hook: 0 = auto, 1 = split-across-warps kernel, 2 = streaming warp-per-item kernel
int b200q_decode_attn_set_variant(int v) {
  B200Q_CHECK_ARG(v >= 0 && v <= 3, "decode variant must be 0..3");
  g_decode_variant = v;
  bump_tuning_epoch();
  return B200Q_OK;
}

int b200q_decode_attn_ex(const void* q, int q_stride, void* out, const void* kv_layer,
                         const int32_t* block_table, int bt_stride, const int32_t* ctx_lens,
                         int n_seqs, int n_q, int n_kv, int D, int block_size, float scale,
                         float softcap, int window, void* stream) {
  B200Q_CHECK_ARG(n_seqs >= 0 && n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= 8,
                  "decode_attn: unsupported heads n_q=%d n_kv=%d (GQA group must be <= 8)", n_q,
                  n_kv);
  B200Q_CHECK_ARG(block_size == 16 && (D == 64 || D == 128 || D == 256),
                  "decode_attn: unsupported D=%d block_size=%d (D in {64,128,256}, block 16)", D,
                  block_size);
  if (n_seqs == 0) return B200Q_OK;
  cudaStream_t st = as_stream(stream);
  // streaming kernel once every resident warp (2 CTAs x 4 warps per SM) gets >= 2 items
  const bool stream_variant =
      g_decode_variant >= 2 ||
      (g_decode_variant == 0 && (long long)n_seqs * n_kv >= 16LL * attn_num_sms());
  if (stream_variant)
    B200Q_ATTN_DISPATCH(launch_decode_stream, q, q_stride, out, kv_layer, block_table, bt_stride,
                        ctx_lens, n_seqs, n_q, n_kv, scale, softcap, window, st);
  B200Q_ATTN_DISPATCH(launch_decode, q, q_stride, out, kv_layer, block_table, bt_stride, ctx_lens,
                      n_seqs, n_q, n_kv, scale, softcap, window, st);
}

int b200q_decode_attn(const void* q, int q_stride, void* out, const void* kv_layer,
                      const int32_t* block_table, int bt_stride, const int32_t* ctx_lens,
                      int n_seqs, int n_q, int n_kv, int D, int block_size, float scale,
                      void* stream) {
  return b200q_decode_attn_ex(q, q_stride, out, kv_layer, block_table, bt_stride, ctx_lens, n_seqs,
                              n_q, n_kv, D, block_size, scale, 0.f, 0, stream);
}

int b200q_prefill_attn_ex(const void* q, int q_stride, void* out, const void* kv_layer,
                          const int32_t* block_table, int bt_stride, const int32_t* tiles,
                          int n_tiles, int n_q, int n_kv, int D, int block_size, float scale,
                          float softcap, int window, void* stream) {
  B200Q_CHECK_ARG(n_tiles >= 0 && n_kv > 0 && n_q % n_kv == 0 && n_q / n_kv <= 8,
                  "prefill_attn: unsupported heads n_q=%d n_kv=%d (GQA group must be <= 8)", n_q,
                  n_kv);
  B200Q_CHECK_ARG(block_size == 16 && (D == 64 || D == 128 || D == 256),
                  "prefill_attn: unsupported D=%d block_size=%d (D in {64,128,256}, block 16)", D,
                  block_size);
  if (n_tiles == 0) return B200Q_OK;
  cudaStream_t st = as_stream(stream);
  B200Q_ATTN_DISPATCH(launch_prefill, q, q_stride, out, kv_layer, block_table, bt_stride, tiles,
                      n_tiles, n_q, n_kv, scale, softcap, window, st);
}

int b200q_prefill_attn(const void* q, int q_stride, void* out, const void* kv_layer,
                       const int32_t* block_table, int bt_stride, const int32_t* tiles,
                       int n_tiles, int n_q, int n_kv, int D, int block_size, float scale,
                       void* stream) {
  return b200q_prefill_attn_ex(q, q_stride, out, kv_layer, block_table, bt_stride, tiles, n_tiles,
                               n_q, n_kv, D, block_size, scale, 0.f, 0, stream);
}

}  // extern "C"
