// elementwise.cu — the HBM-bound, coalesced, 16-byte-vectorised kernels of the hot path:
// embedding gather (K1), RMSNorm / fused add+RMSNorm (K2), RoPE + paged KV write (K4+K5),
// SwiGLU (K10), row gather and greedy argmax (K13).  See SURVEY.md §2.1 for the vLLM ops
// these replace.  All rounding points follow the oracle (oracle/ops.py), which follows
// vLLM's native op definitions / HF transformers' Llama.
#include <stdarg.h>

#include "common.cuh"

namespace b200q {

thread_local char g_err[512] = {0};
static long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }
static int g_epoch = 0;
void bump_tuning_epoch() { __atomic_fetch_add(&g_epoch, 1, __ATOMIC_RELAXED); }
int tuning_epoch() { return __atomic_load_n(&g_epoch, __ATOMIC_RELAXED); }

// ------------------------------------------------------------------------------------------
// K1 embedding gather.  One warp per 512 B (32 lanes x 16 B); grid-stride over (token, chunk).
// ------------------------------------------------------------------------------------------
// A negative id is an indirection into `prev_out`, the sampled ids of the PREVIOUS forward step
// (id = prev_out[-1 - ids[t]]): the engine enqueues step k+1 before it has read step k's tokens
// back, so a decode row's input token is still on the device (engine.cu, async stepping).
__device__ __forceinline__ int resolve_token(const int32_t* __restrict__ ids,
                                             const int32_t* __restrict__ prev_out, int t) {
  int id = __ldg(ids + t);
  if (id < 0) id = prev_out[-1 - id];  // plain load: written by the previous step's sampler
  return id;
}

__global__ void __launch_bounds__(256) embed_kernel(const int32_t* __restrict__ ids,
                                                    const int32_t* __restrict__ prev_out,
                                                    const uint4* __restrict__ table,
                                                    uint4* __restrict__ out, int T, int chunks) {
  long long total = (long long)T * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i / chunks), c = (int)(i % chunks);
    int id = resolve_token(ids, prev_out, t);
    out[i] = ld_nc_v4(table + (long long)id * chunks + c);
  }
}

// Gemma: out = bf16(table[id] * scale) (the normaliser sqrt(H), itself rounded to bf16 by the host)
__global__ void __launch_bounds__(256) embed_scaled_kernel(const int32_t* __restrict__ ids,
                                                           const int32_t* __restrict__ prev_out,
                                                           const uint4* __restrict__ table,
                                                           uint4* __restrict__ out, int T, int chunks,
                                                           float scale) {
  long long total = (long long)T * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i / chunks), c = (int)(i % chunks);
    int id = resolve_token(ids, prev_out, t);
    uint4 a = ld_nc_v4(table + (long long)id * chunks + c);
    uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) ap[k] = pack_bf16x2(bf16_lo(ap[k]) * scale, bf16_hi(ap[k]) * scale);
    out[i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// K2 RMSNorm / fused add + RMSNorm.  One 256-thread CTA per token row; the row lives in
// registers between the reduction and the scale (one HBM read, one write per tensor).
// ------------------------------------------------------------------------------------------
constexpr int NORM_THREADS = 256;
constexpr int NORM_MAX_CHUNKS = 4;  // H <= 256*4*8 = 8192

// 8 consecutive outputs of a split-K GEMM: the fp32 partials ws[s][row][col] added in split order
// (the same order and rounding as the stand-alone reduction) and rounded to bf16
__device__ __forceinline__ uint4 sum_partials_bf16(const float* __restrict__ ws, long long off,
                                                   long long mn, int splits) {
  float4 a = __ldg(reinterpret_cast<const float4*>(ws + off));
  float4 b = __ldg(reinterpret_cast<const float4*>(ws + off + 4));
  for (int sp = 1; sp < splits; ++sp) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(ws + sp * mn + off));
    const float4 d = __ldg(reinterpret_cast<const float4*>(ws + sp * mn + off + 4));
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
  }
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y);
  o.y = pack_bf16x2(a.z, a.w);
  o.z = pack_bf16x2(b.x, b.y);
  o.w = pack_bf16x2(b.z, b.w);
  return o;
}

// kPartials: x is not read from x_in but reduced on the fly from the fp32 partials of a split-K GEMM
// (decode-sized batches): the projection's reduce pass, the residual add and the norm become ONE
// launch instead of two, and the bf16 GEMM output never makes a round trip through memory.
template <bool kAdd, bool kPartials = false>
__global__ void __launch_bounds__(NORM_THREADS)
    rmsnorm_kernel(const uint4* __restrict__ x_in, uint4* __restrict__ residual,
                   const uint4* __restrict__ w, uint4* __restrict__ y, int chunks, float inv_h,
                   float eps, const float* __restrict__ ws = nullptr, int splits = 0,
                   long long mn = 0) {
  const long long row = blockIdx.x;
  const uint4* xr = x_in + row * chunks;
  uint4 v[NORM_MAX_CHUNKS];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NORM_MAX_CHUNKS; ++j) {
    int c = threadIdx.x + j * NORM_THREADS;
    if (c < chunks) {
      uint4 a;
      if constexpr (kPartials) a = sum_partials_bf16(ws, (row * chunks + c) * 8, mn, splits);
      else a = ld_v4(xr + c);
      if (kAdd) {
        uint4 r = ld_v4(residual + row * chunks + c);
        uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float lo = bf16_lo(ap[k]) + bf16_lo(rp[k]);
          float hi = bf16_hi(ap[k]) + bf16_hi(rp[k]);
          ap[k] = pack_bf16x2(lo, hi);  // residual stream is kept in bf16 (rounded sum)
        }
        st_v4(residual + row * chunks + c, a);
      }
      v[j] = a;
      const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo = bf16_lo(ap[k]), hi = bf16_hi(ap[k]);
        ss += lo * lo + hi * hi;
      }
    }
  }
  __shared__ float red[NORM_THREADS / 32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_THREADS / 32; ++i) tot += red[i];
  const float inv = __frsqrt_rn(tot * inv_h + eps);
#pragma unroll
  for (int j = 0; j < NORM_MAX_CHUNKS; ++j) {
    int c = threadIdx.x + j * NORM_THREADS;
    if (c < chunks) {
      uint4 a = v[j];
      uint4 ww = __ldg(w + c);
      uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(&ww);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // y = bf16( bf16(x * inv) * w )  — the normalised value is rounded to the weight
        // dtype before the weight multiply (vllm/ir/ops/layernorm.py:19-20)
        float lo = round_bf16(bf16_lo(ap[k]) * inv) * bf16_lo(wp[k]);
        float hi = round_bf16(bf16_hi(ap[k]) * inv) * bf16_hi(wp[k]);
        ap[k] = pack_bf16x2(lo, hi);
      }
      st_v4(y + row * chunks + c, a);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Gemma-2 norms.  GemmaRMSNorm multiplies by (1 + w) and rounds ONCE:
//   y = bf16( (x * rsqrt(mean(x^2) + eps)) * (1 + w) )      all in fp32 until the final cast.
// gemma_norm_add_norm fuses the sandwich around a residual add (two row reductions, the row
// stays in registers):  a = norm(x, w_post);  residual = bf16(residual + a);  x = norm(residual, w_next)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();  // red may still be read by the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_THREADS / 32; ++i) tot += red[i];
  return tot;
}

__device__ __forceinline__ float chunk_sumsq(const uint4& a) {
  const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo = bf16_lo(ap[k]), hi = bf16_hi(ap[k]);
    ss += lo * lo + hi * hi;
  }
  return ss;
}

// a <- bf16((a * inv) * (1 + w))
__device__ __forceinline__ void gemma_scale_chunk(uint4& a, const uint4& ww, float inv) {
  uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(&ww);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo = (bf16_lo(ap[k]) * inv) * (1.f + bf16_lo(wp[k]));
    float hi = (bf16_hi(ap[k]) * inv) * (1.f + bf16_hi(wp[k]));
    ap[k] = pack_bf16x2(lo, hi);
  }
}

template <bool kSandwich>
__global__ void __launch_bounds__(NORM_THREADS)
    gemma_norm_kernel(const uint4* __restrict__ x_in, uint4* __restrict__ residual,
                      const uint4* __restrict__ w_post, const uint4* __restrict__ w_next,
                      uint4* __restrict__ y, int chunks, float inv_h, float eps) {
  __shared__ float red[NORM_THREADS / 32];
  const long long row = blockIdx.x;
  const uint4* xr = x_in + row * chunks;
  uint4 v[NORM_MAX_CHUNKS];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NORM_MAX_CHUNKS; ++j) {
    int c = threadIdx.x + j * NORM_THREADS;
    if (c < chunks) {
      v[j] = ld_v4(xr + c);
      ss += chunk_sumsq(v[j]);
    }
  }
  float inv = __frsqrt_rn(block_sum_256(ss, red) * inv_h + eps);
  if (kSandwich) {
    // a = norm(x, w_post); residual <- bf16(residual + a); then the second norm over the new residual
    ss = 0.f;
#pragma unroll
    for (int j = 0; j < NORM_MAX_CHUNKS; ++j) {
      int c = threadIdx.x + j * NORM_THREADS;
      if (c < chunks) {
        uint4 a = v[j];
        gemma_scale_chunk(a, __ldg(w_post + c), inv);
        uint4 r = ld_v4(residual + row * chunks + c);
        uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ap[k] = pack_bf16x2(bf16_lo(rp[k]) + bf16_lo(ap[k]), bf16_hi(rp[k]) + bf16_hi(ap[k]));
        st_v4(residual + row * chunks + c, a);
        v[j] = a;
        ss += chunk_sumsq(a);
      }
    }
    inv = __frsqrt_rn(block_sum_256(ss, red) * inv_h + eps);
  }
  const uint4* w_out = kSandwich ? w_next : w_post;
#pragma unroll
  for (int j = 0; j < NORM_MAX_CHUNKS; ++j) {
    int c = threadIdx.x + j * NORM_THREADS;
    if (c < chunks) {
      uint4 a = v[j];
      gemma_scale_chunk(a, __ldg(w_out + c), inv);
      st_v4(y + row * chunks + c, a);
    }
  }
}

// final-logit soft-capping with every eager bf16 rounding: l <- bf16(bf16(tanh(bf16(l / cap))) * cap)
__global__ void __launch_bounds__(256)
    softcap_kernel(uint4* __restrict__ logits, long long n_chunks, float cap) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    uint4 a = ld_v4(logits + i);
    uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float lo = round_bf16(tanhf(round_bf16(__fdiv_rn(bf16_lo(ap[k]), cap)))) * cap;
      float hi = round_bf16(tanhf(round_bf16(__fdiv_rn(bf16_hi(ap[k]), cap)))) * cap;
      ap[k] = pack_bf16x2(lo, hi);
    }
    st_v4(logits + i, a);
  }
}

// ------------------------------------------------------------------------------------------
// K4+K5 RoPE (neox pairs i, i+D/2) on q and k, and the paged KV write.
// One CTA per token.  Work items: q heads x (D/16) chunk-pairs (in place), k heads x (D/16)
// chunk-pairs (rotated, written to the cache only), v heads x (D/8) chunks (copied to the cache).
// Cache rows are [token][D] with 16 B chunk c stored at chunk (c ^ (token_in_block & 7)) so
// that the attention kernels' ldmatrix reads of TMA-copied pages are bank-conflict free.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rope_chunk_pair(uint4& x1, uint4& x2, const uint4& cs,
                                                const uint4& sn) {
  uint32_t* a = reinterpret_cast<uint32_t*>(&x1);
  uint32_t* b = reinterpret_cast<uint32_t*>(&x2);
  const uint32_t* c = reinterpret_cast<const uint32_t*>(&cs);
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&sn);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // bf16 arithmetic with a rounding after every op, as HF / vLLM eager do on bf16 tensors:
    // o1 = x1*cos - x2*sin ; o2 = x2*cos + x1*sin
    float x1l = bf16_lo(a[k]), x1h = bf16_hi(a[k]);
    float x2l = bf16_lo(b[k]), x2h = bf16_hi(b[k]);
    float cl = bf16_lo(c[k]), ch = bf16_hi(c[k]);
    float sl = bf16_lo(s[k]), sh = bf16_hi(s[k]);
    float o1l = round_bf16(x1l * cl) - round_bf16(x2l * sl);
    float o1h = round_bf16(x1h * ch) - round_bf16(x2h * sh);
    float o2l = round_bf16(x2l * cl) + round_bf16(x1l * sl);
    float o2h = round_bf16(x2h * ch) + round_bf16(x1h * sh);
    a[k] = pack_bf16x2(o1l, o1h);
    b[k] = pack_bf16x2(o2l, o2h);
  }
}

// kPartials: the qkv row is reduced on the fly from the fp32 partials of the split-K qkv projection
// (decode-sized batches) instead of being read back as bf16 — reduce + RoPE + KV write in one launch
template <bool kPartials>
__global__ void __launch_bounds__(256)
    rope_kvwrite_kernel(uint4* __restrict__ qkv, const uint4* __restrict__ cos_sin,
                        const int32_t* __restrict__ positions,
                        const int32_t* __restrict__ slot_mapping, uint4* __restrict__ kv_layer,
                        int n_q, int n_kv, int D, int block_size, const float* __restrict__ ws,
                        int splits, long long mn) {
  const int t = blockIdx.x;
  const int cpr = D / 8;        // 16 B chunks per head row
  const int half = cpr / 2;     // chunk pairs per head
  const int row_chunks = (n_q + 2 * n_kv) * cpr;
  uint4* row = qkv + (long long)t * row_chunks;
  const int pos = __ldg(positions + t);
  const int slot = __ldg(slot_mapping + t);
  const uint4* cs_row = cos_sin + (long long)pos * cpr;  // [cos (D/2) | sin (D/2)]
  const int blk = slot >= 0 ? slot / block_size : 0;
  const int off = slot >= 0 ? slot % block_size : 0;
  const int n_rope_items = (n_q + n_kv) * half;
  const int n_items = n_rope_items + n_kv * cpr;
  for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
    if (it < n_rope_items) {
      int head = it / half, c = it % half;
      uint4 x1, x2;
      if constexpr (kPartials) {
        const long long e = ((long long)t * row_chunks + head * cpr + c) * 8;
        x1 = sum_partials_bf16(ws, e, mn, splits);
        x2 = sum_partials_bf16(ws, e + (long long)half * 8, mn, splits);
      } else {
        x1 = ld_v4(row + head * cpr + c);
        x2 = ld_v4(row + head * cpr + c + half);
      }
      uint4 cs = __ldg(cs_row + c);
      uint4 sn = __ldg(cs_row + half + c);
      rope_chunk_pair(x1, x2, cs, sn);
      if (head < n_q) {
        st_v4(row + head * cpr + c, x1);
        st_v4(row + head * cpr + c + half, x2);
      } else if (slot >= 0) {
        int h = head - n_q;
        uint4* dst =
            kv_layer + ((((long long)blk * 2 + 0) * n_kv + h) * block_size + off) * cpr;
        int sw = off & 7;
        st_v4(dst + (c ^ sw), x1);
        st_v4(dst + ((c + half) ^ sw), x2);
      }
    } else if (slot >= 0) {
      int j = it - n_rope_items;
      int h = j / cpr, c = j % cpr;
      uint4 v;
      if constexpr (kPartials)
        v = sum_partials_bf16(ws, ((long long)t * row_chunks + (n_q + n_kv + h) * cpr + c) * 8, mn, splits);
      else
        v = ld_v4(row + (n_q + n_kv + h) * cpr + c);
      uint4* dst = kv_layer + ((((long long)blk * 2 + 1) * n_kv + h) * block_size + off) * cpr;
      st_v4(dst + (c ^ (off & 7)), v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K10 SwiGLU.  grid-stride over 16 B chunks of the output.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    swiglu_kernel(const uint4* __restrict__ gate_up, uint4* __restrict__ out, int T, int ichunks) {
  long long total = (long long)T * ichunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i / ichunks;
    int c = (int)(i % ichunks);
    uint4 g = ld_nc_v4(gate_up + t * 2 * ichunks + c);
    uint4 u = ld_nc_v4(gate_up + t * 2 * ichunks + ichunks + c);
    uint32_t* gp = reinterpret_cast<uint32_t*>(&g);
    const uint32_t* up = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gl = bf16_lo(gp[k]), gh = bf16_hi(gp[k]);
      // silu rounded to bf16, then the product rounded to bf16 (two bf16 ops, as eager torch)
      float sl = round_bf16(gl / (1.f + __expf(-gl)));
      float sh = round_bf16(gh / (1.f + __expf(-gh)));
      gp[k] = pack_bf16x2(sl * bf16_lo(up[k]), sh * bf16_hi(up[k]));
    }
    out[i] = g;
  }
}

__global__ void __launch_bounds__(256)
    gather_rows_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ rows,
                       uint4* __restrict__ out, int n, int chunks) {
  long long total = (long long)n * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / chunks), c = (int)(i % chunks);
    out[i] = ld_v4(x + (long long)__ldg(rows + r) * chunks + c);
  }
}

// ------------------------------------------------------------------------------------------
// K13 greedy argmax over bf16 logits; lowest index wins ties (torch.argmax behaviour).
// One 1024-thread CTA per row.  V need not be a multiple of 8 (tail handled scalar).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void amax_update(float v, int i, float& best, int& bi) {
  if (v > best || (v == best && i < bi)) {
    best = v;
    bi = i;
  }
}

__global__ void __launch_bounds__(1024)
    argmax_kernel(const bf16* __restrict__ logits, int32_t* __restrict__ ids, int V) {
  const bf16* row = logits + (long long)blockIdx.x * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const int nvec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? V / 8 : 0;
  const uint4* rv = reinterpret_cast<const uint4*>(row);
  for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
    uint4 a = ld_nc_v4(rv + c);
    const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      amax_update(bf16_lo(ap[k]), c * 8 + 2 * k, best, bi);
      amax_update(bf16_hi(ap[k]), c * 8 + 2 * k + 1, best, bi);
    }
  }
  for (int i = nvec * 8 + threadIdx.x; i < V; i += blockDim.x)
    amax_update(__bfloat162float(row[i]), i, best, bi);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    amax_update(ov, oi, best, bi);
  }
  __shared__ float sv[32];
  __shared__ int si[32];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      amax_update(ov, oi, best, bi);
    }
    if (threadIdx.x == 0) ids[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

// ------------------------------------------------------------------------------------------
// K13 sampler at temperature > 0: the reference's default path (SamplingParams(temperature=0.7),
// ref:llmq/workers/vllm_worker.py:161-165; vLLM: probs = softmax(logits/T); probs.div_(q).argmax
// with q ~ Exp(1), vllm/v1/sample/ops/topk_topp_sampler.py:395-416).  argmax_i p_i/q_i ==
// argmax_i (logit_i/T - log q_i), so no softmax pass is needed: one read of the bf16 logits row.
// q_i = -log(u_i), u_i from Philox4x32-10 keyed by the request's seed with counter
// (element index / 4, token position) — counter-based, so a request's samples do not depend on
// its batch-mates.  temperature == 0 rows fall back to the greedy argmax (lowest index on ties).
// params: int32[B][4] = {float bits of temperature, seed lo, seed hi, position}.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__device__ __forceinline__ float gumbel_from_bits(uint32_t x) {
  const float u = ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1), 24 bits
  return -logf(-logf(u));
}

__global__ void __launch_bounds__(1024)
    sample_kernel(const bf16* __restrict__ logits, const int32_t* __restrict__ params,
                  int32_t* __restrict__ ids, int V) {
  const bf16* row = logits + (long long)blockIdx.x * V;
  const float temp = __int_as_float(params[4 * blockIdx.x + 0]);
  const uint32_t k0 = (uint32_t)params[4 * blockIdx.x + 1], k1 = (uint32_t)params[4 * blockIdx.x + 2];
  const uint32_t pos = (uint32_t)params[4 * blockIdx.x + 3];
  const bool greedy = !(temp > 0.f);
  const float inv_t = greedy ? 1.f : 1.f / temp;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  // groups of 4 consecutive elements share one Philox block: counter = (group, position, 0, 0)
  const int n4 = V / 4;
  for (int g = threadIdx.x; g < n4; g += blockDim.x) {
    const uint2 a = *reinterpret_cast<const uint2*>(row + 4 * g);  // 4 bf16 (V*2 B rows are 8 B aligned)
    float v[4] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y)};
    if (!greedy) {
      uint32_t r[4];
      philox4x32_10((uint32_t)g, pos, 0u, 0u, k0, k1, r);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = v[k] * inv_t + gumbel_from_bits(r[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) amax_update(v[k], 4 * g + k, best, bi);
  }
  for (int i = n4 * 4 + threadIdx.x; i < V; i += blockDim.x) {  // tail (V % 4 elements)
    float v = __bfloat162float(row[i]);
    if (!greedy) {
      uint32_t r[4];
      philox4x32_10((uint32_t)(i / 4), pos, 0u, 0u, k0, k1, r);
      v = v * inv_t + gumbel_from_bits(r[i & 3]);
    }
    amax_update(v, i, best, bi);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    amax_update(ov, oi, best, bi);
  }
  __shared__ float sv[32];
  __shared__ int si[32];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      amax_update(ov, oi, best, bi);
    }
    if (threadIdx.x == 0) ids[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

static inline int grid_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  const long long cap = 148LL * 16;  // 16 resident 256-thread CTAs per SM
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_version(void) { return B200Q_VERSION; }
const char* b200q_last_error(void) { return g_err; }
int64_t b200q_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int b200q_device_check(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available: libb200q has no CPU fallback");
    return B200Q_ENODEV;
  }
  if (p.major != 10) {
    set_error("device %s is sm_%d%d; libb200q is built for sm_100a only", p.name, p.major,
              p.minor);
    return B200Q_ENODEV;
  }
  return B200Q_OK;
}

int b200q_embed_ex(const int32_t* ids, const int32_t* prev_out, const void* table, void* out, int T,
                   int H, float scale, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && H > 0 && H % 8 == 0, "embed: bad shape T=%d H=%d", T, H);
  if (T == 0) return B200Q_OK;
  int chunks = H / 8;
  if (scale > 0.f && scale != 1.f)
    embed_scaled_kernel<<<grid_for((long long)T * chunks, 256), 256, 0, as_stream(stream)>>>(
        ids, prev_out, (const uint4*)table, (uint4*)out, T, chunks, scale);
  else
    embed_kernel<<<grid_for((long long)T * chunks, 256), 256, 0, as_stream(stream)>>>(
        ids, prev_out, (const uint4*)table, (uint4*)out, T, chunks);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_embed(const int32_t* ids, const void* table, void* out, int T, int H, void* stream) {
  return b200q_embed_ex(ids, nullptr, table, out, T, H, 0.f, stream);
}

int b200q_embed_scaled(const int32_t* ids, const void* table, void* out, int T, int H, float scale,
                       void* stream) {
  B200Q_CHECK_ARG(scale > 0.f, "embed_scaled: scale must be positive");
  if (scale == 1.f) return b200q_embed_ex(ids, nullptr, table, out, T, H, 0.f, stream);
  return b200q_embed_ex(ids, nullptr, table, out, T, H, scale, stream);
}

int b200q_gemma_rmsnorm(const void* x, const void* w, void* y, int T, int H, float eps,
                        void* stream) {
  B200Q_CHECK_ARG(T >= 0 && H > 0 && H % 8 == 0 && H <= NORM_THREADS * NORM_MAX_CHUNKS * 8,
                  "gemma_rmsnorm: bad shape T=%d H=%d", T, H);
  if (T == 0) return B200Q_OK;
  gemma_norm_kernel<false><<<T, NORM_THREADS, 0, as_stream(stream)>>>(
      (const uint4*)x, nullptr, (const uint4*)w, nullptr, (uint4*)y, H / 8, 1.f / (float)H, eps);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_gemma_norm_add_norm(void* x, void* residual, const void* w_post, const void* w_next, int T,
                              int H, float eps, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && H > 0 && H % 8 == 0 && H <= NORM_THREADS * NORM_MAX_CHUNKS * 8,
                  "gemma_norm_add_norm: bad shape T=%d H=%d", T, H);
  B200Q_CHECK_ARG(x && residual && w_post && w_next, "gemma_norm_add_norm: null argument");
  if (T == 0) return B200Q_OK;
  gemma_norm_kernel<true><<<T, NORM_THREADS, 0, as_stream(stream)>>>(
      (const uint4*)x, (uint4*)residual, (const uint4*)w_post, (const uint4*)w_next, (uint4*)x, H / 8,
      1.f / (float)H, eps);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_softcap_bf16(void* logits, int64_t n, float cap, void* stream) {
  B200Q_CHECK_ARG(n >= 0 && n % 8 == 0 && cap > 0.f, "softcap: bad arguments n=%lld cap=%f",
                  (long long)n, (double)cap);
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "softcap: logits not 16-byte aligned");
  if (n == 0) return B200Q_OK;
  softcap_kernel<<<grid_for(n / 8, 256), 256, 0, as_stream(stream)>>>((uint4*)logits, n / 8, cap);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_rmsnorm(const void* x, const void* w, void* y, int T, int H, float eps, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && H > 0 && H % 8 == 0 && H <= NORM_THREADS * NORM_MAX_CHUNKS * 8,
                  "rmsnorm: bad shape T=%d H=%d", T, H);
  if (T == 0) return B200Q_OK;
  rmsnorm_kernel<false><<<T, NORM_THREADS, 0, as_stream(stream)>>>(
      (const uint4*)x, nullptr, (const uint4*)w, (uint4*)y, H / 8, 1.f / (float)H, eps);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_add_rmsnorm(void* x, void* residual, const void* w, int T, int H, float eps,
                      void* stream) {
  B200Q_CHECK_ARG(T >= 0 && H > 0 && H % 8 == 0 && H <= NORM_THREADS * NORM_MAX_CHUNKS * 8,
                  "add_rmsnorm: bad shape T=%d H=%d", T, H);
  if (T == 0) return B200Q_OK;
  rmsnorm_kernel<true><<<T, NORM_THREADS, 0, as_stream(stream)>>>(
      (const uint4*)x, (uint4*)residual, (const uint4*)w, (uint4*)x, H / 8, 1.f / (float)H, eps);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_add_rmsnorm_splitk(void* x, void* residual, const void* w, const float* partials,
                             int splits, int T, int H, float eps, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && H > 0 && H % 8 == 0 && H <= NORM_THREADS * NORM_MAX_CHUNKS * 8,
                  "add_rmsnorm_splitk: bad shape T=%d H=%d", T, H);
  B200Q_CHECK_ARG(partials && splits >= 1 && splits <= 16, "add_rmsnorm_splitk: bad partials (splits=%d)", splits);
  if (T == 0) return B200Q_OK;
  rmsnorm_kernel<true, true><<<T, NORM_THREADS, 0, as_stream(stream)>>>(
      nullptr, (uint4*)residual, (const uint4*)w, (uint4*)x, H / 8, 1.f / (float)H, eps, partials, splits,
      (long long)T * H);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_rope_kvwrite(void* qkv, const void* cos_sin, const int32_t* positions,
                       const int32_t* slot_mapping, void* kv_layer, int T, int n_q, int n_kv,
                       int D, int block_size, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && (D == 64 || D == 128 || D == 256) && n_q > 0 && n_kv > 0 && block_size > 0,
                  "rope_kvwrite: bad shape T=%d n_q=%d n_kv=%d D=%d bs=%d", T, n_q, n_kv, D,
                  block_size);
  if (T == 0) return B200Q_OK;
  rope_kvwrite_kernel<false><<<T, 256, 0, as_stream(stream)>>>(
      (uint4*)qkv, (const uint4*)cos_sin, positions, slot_mapping, (uint4*)kv_layer, n_q, n_kv, D, block_size,
      nullptr, 0, 0);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_rope_kvwrite_splitk(void* qkv, const float* partials, int splits, const void* cos_sin,
                              const int32_t* positions, const int32_t* slot_mapping, void* kv_layer,
                              int T, int n_q, int n_kv, int D, int block_size, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && (D == 64 || D == 128 || D == 256) && n_q > 0 && n_kv > 0 && block_size > 0,
                  "rope_kvwrite_splitk: bad shape T=%d n_q=%d n_kv=%d D=%d bs=%d", T, n_q, n_kv, D,
                  block_size);
  B200Q_CHECK_ARG(partials && splits >= 1 && splits <= 16, "rope_kvwrite_splitk: bad partials (splits=%d)", splits);
  if (T == 0) return B200Q_OK;
  rope_kvwrite_kernel<true><<<T, 256, 0, as_stream(stream)>>>(
      (uint4*)qkv, (const uint4*)cos_sin, positions, slot_mapping, (uint4*)kv_layer, n_q, n_kv, D, block_size,
      partials, splits, (long long)T * (n_q + 2 * n_kv) * D);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_swiglu(const void* gate_up, void* out, int T, int I, void* stream) {
  B200Q_CHECK_ARG(T >= 0 && I > 0 && I % 8 == 0, "swiglu: bad shape T=%d I=%d", T, I);
  if (T == 0) return B200Q_OK;
  int ic = I / 8;
  swiglu_kernel<<<grid_for((long long)T * ic, 256), 256, 0, as_stream(stream)>>>(
      (const uint4*)gate_up, (uint4*)out, T, ic);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_gather_rows(const void* x, const int32_t* rows, void* out, int n, int H, void* stream) {
  B200Q_CHECK_ARG(n >= 0 && H > 0 && H % 8 == 0, "gather_rows: bad shape n=%d H=%d", n, H);
  if (n == 0) return B200Q_OK;
  int chunks = H / 8;
  gather_rows_kernel<<<grid_for((long long)n * chunks, 256), 256, 0, as_stream(stream)>>>(
      (const uint4*)x, rows, (uint4*)out, n, chunks);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_argmax_bf16(const void* logits, int32_t* ids, int B, int V, void* stream) {
  B200Q_CHECK_ARG(B >= 0 && V > 0, "argmax: bad shape B=%d V=%d", B, V);
  if (B == 0) return B200Q_OK;
  argmax_kernel<<<B, 1024, 0, as_stream(stream)>>>((const bf16*)logits, ids, V);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

int b200q_sample_bf16(const void* logits, const int32_t* params, int32_t* ids, int B, int V,
                      void* stream) {
  B200Q_CHECK_ARG(B >= 0 && V > 0 && V % 4 == 0, "sample: bad shape B=%d V=%d (V %% 4 == 0)", B, V);
  B200Q_CHECK_ARG(params != nullptr, "sample: params is null");
  if (B == 0) return B200Q_OK;
  sample_kernel<<<B, 1024, 0, as_stream(stream)>>>((const bf16*)logits, params, ids, V);
  B200Q_LAUNCH_CHECK();
  return B200Q_OK;
}

}  // extern "C"
