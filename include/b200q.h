/*
 * b200q.h — C ABI of libb200q.so, the B200-native (sm_100a) generation engine that
 * replaces the arithmetic behind llmq's vLLM worker.
 *
 * The reference (iPieter/llmq) has NO FFI of its own: its hot path is four Python calls
 * into the un-vendored vLLM package (ref:llmq/workers/vllm_worker.py:105-123 engine
 * construction, :146 tokenizer, :161-165 sampling params, :183-186 engine.generate).
 * Each group of entry points below cites the reference call (or the vLLM op behind it)
 * that it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types in any signature.
 *   - every function returns 0 on success or a negative B200Q_E* code; it never throws.
 *     b200q_last_error() returns a thread-local message for the last failure.
 *   - all `dev` pointers are device pointers owned by the caller (PyTorch is only the
 *     allocator); the library owns TMA descriptors, small metadata buffers and its
 *     scheduler state.
 *   - no hidden synchronisation in the op-level and model-level calls: work is enqueued
 *     on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *     Engine-level calls (b200q_engine_step) synchronise their own stream because they
 *     hand token ids back to the host.
 *   - bf16 tensors are passed as `void*` / `const void*` (2-byte elements, row-major).
 */
#ifndef B200Q_H_
#define B200Q_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_VERSION 2

#define B200Q_OK 0
#define B200Q_EINVAL (-1)   /* bad argument / unsupported shape            */
#define B200Q_ECUDA (-2)    /* CUDA runtime / driver error                 */
#define B200Q_ENOMEM (-3)   /* KV pool or workspace exhausted              */
#define B200Q_ESTATE (-4)   /* call made in the wrong state (unbound etc.) */
#define B200Q_ENODEV (-5)   /* no sm_100 device present                    */

int b200q_version(void);
const char* b200q_last_error(void);
/* 0 if a CUDA device with compute capability 10.x is current, else B200Q_ENODEV. */
int b200q_device_check(void);

/* ------------------------------------------------------------------------------------
 * Op-level entry points (used by the parity tests and by b200q_model_forward).
 * Each replaces one op vLLM dispatches for the reference worker.
 * ---------------------------------------------------------------------------------- */

/* K1  embedding gather: out[t,:] = table[ids[t],:]            (vLLM VocabParallelEmbedding) */
int b200q_embed(const int32_t* ids_dev, const void* table_dev, void* out_dev,
                int T, int H, void* stream);
/* same, where ids[t] < 0 stands for prev_out_dev[-1 - ids[t]] — a token the previous step sampled
 * and left on the device (the engine enqueues step k+1 before reading step k's ids back, as
 * vLLM's async scheduling does: vllm/v1/worker/gpu_model_runner.py prev_sampled_token_ids);
 * scale > 0 and != 1 multiplies the row (Gemma normaliser), 0 = plain gather */
int b200q_embed_ex(const int32_t* ids_dev, const int32_t* prev_out_dev, const void* table_dev,
                   void* out_dev, int T, int H, float scale, void* stream);

/* K2  RMSNorm: y = bf16(bf16(x * rsqrt(mean(x^2)+eps)) * w)      (vllm/ir/ops/layernorm.py:9-21) */
int b200q_rmsnorm(const void* x_dev, const void* w_dev, void* y_dev,
                  int T, int H, float eps, void* stream);

/* K2' fused residual add + RMSNorm, in place:
 *     residual <- bf16(x + residual);  x <- rmsnorm(residual) * w
 *     (vllm/ir/ops/layernorm.py:39-58; rounding of the sum as in vLLM's _C kernel / HF) */
int b200q_add_rmsnorm(void* x_dev, void* residual_dev, const void* w_dev,
                      int T, int H, float eps, void* stream);

/* K4+K5  neox RoPE on q,k (in place on the fused qkv rows) and scatter of k,v into the
 *     paged KV cache.  qkv: [T, (n_q + 2*n_kv) * D]; cos_sin: bf16 [max_pos, D] (cos | sin);
 *     slot_mapping[t] = block * block_size + offset, or < 0 to skip the cache write.
 *     (vllm rotary_embedding/base.py:140-180 + _C_cache_ops.reshape_and_cache_flash) */
int b200q_rope_kvwrite(void* qkv_dev, const void* cos_sin_dev, const int32_t* positions_dev,
                       const int32_t* slot_mapping_dev, void* kv_layer_dev,
                       int T, int n_q, int n_kv, int D, int block_size, void* stream);

/* K7  paged decode attention (one query token per sequence, GQA).
 *     q rows are the first n_seqs rows of the fused qkv buffer (row stride q_stride elems);
 *     out: [n_seqs, n_q * D].  kv_layer: [num_blocks][2][n_kv][block_size][D] bf16, each
 *     token row stored with its 16-byte chunks XOR-swizzled by (token_in_block & 7).
 *     (replaces flashinfer trtllm_batch_decode_with_kv_cache, vllm flashinfer.py:1803) */
int b200q_decode_attn(const void* q_dev, int q_stride, void* out_dev, const void* kv_layer_dev,
                      const int32_t* block_table_dev, int bt_stride, const int32_t* ctx_lens_dev,
                      int n_seqs, int n_q, int n_kv, int D, int block_size, float scale,
                      void* stream);

/* K6  paged causal prefill attention over q-tiles of <=16 consecutive tokens of one sequence.
 *     tiles: int32[n_tiles][4] = {block_table_row, first_batch_row, n_rows, first_position}.
 *     (replaces flashinfer trtllm_batch_context_with_kv_cache, vllm flashinfer.py:1665) */
int b200q_prefill_attn(const void* q_dev, int q_stride, void* out_dev, const void* kv_layer_dev,
                       const int32_t* block_table_dev, int bt_stride, const int32_t* tiles_dev,
                       int n_tiles, int n_q, int n_kv, int D, int block_size, float scale,
                       void* stream);

/* K3/K8/K9/K11/K12  C[M,N] = A[M,K] * W[N,K]^T, bf16 in, fp32 accumulate in TMEM (tcgen05),
 *     bf16 out.  A, W, C row-major and contiguous; K % 64 == 0, N % 64 == 0.
 *     (replaces F.linear -> cuBLASLt, vllm/model_executor/layers/utils.py:92-98) */
int b200q_gemm_bf16(const void* A_dev, const void* W_dev, void* C_dev,
                    int M, int N, int K, void* stream);

/* Decode-sized batches: the split-K form of b200q_gemm_bf16 WITHOUT its reduce pass.  If the library
 *     would split the reduction for this shape, the fp32 partial tiles are left in its scratch
 *     (*partials = [splits][M][N], valid until the next split-K GEMM on the stream) and *splits > 1;
 *     the caller folds sum-in-split-order + bf16 rounding into the consumer (the two *_splitk ops
 *     below).  *splits == 1: nothing was launched, call b200q_gemm_bf16.  Results are bit-identical
 *     to the unfused sequence. */
int b200q_gemm_bf16_splitk(const void* A_dev, const void* W_dev, int M, int N, int K, void* stream,
                           const float** partials_dev, int* splits);
/* K2' reading x from split-K partials: residual <- bf16(bf16(sum_s partials[s]) + residual);
 *     x <- rmsnorm(residual) * w        (o-projection / down-projection -> next norm, one launch) */
int b200q_add_rmsnorm_splitk(void* x_dev, void* residual_dev, const void* w_dev,
                             const float* partials_dev, int splits, int T, int H, float eps,
                             void* stream);
/* K4+K5 reading the qkv rows from split-K partials (rotated q is written to qkv_dev) */
int b200q_rope_kvwrite_splitk(void* qkv_dev, const float* partials_dev, int splits,
                              const void* cos_sin_dev, const int32_t* positions_dev,
                              const int32_t* slot_mapping_dev, void* kv_layer_dev,
                              int T, int n_q, int n_kv, int D, int block_size, void* stream);

/* K9+K10 fused: out[M, N/2] = SwiGLU(A . W^T); W [N=2I, K] holds gate/up rows INTERLEAVED in blocks
 *     of 128: rows [256j, 256j+128) = gate rows [128j, 128j+128), rows [256j+128, 256j+256) = the
 *     matching up rows.  Rounding identical to b200q_gemm_bf16 followed by b200q_swiglu. */
int b200q_gemm_swiglu_bf16(const void* A_dev, const void* W_dev, void* C_dev,
                           int M, int N, int K, void* stream);

/* K10 SwiGLU: out[t,i] = bf16(bf16(silu(g[t,i])) * u[t,i]), gate_up = [T, 2I] = (g | u)
 *     (vllm activation.py:138-141 SiluAndMul.forward_native) */
int b200q_swiglu(const void* gate_up_dev, void* out_dev, int T, int I, void* stream);

/* row gather: out[i,:] = x[rows[i],:]  (last-token selection before the LM head) */
int b200q_gather_rows(const void* x_dev, const int32_t* rows_dev, void* out_dev,
                      int n, int H, void* stream);

/* K13 greedy sampler: ids[b] = argmax_v logits[b,v] over bf16 logits, lowest index wins ties
 *     (vllm/v1/sample/sampler.py:91,235-236 in the oracle's temperature-0 mode) */
int b200q_argmax_bf16(const void* logits_dev, int32_t* ids_dev, int B, int V, void* stream);

/* K13 sampler, the reference's default (temperature 0.7, no top-p/k;
 *     ref:llmq/workers/vllm_worker.py:161-165; vllm/v1/sample/ops/topk_topp_sampler.py:395-416:
 *     probs.div_(q).argmax, q ~ Exp(1)): ids[b] = argmax_v (logits[b,v]/T_b - log q_bv) with
 *     q_bv = -log(u), u from Philox4x32-10(key = seed_b, counter = (v/4, position_b, 0, 0)).
 *     params_dev: int32[B][4] = {float bits of temperature, seed lo, seed hi, position};
 *     temperature <= 0 selects the greedy argmax for that row. */
int b200q_sample_bf16(const void* logits_dev, const int32_t* params_dev, int32_t* ids_dev,
                      int B, int V, void* stream);

/* ---- Gemma-2 family (vllm/model_executor/models/gemma2.py) ----------------- */

/* embedding gather with the Gemma normaliser: out[t,:] = bf16(table[ids[t],:] * scale), scale =
 *     bf16(sqrt(H))                                                    (gemma2.py embed_tokens * normalizer) */
int b200q_embed_scaled(const int32_t* ids_dev, const void* table_dev, void* out_dev,
                       int T, int H, float scale, void* stream);

/* GemmaRMSNorm: y = bf16(x * rsqrt(mean(x^2)+eps) * (1 + w)), ONE rounding
 *     (vllm/model_executor/layers/layernorm.py GemmaRMSNorm.forward_static) */
int b200q_gemma_rmsnorm(const void* x_dev, const void* w_dev, void* y_dev,
                        int T, int H, float eps, void* stream);

/* Gemma-2 sandwich norms around the residual add, fused, in place:
 *     residual <- bf16(residual + gemma_rmsnorm(x, w_post));  x <- gemma_rmsnorm(residual, w_next)
 *     (post_attention_layernorm + pre_feedforward_layernorm, or post_feedforward_layernorm + the
 *     next layer's input_layernorm / the final norm; gemma2.py Gemma2DecoderLayer.forward) */
int b200q_gemma_norm_add_norm(void* x_dev, void* residual_dev, const void* w_post_dev,
                              const void* w_next_dev, int T, int H, float eps, void* stream);

/* K7/K6 with Gemma-2's attention extras: scores <- softcap * tanh(scores * scale / softcap)
 *     (softcap <= 0: off) and a sliding window: the query at position p sees keys j with
 *     p - window < j <= p (window <= 0: off).  D in {64, 128, 256}. */
int b200q_decode_attn_ex(const void* q_dev, int q_stride, void* out_dev, const void* kv_layer_dev,
                         const int32_t* block_table_dev, int bt_stride, const int32_t* ctx_lens_dev,
                         int n_seqs, int n_q, int n_kv, int D, int block_size, float scale,
                         float softcap, int window, void* stream);
int b200q_prefill_attn_ex(const void* q_dev, int q_stride, void* out_dev, const void* kv_layer_dev,
                          const int32_t* block_table_dev, int bt_stride, const int32_t* tiles_dev,
                          int n_tiles, int n_q, int n_kv, int D, int block_size, float scale,
                          float softcap, int window, void* stream);

/* fused gate_up GEMM + GeGLU: out = bf16(bf16(gelu_tanh(g)) * u), weight rows interleaved exactly as
 *     for b200q_gemm_swiglu_bf16                          (vllm activation.py GeluAndMul, approximate="tanh") */
int b200q_gemm_geglu_bf16(const void* A_dev, const void* W_dev, void* C_dev,
                          int M, int N, int K, void* stream);

/* final-logit soft-capping in place over n bf16 values, with the bf16 rounding of every eager op:
 *     l <- bf16(bf16(tanh(bf16(l / cap))) * cap)               (vllm logits_processor.py:66-69) */
int b200q_softcap_bf16(void* logits_dev, int64_t n, float cap, void* stream);

/* ------------------------------------------------------------------------------------
 * Model-level: one forward step over a mixed decode+prefill token batch.
 * Replaces GPUModelRunner.execute_model for LlamaForCausalLM
 * (vllm/model_executor/models/llama.py:316-333,395-431).
 * ---------------------------------------------------------------------------------- */

typedef struct b200q_model_config {
  int32_t hidden;        /* H                              */
  int32_t n_layers;      /* L                              */
  int32_t n_q_heads;
  int32_t n_kv_heads;
  int32_t head_dim;      /* 64, 128 or 256                 */
  int32_t intermediate;  /* I                              */
  int32_t vocab;         /* V                              */
  int32_t block_size;    /* KV page size in tokens (16)    */
  int32_t max_tokens;    /* workspace capacity: max tokens per step */
  int32_t max_seqs;      /* max sequences per step (block-table rows, sampled rows) */
  int32_t max_pos;       /* rows of the RoPE table         */
  int32_t tie_embeddings;/* 1: lm_head shares embed_tokens */
  float rms_eps;
  float attn_scale;      /* 1/sqrt(head_dim); gemma2: query_pre_attn_scalar^-0.5 */
  /* architecture extras; all zero = Llama */
  int32_t arch;           /* B200Q_ARCH_LLAMA | B200Q_ARCH_GEMMA2                                   */
  int32_t sliding_window; /* gemma2: even layers see only the last `sliding_window` keys; 0 = off   */
  float attn_softcap;     /* attention logit soft-capping (gemma2: 50); 0 = off                     */
  float final_softcap;    /* final logit soft-capping (gemma2: 30); 0 = off                         */
  float embed_scale;      /* embeddings are multiplied by this value (gemma2: bf16(sqrt(H))); 0 = off */
} b200q_model_config;

#define B200Q_ARCH_LLAMA 0
#define B200Q_ARCH_GEMMA2 1

typedef struct b200q_model* b200q_model_t;

/* device metadata for one step; every pointer is a device pointer to int32 data */
typedef struct b200q_batch {
  int32_t T;             /* total tokens this step; decode tokens occupy rows [0, n_dec)        */
  int32_t n_dec;         /* decode sequences (1 token each); block-table rows [0, n_dec)        */
  int32_t n_tiles;       /* prefill q-tiles                                                     */
  int32_t n_sample;      /* rows for which a next token is sampled                              */
  int32_t bt_stride;     /* int32 elements per block-table row                                  */
  const int32_t* token_ids;     /* [T]                                                          */
  const int32_t* positions;     /* [T]                                                          */
  const int32_t* slot_mapping;  /* [T]                                                          */
  const int32_t* block_table;   /* [rows][bt_stride]                                            */
  const int32_t* ctx_lens;      /* [n_dec] context length including the current token           */
  const int32_t* tiles;         /* [n_tiles][4]                                                 */
  const int32_t* sample_rows;   /* [n_sample] batch rows whose hidden state feeds the LM head   */
  int32_t* out_ids;             /* [n_sample] sampled token ids                                 */
  const int32_t* sample_params; /* [n_sample][4] {temperature bits, seed lo, seed hi, position};
                                   NULL = greedy argmax for every row                           */
  /* host-side bookkeeping for the profiler (algorithmic work of this step's attention) */
  int64_t sum_ctx_dec;          /* sum of ctx_lens over the decode sequences                    */
  int64_t prefill_flops_per_layer; /* causal QK^T + PV flops of the prefill tiles, one layer     */
  /* async stepping: token_ids[t] < 0 stands for prev_out_ids[-1 - token_ids[t]], a token the
   * previous step sampled and the host has not read back yet; NULL when no id is negative      */
  const int32_t* prev_out_ids;
} b200q_batch;

int b200q_model_create(const b200q_model_config* cfg, b200q_model_t* out);
int b200q_model_destroy(b200q_model_t m);
/* names: "embed", "final_norm", "lm_head", and per layer i: "layers.i.input_norm",
 * "layers.i.qkv" [(n_q+2n_kv)D, H], "layers.i.o" [H, n_q D], "layers.i.post_norm",
 * "layers.i.gate_up" [2I, H] (gate/up rows interleaved in 128-row blocks, see
 * b200q_gemm_swiglu_bf16), "layers.i.down" [H, I].  bf16, row-major, contiguous.
 * B200Q_ARCH_GEMMA2 replaces "post_norm" by the three norms "layers.i.post_attn_norm",
 * "layers.i.pre_ffn_norm", "layers.i.post_ffn_norm" (weights as stored: the kernels add the 1). */
int b200q_model_bind_weight(b200q_model_t m, const char* name, const void* dev_ptr,
                            int64_t rows, int64_t cols);
/* kv: [L][num_blocks][2][n_kv][block_size][D] bf16, zero-initialised by the caller */
int b200q_model_bind_kv(b200q_model_t m, void* dev_ptr, int64_t num_blocks);
/* cos_sin: bf16 [max_pos][D] */
int b200q_model_bind_rope(b200q_model_t m, const void* dev_ptr);
/* bytes needed for activations at cfg.max_tokens; caller allocates and binds */
int64_t b200q_model_workspace_bytes(const b200q_model_config* cfg);
int b200q_model_bind_workspace(b200q_model_t m, void* dev_ptr, int64_t bytes);
int b200q_model_forward(b200q_model_t m, const b200q_batch* batch, void* stream);
/* debugging / parity: copy the bf16 logits of the last forward ([n_sample, V]) location */
const void* b200q_model_logits_ptr(b200q_model_t m);
/* per-category device timing: CUDA events recorded on the forward's stream around every launch
 * (bench.py's live roofline numbers).  work = algorithmic flops (GEMM, prefill attention) or
 * bytes (decode attention, elementwise) summed over the timed launches. */
#define B200Q_PROF_GEMM 0
#define B200Q_PROF_DECODE_ATTN 1
#define B200Q_PROF_PREFILL_ATTN 2
#define B200Q_PROF_ELEMENTWISE 3
typedef struct b200q_profile {
  double ms[4];
  double work[4];
  int64_t launches[4];
} b200q_profile;
int b200q_model_set_profiling(b200q_model_t m, int on);
int b200q_model_profile_collect(b200q_model_t m, b200q_profile* out, int reset);
/* number of kernels launched by this library since load (gpu_launches evidence) */
int64_t b200q_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Engine-level: continuous-batching scheduler + paged-KV block manager + step loop.
 * Replaces AsyncLLMEngine.generate / EngineCore.step for the reference worker
 * (ref:llmq/workers/vllm_worker.py:183-186; vllm/v1/core/sched/scheduler.py:329-340
 * behaviour: running requests first, then waiting/prefill under a token budget, chunked
 * prefill, preempt-by-recompute when the KV pool is exhausted).
 * ---------------------------------------------------------------------------------- */

typedef struct b200q_engine_config {
  int32_t max_num_seqs;            /* VLLM_MAX_NUM_SEQS  (ref:llmq/core/config.py:27-32)  */
  int32_t max_num_batched_tokens;  /* token budget per step (<= model max_tokens)          */
  int32_t max_model_len;           /* VLLM_MAX_MODEL_LEN (ref:llmq/core/config.py:34-39)  */
  int32_t eos_token_id;            /* -1: none                                            */
  int32_t policy;                  /* 0: running requests first (vLLM order); 1: prefill first */
} b200q_engine_config;

typedef struct b200q_engine* b200q_engine_t;

#define B200Q_FLAG_FINISHED_EOS 1
#define B200Q_FLAG_FINISHED_LENGTH 2
#define B200Q_FLAG_FINISHED_ABORT 4

typedef struct b200q_engine_stats {
  int64_t steps;
  int64_t tokens_prefilled;
  int64_t tokens_decoded;
  int64_t preemptions;
  int32_t running;
  int32_t waiting;
  int32_t free_blocks;
  int32_t total_blocks;
  int32_t last_step_tokens;
  int32_t last_step_seqs;
  int64_t h2d_bytes;   /* per-step metadata copies (token ids, block tables, ...) */
  int64_t d2h_bytes;   /* sampled token ids read back                              */
} b200q_engine_stats;

int b200q_engine_create(b200q_model_t model, const b200q_engine_config* cfg, b200q_engine_t* out);
/* Scheduler self-test engine (host-logic tests on machines without a GPU): no model, no CUDA.
 * step() builds the per-step metadata exactly as in production, checks its invariants (budget,
 * unique KV slots, tile cover, positions) and fabricates each "sampled" token as
 * (previous token + 1) mod vocab.  It cannot run a model and is never used by the worker. */
int b200q_engine_create_dryrun(const b200q_engine_config* cfg, int32_t vocab, int32_t block_size,
                               int32_t num_blocks, b200q_engine_t* out);
int b200q_engine_destroy(b200q_engine_t e);
/* prompt ids are copied.  Returns B200Q_EINVAL if n_prompt + 1 > max_model_len (the worker
 * turns that into ValueError => job dropped, ref:llmq/workers/base.py:228-235). */
int b200q_engine_add_request(b200q_engine_t e, int64_t req_id, const int32_t* prompt_ids,
                             int32_t n_prompt, int32_t max_new_tokens, int32_t ignore_eos);
/* same, with the sampling the reference worker uses: temperature (0 = greedy) and a per-request
 * seed; token t of the request is drawn with Philox counter position t, so the result does not
 * depend on batching, chunking or preemption. */
int b200q_engine_add_request_sampled(b200q_engine_t e, int64_t req_id, const int32_t* prompt_ids,
                                     int32_t n_prompt, int32_t max_new_tokens, int32_t ignore_eos,
                                     float temperature, uint64_t seed);
/* every id in ids[0..n) ends a request (replaces the single cfg.eos_token_id): vLLM stops on
 * tokenizer.eos_token_id plus all eos ids of generation_config.json (vllm/sampling_params.py
 * update_from_generation_config), which the reference worker inherits (vllm_worker.py:161-165). */
int b200q_engine_set_stop_ids(b200q_engine_t e, const int32_t* ids, int32_t n);
int b200q_engine_abort(b200q_engine_t e, int64_t req_id);
/* Async stepping (default on; env B200Q_ASYNC=0 or on=0 turns it off): b200q_engine_step enqueues
 * step k+1 BEFORE it reads step k's sampled ids back, so the events a call returns are those of
 * the step enqueued by the PREVIOUS call, and b200q_engine_has_work stays 1 while a step is in
 * flight.  Off: every call returns the events of the step it ran (one sync per step).  Token
 * sequences are identical in both modes.  Only while no step is in flight (B200Q_ESTATE else). */
int b200q_engine_set_async(b200q_engine_t e, int32_t on);
/* 1 if any request is waiting or running */
int b200q_engine_has_work(b200q_engine_t e);
/* run one scheduler step + forward; writes up to cap (req_id, token, flags) events — those of this
 * step (sync mode) or of the previously enqueued step (async mode, see b200q_engine_set_async).
 * cap must be >= max_num_seqs. */
int b200q_engine_step(b200q_engine_t e, int64_t* out_req_ids, int32_t* out_tokens,
                      int32_t* out_flags, int32_t cap, int32_t* n_out);
int b200q_engine_get_stats(b200q_engine_t e, b200q_engine_stats* out);
/* the engine's CUDA stream (cudaStream_t) so callers can record their own events on it */
void* b200q_engine_stream(b200q_engine_t e);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H_ */
