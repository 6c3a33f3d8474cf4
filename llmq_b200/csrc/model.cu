// model.cu — one forward step of a Llama-architecture decoder over a mixed decode+prefill
// token batch, composed from the op-level kernels.  Replaces GPUModelRunner.execute_model for
// LlamaForCausalLM behind the reference worker (vllm/model_executor/models/llama.py:316-333,
// 395-431): same layer wiring and rounding points, B200-native kernels.
// B200Q_ARCH_GEMMA2 (vllm/model_executor/models/gemma2.py): same skeleton with
// scaled embeddings, (1+w) sandwich norms around both residual adds, soft-capped attention with
// alternating sliding-window layers, GeGLU, tied LM head and final-logit soft-capping.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

using namespace b200q;

struct b200q_layer {
  const void* input_norm = nullptr;
  const void* qkv = nullptr;
  const void* o = nullptr;
  const void* post_norm = nullptr;  // llama: post_attention_layernorm (= the pre-MLP norm)
  const void* gate_up = nullptr;
  const void* down = nullptr;
  // gemma2: the sandwich norms
  const void* post_attn_norm = nullptr;
  const void* pre_ffn_norm = nullptr;
  const void* post_ffn_norm = nullptr;
};

struct b200q_model {
  b200q_model_config cfg;
  const void* embed = nullptr;
  const void* final_norm = nullptr;
  const void* lm_head = nullptr;
  std::vector<b200q_layer> layers;
  uint8_t* kv = nullptr;
  int64_t num_blocks = 0;
  const void* rope = nullptr;
  // workspace views
  uint8_t* ws = nullptr;
  int64_t ws_bytes = 0;
  // (no [T, 2I] gate_up buffer: the SwiGLU is applied in the gate_up GEMM's epilogue)
  bf16 *x = nullptr, *residual = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr,
       *sel = nullptr, *logits = nullptr;
  // fold the split-K reduce of decode-sized projections into the consuming kernel (B200Q_FUSE_SPLITK=0: off)
  bool fuse_splitk = true;
  // optional per-category device timing (CUDA events on the forward's stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int cat; double work; int e0, e1; };
  std::vector<Span> spans;
  int ev_used = 0;
  b200q_profile prof{};
  int get_event() {
    if (ev_used == (int)ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    return ev_used++;
  }
};

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

static int64_t qkv_dim(const b200q_model_config& c) {
  return (int64_t)(c.n_q_heads + 2 * c.n_kv_heads) * c.head_dim;
}

static int check_cfg(const b200q_model_config* c) {
  B200Q_CHECK_ARG(c != nullptr, "model config is null");
  B200Q_CHECK_ARG(c->hidden > 0 && c->hidden % 64 == 0 && c->hidden <= 8192,
                  "hidden=%d unsupported (multiple of 64, <= 8192)", c->hidden);
  B200Q_CHECK_ARG(c->arch == B200Q_ARCH_LLAMA || c->arch == B200Q_ARCH_GEMMA2, "arch=%d unknown", c->arch);
  B200Q_CHECK_ARG(c->head_dim == 64 || c->head_dim == 128 || c->head_dim == 256,
                  "head_dim=%d unsupported (64, 128, 256)", c->head_dim);
  B200Q_CHECK_ARG(c->n_kv_heads > 0 && c->n_q_heads % c->n_kv_heads == 0 &&
                      c->n_q_heads / c->n_kv_heads <= (c->head_dim == 256 ? 4 : 8),
                  "heads n_q=%d n_kv=%d unsupported (GQA group <= 8; <= 4 at head_dim 256)",
                  c->n_q_heads, c->n_kv_heads);
  B200Q_CHECK_ARG(c->sliding_window >= 0 && c->attn_softcap >= 0.f && c->final_softcap >= 0.f &&
                      c->embed_scale >= 0.f,
                  "negative sliding_window / softcap / embed_scale");
  if (c->arch == B200Q_ARCH_LLAMA)
    B200Q_CHECK_ARG(c->sliding_window == 0 && c->attn_softcap == 0.f && c->final_softcap == 0.f &&
                        (c->embed_scale == 0.f || c->embed_scale == 1.f),
                    "llama arch takes no sliding window / soft-capping / embedding scale");
  B200Q_CHECK_ARG(c->intermediate > 0 && c->intermediate % 128 == 0,
                  "intermediate=%d unsupported (multiple of 128: gate/up rows are interleaved in "
                  "128-row blocks for the fused SwiGLU epilogue)", c->intermediate);
  B200Q_CHECK_ARG(c->vocab > 0 && c->vocab % 64 == 0, "vocab=%d must be a multiple of 64", c->vocab);
  B200Q_CHECK_ARG(c->block_size == 16, "block_size=%d unsupported (16)", c->block_size);
  B200Q_CHECK_ARG(c->n_layers > 0 && c->max_tokens > 0 && c->max_seqs > 0 && c->max_pos > 0,
                  "bad sizes L=%d max_tokens=%d max_seqs=%d max_pos=%d", c->n_layers,
                  c->max_tokens, c->max_seqs, c->max_pos);
  B200Q_CHECK_ARG((c->n_q_heads * c->head_dim) % 64 == 0 && qkv_dim(*c) % 64 == 0,
                  "projection widths must be multiples of 64");
  return B200Q_OK;
}

extern "C" {

int64_t b200q_model_workspace_bytes(const b200q_model_config* c) {
  if (check_cfg(c) != B200Q_OK) return -1;
  const int64_t T = c->max_tokens, S = c->max_seqs;
  int64_t b = 0;
  b += 2 * align256(T * c->hidden * 2);                                  // x, residual
  b += align256(T * qkv_dim(*c) * 2);                                    // qkv
  b += align256(T * (int64_t)c->n_q_heads * c->head_dim * 2);            // attn
  b += align256(T * (int64_t)c->intermediate * 2);                       // act
  b += align256(S * (int64_t)c->hidden * 2);                             // sel
  b += align256(S * (int64_t)c->vocab * 2);                              // logits
  return b;
}

int b200q_model_create(const b200q_model_config* cfg, b200q_model_t* out) {
  B200Q_CHECK_ARG(out != nullptr, "out is null");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  rc = b200q_device_check();
  if (rc) return rc;
  b200q_model* m = new b200q_model();
  m->cfg = *cfg;
  {
    const char* v = getenv("B200Q_FUSE_SPLITK");
    m->fuse_splitk = !(v && v[0] == '0');
  }
  m->layers.resize(cfg->n_layers);
  *out = m;
  return B200Q_OK;
}

int b200q_model_destroy(b200q_model_t m) {
  if (m)
    for (cudaEvent_t e : m->ev_pool) cudaEventDestroy(e);
  delete m;
  return B200Q_OK;
}

int b200q_model_bind_weight(b200q_model_t m, const char* name, const void* p, int64_t rows,
                            int64_t cols) {
  B200Q_CHECK_ARG(m && name && p, "bind_weight: null argument");
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15) == 0, "weight %s not 16-byte aligned", name);
  const b200q_model_config& c = m->cfg;
  const int64_t H = c.hidden, I = c.intermediate, QKV = qkv_dim(c),
                QD = (int64_t)c.n_q_heads * c.head_dim;
  auto expect = [&](int64_t r, int64_t cc) -> int {
    B200Q_CHECK_ARG(rows == r && cols == cc, "weight %s has shape [%lld,%lld], expected [%lld,%lld]",
                    name, (long long)rows, (long long)cols, (long long)r, (long long)cc);
    return B200Q_OK;
  };
  int rc;
  if (!strcmp(name, "embed")) {
    if ((rc = expect(c.vocab, H))) return rc;
    m->embed = p;
    if (c.tie_embeddings) m->lm_head = p;
    return B200Q_OK;
  }
  if (!strcmp(name, "final_norm")) {
    if ((rc = expect(1, H))) return rc;
    m->final_norm = p;
    return B200Q_OK;
  }
  if (!strcmp(name, "lm_head")) {
    if ((rc = expect(c.vocab, H))) return rc;
    m->lm_head = p;
    return B200Q_OK;
  }
  int li = -1;
  char field[32] = {0};
  if (sscanf(name, "layers.%d.%31s", &li, field) == 2 && li >= 0 && li < c.n_layers) {
    b200q_layer& L = m->layers[li];
    if (!strcmp(field, "input_norm")) { if ((rc = expect(1, H))) return rc; L.input_norm = p; return B200Q_OK; }
    if (!strcmp(field, "post_norm")) { if ((rc = expect(1, H))) return rc; L.post_norm = p; return B200Q_OK; }
    if (!strcmp(field, "post_attn_norm")) { if ((rc = expect(1, H))) return rc; L.post_attn_norm = p; return B200Q_OK; }
    if (!strcmp(field, "pre_ffn_norm")) { if ((rc = expect(1, H))) return rc; L.pre_ffn_norm = p; return B200Q_OK; }
    if (!strcmp(field, "post_ffn_norm")) { if ((rc = expect(1, H))) return rc; L.post_ffn_norm = p; return B200Q_OK; }
    if (!strcmp(field, "qkv")) { if ((rc = expect(QKV, H))) return rc; L.qkv = p; return B200Q_OK; }
    if (!strcmp(field, "o")) { if ((rc = expect(H, QD))) return rc; L.o = p; return B200Q_OK; }
    if (!strcmp(field, "gate_up")) { if ((rc = expect(2 * I, H))) return rc; L.gate_up = p; return B200Q_OK; }
    if (!strcmp(field, "down")) { if ((rc = expect(H, I))) return rc; L.down = p; return B200Q_OK; }
  }
  set_error("bind_weight: unknown weight name '%s'", name);
  return B200Q_EINVAL;
}

int b200q_model_bind_kv(b200q_model_t m, void* p, int64_t num_blocks) {
  B200Q_CHECK_ARG(m && p && num_blocks > 0, "bind_kv: bad argument");
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 127) == 0, "kv cache must be 128-byte aligned");
  m->kv = (uint8_t*)p;
  m->num_blocks = num_blocks;
  return B200Q_OK;
}

int b200q_model_bind_rope(b200q_model_t m, const void* p) {
  B200Q_CHECK_ARG(m && p, "bind_rope: bad argument");
  m->rope = p;
  return B200Q_OK;
}

int b200q_model_bind_workspace(b200q_model_t m, void* p, int64_t bytes) {
  B200Q_CHECK_ARG(m && p, "bind_workspace: bad argument");
  const b200q_model_config& c = m->cfg;
  const int64_t need = b200q_model_workspace_bytes(&c);
  B200Q_CHECK_ARG(bytes >= need, "workspace too small: %lld < %lld", (long long)bytes,
                  (long long)need);
  B200Q_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 255) == 0, "workspace must be 256-byte aligned");
  const int64_t T = c.max_tokens, S = c.max_seqs;
  uint8_t* q = (uint8_t*)p;
  auto take = [&](int64_t n) {
    uint8_t* r = q;
    q += align256(n);
    return (bf16*)r;
  };
  m->ws = (uint8_t*)p;
  m->ws_bytes = bytes;
  m->x = take(T * c.hidden * 2);
  m->residual = take(T * c.hidden * 2);
  m->qkv = take(T * qkv_dim(c) * 2);
  m->attn = take(T * (int64_t)c.n_q_heads * c.head_dim * 2);
  m->act = take(T * (int64_t)c.intermediate * 2);
  m->sel = take(S * (int64_t)c.hidden * 2);
  m->logits = take(S * (int64_t)c.vocab * 2);
  return B200Q_OK;
}

const void* b200q_model_logits_ptr(b200q_model_t m) { return m ? m->logits : nullptr; }

int b200q_model_set_profiling(b200q_model_t m, int on) {
  B200Q_CHECK_ARG(m, "set_profiling: null model");
  m->profiling = on != 0;
  return B200Q_OK;
}

// resolve the event pairs recorded since the last collect (the stream must be idle) and add
// them to the running per-category totals
int b200q_model_profile_collect(b200q_model_t m, b200q_profile* out, int reset) {
  B200Q_CHECK_ARG(m && out, "profile_collect: null argument");
  for (const auto& sp : m->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->ev_pool[sp.e0], m->ev_pool[sp.e1]) == cudaSuccess) {
      m->prof.ms[sp.cat] += ms;
      m->prof.work[sp.cat] += sp.work;
      m->prof.launches[sp.cat] += 1;
    } else {
      cudaGetLastError();
    }
  }
  m->spans.clear();
  m->ev_used = 0;
  *out = m->prof;
  if (reset) m->prof = b200q_profile{};
  return B200Q_OK;
}

int b200q_model_forward(b200q_model_t m, const b200q_batch* b, void* stream) {
  B200Q_CHECK_ARG(m && b, "forward: null argument");
  const b200q_model_config& c = m->cfg;
  if (!m->embed || !m->final_norm || !m->lm_head || !m->kv || !m->rope || !m->ws) {
    set_error("forward: model not fully bound (embed/final_norm/lm_head/kv/rope/workspace)");
    return B200Q_ESTATE;
  }
  for (int i = 0; i < c.n_layers; ++i) {
    const b200q_layer& L = m->layers[i];
    const bool norms_ok = c.arch == B200Q_ARCH_GEMMA2
                              ? (L.post_attn_norm && L.pre_ffn_norm && L.post_ffn_norm)
                              : L.post_norm != nullptr;
    if (!L.input_norm || !L.qkv || !L.o || !norms_ok || !L.gate_up || !L.down) {
      set_error("forward: layer %d has unbound weights", i);
      return B200Q_ESTATE;
    }
  }
  const int T = b->T;
  B200Q_CHECK_ARG(T >= 0 && T <= c.max_tokens, "forward: T=%d exceeds workspace (%d)", T,
                  c.max_tokens);
  B200Q_CHECK_ARG(b->n_dec >= 0 && b->n_dec <= T && b->n_sample >= 0 && b->n_sample <= c.max_seqs,
                  "forward: bad batch n_dec=%d n_sample=%d", b->n_dec, b->n_sample);
  if (T == 0) return B200Q_OK;

  const int H = c.hidden, D = c.head_dim, NQ = c.n_q_heads, NKV = c.n_kv_heads, I = c.intermediate;
  const int QKV = (int)qkv_dim(c), QD = NQ * D;
  const int64_t kv_layer_bytes = m->num_blocks * 2 * (int64_t)NKV * c.block_size * D * 2;
  int rc;
  cudaStream_t cst = as_stream(stream);
#define B200Q_TRY(cat, work, call)                          \
  do {                                                      \
    int _e0 = -1;                                           \
    if (m->profiling) {                                     \
      _e0 = m->get_event();                                 \
      cudaEventRecord(m->ev_pool[_e0], cst);                \
    }                                                       \
    rc = (call);                                            \
    if (rc) return rc;                                      \
    if (m->profiling) {                                     \
      int _e1 = m->get_event();                             \
      cudaEventRecord(m->ev_pool[_e1], cst);                \
      m->spans.push_back({(cat), (double)(work), _e0, _e1}); \
    }                                                       \
  } while (0)
  const double kv_tok_bytes = 2.0 * NKV * D * 2;  // K+V bytes per cached token per layer

  const bool gemma = c.arch == B200Q_ARCH_GEMMA2;
  // token ids < 0 are indirections into the previous step's sampled ids (async stepping)
  B200Q_TRY(B200Q_PROF_ELEMENTWISE, 2.0 * T * H * 2,
            b200q_embed_ex(b->token_ids, b->prev_out_ids, m->embed, m->residual, T, H,
                           gemma ? c.embed_scale : 0.f, stream));
  // Decode-sized batches (llama arch): a projection that the GEMM splits along K leaves its fp32
  // partials in the library's scratch and the NEXT kernel of the chain reduces them on the fly —
  // qkv -> RoPE/KV write, o -> add+RMSNorm, down -> the next add+RMSNorm — instead of a separate
  // reduce pass per projection (96 launches per 32-layer step).  Bit-identical to the unfused chain.
  const bool fuse = !gemma && m->fuse_splitk;
  const float* pend = nullptr;  // partials of the down projection, consumed by the next norm
  int pend_splits = 1;
  // y = norm(residual += x) with x either in m->x or still in split-K partials
  auto add_norm = [&](const void* w, const float* part, int splits) -> int {
    if (splits > 1) return b200q_add_rmsnorm_splitk(m->x, m->residual, w, part, splits, T, H, c.rms_eps, stream);
    return b200q_add_rmsnorm(m->x, m->residual, w, T, H, c.rms_eps, stream);
  };
  // C = A . W^T, or (fuse && the library splits K) partials only
  auto gemm_maybe_split = [&](const void* A, const void* W, void* C, int N, int K, const float** part,
                              int* splits) -> int {
    *part = nullptr;
    *splits = 1;
    if (fuse) {
      int rc2 = b200q_gemm_bf16_splitk(A, W, T, N, K, stream, part, splits);
      if (rc2 || *splits > 1) return rc2;
    }
    return b200q_gemm_bf16(A, W, C, T, N, K, stream);
  };
  for (int li = 0; li < c.n_layers; ++li) {
    const b200q_layer& L = m->layers[li];
    uint8_t* kv_layer = m->kv + li * kv_layer_bytes;
    // input norm.  llama: layer > 0 fuses the previous MLP's residual add; gemma2: layer > 0 got its
    // normalised input from the previous layer's post-FFN sandwich kernel already
    if (li == 0) {
      if (gemma)
        B200Q_TRY(B200Q_PROF_ELEMENTWISE, 2.0 * T * H * 2, b200q_gemma_rmsnorm(m->residual, L.input_norm, m->x, T, H, c.rms_eps, stream));
      else
        B200Q_TRY(B200Q_PROF_ELEMENTWISE, 2.0 * T * H * 2, b200q_rmsnorm(m->residual, L.input_norm, m->x, T, H, c.rms_eps, stream));
    } else if (!gemma) {
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 4.0 * T * H * 2, add_norm(L.input_norm, pend, pend_splits));
    }
    const float* part = nullptr;
    int splits = 1;
    B200Q_TRY(B200Q_PROF_GEMM, 2.0 * T * QKV * H, gemm_maybe_split(m->x, L.qkv, m->qkv, QKV, H, &part, &splits));
    if (splits > 1)
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, (double)T * (QKV + QD + 2.0 * NKV * D) * 2,
                b200q_rope_kvwrite_splitk(m->qkv, part, splits, m->rope, b->positions, b->slot_mapping, kv_layer, T,
                                          NQ, NKV, D, c.block_size, stream));
    else
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, (double)T * (QKV + QD + 2.0 * NKV * D) * 2, b200q_rope_kvwrite(m->qkv, m->rope, b->positions, b->slot_mapping, kv_layer, T, NQ,
                                 NKV, D, c.block_size, stream));
    // gemma2: even layers are the sliding-window ones (HF layer_types: sliding, full, sliding, ...)
    const int window = (gemma && li % 2 == 0) ? c.sliding_window : 0;
    B200Q_TRY(B200Q_PROF_DECODE_ATTN, (double)b->sum_ctx_dec * kv_tok_bytes, b200q_decode_attn_ex(m->qkv, QKV, m->attn, kv_layer, b->block_table, b->bt_stride,
                                b->ctx_lens, b->n_dec, NQ, NKV, D, c.block_size, c.attn_scale,
                                c.attn_softcap, window, stream));
    B200Q_TRY(B200Q_PROF_PREFILL_ATTN, (double)b->prefill_flops_per_layer, b200q_prefill_attn_ex(m->qkv, QKV, m->attn, kv_layer, b->block_table, b->bt_stride,
                                 b->tiles, b->n_tiles, NQ, NKV, D, c.block_size, c.attn_scale,
                                 c.attn_softcap, window, stream));
    B200Q_TRY(B200Q_PROF_GEMM, 2.0 * T * H * QD, gemm_maybe_split(m->attn, L.o, m->x, H, QD, &part, &splits));
    if (gemma) {
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 4.0 * T * H * 2, b200q_gemma_norm_add_norm(m->x, m->residual, L.post_attn_norm, L.pre_ffn_norm, T, H, c.rms_eps, stream));
      B200Q_TRY(B200Q_PROF_GEMM, 4.0 * T * I * H, b200q_gemm_geglu_bf16(m->x, L.gate_up, m->act, T, 2 * I, H, stream));
    } else {
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 4.0 * T * H * 2, add_norm(L.post_norm, part, splits));
      // gate_up GEMM with the SwiGLU fused into its epilogue (weights interleaved at bind time)
      B200Q_TRY(B200Q_PROF_GEMM, 4.0 * T * I * H, b200q_gemm_swiglu_bf16(m->x, L.gate_up, m->act, T, 2 * I, H, stream));
    }
    B200Q_TRY(B200Q_PROF_GEMM, 2.0 * T * H * I, gemm_maybe_split(m->act, L.down, m->x, H, I, &pend, &pend_splits));
    if (gemma && li + 1 < c.n_layers)
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 4.0 * T * H * 2, b200q_gemma_norm_add_norm(m->x, m->residual, L.post_ffn_norm, m->layers[li + 1].input_norm, T, H, c.rms_eps, stream));
  }
  // the last down projection's residual add (+ final norm).  Without sampling rows nothing reads the
  // result, but pending partials must not outlive the step: fold them in all the same.
  if (b->n_sample > 0 || pend_splits > 1) {
    // only the rows that sample need the final norm + LM head; the fused add+norm runs on all rows
    // because x/residual are per-row anyway and the gather wants the normalised value.
    if (gemma)
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 4.0 * T * H * 2, b200q_gemma_norm_add_norm(m->x, m->residual, m->layers[c.n_layers - 1].post_ffn_norm, m->final_norm, T, H, c.rms_eps, stream));
    else
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 4.0 * T * H * 2, add_norm(m->final_norm, pend, pend_splits));
  }
  if (b->n_sample > 0) {
    B200Q_TRY(B200Q_PROF_ELEMENTWISE, 2.0 * b->n_sample * H * 2, b200q_gather_rows(m->x, b->sample_rows, m->sel, b->n_sample, H, stream));
    B200Q_TRY(B200Q_PROF_GEMM, 2.0 * b->n_sample * (double)c.vocab * H, b200q_gemm_bf16(m->sel, m->lm_head, m->logits, b->n_sample, c.vocab, H, stream));
    if (c.final_softcap > 0.f)
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, 2.0 * b->n_sample * c.vocab * 2,
                b200q_softcap_bf16(m->logits, (int64_t)b->n_sample * c.vocab, c.final_softcap, stream));
    if (b->sample_params)
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, (double)b->n_sample * c.vocab * 2,
                b200q_sample_bf16(m->logits, b->sample_params, b->out_ids, b->n_sample, c.vocab, stream));
    else
      B200Q_TRY(B200Q_PROF_ELEMENTWISE, (double)b->n_sample * c.vocab * 2,
                b200q_argmax_bf16(m->logits, b->out_ids, b->n_sample, c.vocab, stream));
  }
#undef B200Q_TRY
  return B200Q_OK;
}

}  // extern "C"

// accessors for engine.cu
namespace b200q {
const b200q_model_config& model_cfg(b200q_model_t m) { return m->cfg; }
int64_t model_num_blocks(b200q_model_t m) { return m->num_blocks; }
bool model_is_profiling(b200q_model_t m) { return m->profiling; }
}  // namespace b200q
