// engine.cu — continuous-batching scheduler, paged-KV block manager and step loop.
//
// Replaces the part of vLLM that sits behind `engine.generate` in the reference worker
// (ref:llmq/workers/vllm_worker.py:183-186): EngineCore.step = Scheduler.schedule() +
// execute_model + update_from_output.  The *behaviour* mirrored from
// vllm/v1/core/sched/scheduler.py:329-340 is: there is no separate prefill/decode phase — every
// step hands each request `num_tokens - num_computed_tokens` new tokens under a global token
// budget; RUNNING requests are served first, then WAITING ones (FIFO), long prompts are chunked
// by the budget, and when the KV pool is exhausted the most recently admitted running request
// is preempted and later recomputed from its tokens.
//
// The host side is deliberately plain C++: per step it fills ONE pinned staging buffer with all
// int32 metadata (token ids, positions, slot mapping, context lengths, prefill q-tiles, sample
// rows, block table), ships it with one H2D copy, runs the forward on the engine's stream and
// reads back the sampled ids with one D2H copy.
//
// Async stepping (default; B200Q_ASYNC=0 turns it off): step k+1 is scheduled and enqueued while
// step k still runs on the GPU — what vLLM calls async scheduling.  The scheduler only needs to
// know THAT a running request gets one more token, not which: the token's place in the sequence
// is held by a placeholder, and the row that consumes it in step k+1 carries a negative token id
// that the embedding kernel resolves against step k's sampled ids, still on the device
// (b200q_batch.prev_out_ids).  Step k's ids are read back after step k+1 has been enqueued, so the
// GPU never waits for the host's scheduling / launch / Python turn-around.  The cost: a request
// that samples a stop id in step k has already been given one more (discarded) token in step k+1.
// Length stops are known when a step is scheduled and cost nothing.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace b200q {
const b200q_model_config& model_cfg(b200q_model_t m);
int64_t model_num_blocks(b200q_model_t m);
bool model_is_profiling(b200q_model_t m);
}  // namespace b200q

using namespace b200q;

namespace {

struct Request {
  int64_t id;
  std::vector<int32_t> tokens;  // prompt + generated
  int32_t n_prompt;
  int32_t n_computed = 0;       // tokens whose KV is in the cache
  int32_t n_generated = 0;
  int32_t max_new;
  bool ignore_eos;
  std::vector<int32_t> blocks;
  float temperature = 0.f;      // 0 = greedy
  uint64_t seed = 0;
  int32_t n_sched = 0;          // tokens scheduled in the step being built
  // async stepping: tokens[pending_pos] is a placeholder (-1) for the token an in-flight step is
  // sampling into out_ids[pending_slot]; -1 = no unresolved token
  int32_t pending_pos = -1;
  int32_t pending_slot = -1;
  int32_t refs = 0;             // in-flight steps that hold this request (0..2)
  bool dead = false;            // finished / aborted while a later step still holds it: no more events
};

// one scheduled request of an in-flight step
struct StepEntry {
  Request* r;
  int32_t sample_slot;   // index into the step's out_ids, or -1 (mid-prompt chunk)
  int32_t token_pos;     // position in r->tokens the sampled token belongs to
  bool finished_len;     // the request reached max_new / max_model_len with this token
};

struct InFlightStep {
  bool active = false;
  int buf = 0;           // which pinned metadata / output buffer the step uses
  int n_sample = 0;
  std::vector<StepEntry> entries;
};

}  // namespace

struct b200q_engine {
  b200q_model_t model;
  b200q_engine_config cfg;
  b200q_model_config mcfg;
  int block_size;
  int max_blocks_per_seq;
  cudaStream_t stream = nullptr;

  std::unordered_map<int64_t, Request*> by_id;
  std::deque<Request*> waiting;
  std::vector<Request*> running;
  std::vector<int32_t> free_blocks;
  int32_t total_blocks = 0;

  // pinned host buffers are double-buffered (the host fills step k+1's while step k's H2D copy /
  // D2H read-back may still be pending); the device copies are single: everything is stream-ordered
  int32_t* h_meta_buf[2] = {nullptr, nullptr};
  int32_t* h_meta = nullptr;  // = h_meta_buf[buffer of the step being built]
  int32_t* d_meta = nullptr;
  int64_t meta_cap = 0;       // int32 elements
  int32_t* h_out_buf[2] = {nullptr, nullptr};
  int32_t* d_out = nullptr;
  cudaEvent_t done_ev[2] = {nullptr, nullptr};
  bool async_steps = true;    // B200Q_ASYNC
  int next_buf = 0;
  InFlightStep inflight;      // the step running on the GPU while the next one is scheduled
  int32_t prev_n_sample = 0;  // sample slots of the in-flight step (bounds the negative token ids)

  b200q_engine_stats stats{};
  // every token id that ends a request (cfg.eos_token_id plus b200q_engine_set_stop_ids):
  // Llama-3.x-Instruct lists three ids, gemma-2-it adds <end_of_turn> in generation_config.json
  std::vector<int32_t> stop_ids;
  bool is_stop_id(int32_t t) const {
    for (int32_t s : stop_ids)
      if (s == t) return true;
    return false;
  }

  // CUDA graphs of the forward for decode-only steps, keyed by (tokens, block-table stride,
  // sampled?) — every device pointer in the batch is a fixed function of that key.  Removes the
  // ~330 launch gaps per step, which is what small decode batches are bound by.
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::unordered_map<uint64_t, GraphEntry> graphs;
  int graph_epoch = 0;  // tuning_epoch() the cached graphs were captured under
  bool admitting = false;  // admission hysteresis state (policy 1)
  int admit_batch = 0;     // free slots that (re)open admission; 0 = max_num_seqs / 16 (B200Q_ADMIT_BATCH)
  // growth-aware admission (policy 1): running mean of the generated length of finished requests;
  // < 0 until the first request has finished (then every request is assumed to run to max_new)
  double est_gen = -1.0;
  // scheduler self-test mode (b200q_engine_create_dryrun): no model, no CUDA — step() builds the
  // batch metadata exactly as in production, checks its invariants and fabricates the "sampled"
  // token as (previous token + 1) mod vocab.  Host-logic tests only; never reachable from the worker.
  bool dry_run = false;
  bool use_graphs = true;
};

static void free_request_blocks(b200q_engine* e, Request* r) {
  for (int32_t b : r->blocks) e->free_blocks.push_back(b);
  r->blocks.clear();
}

// KV blocks request r is still expected to claim beyond `held` blocks: it is assumed to generate
// min(max_new, max(est_gen, n_generated + 1)) tokens (est_gen < 0: no history yet => max_new)
static int expected_growth_blocks(const b200q_engine* e, const Request* r, int held) {
  int gen = r->max_new;
  if (e->est_gen >= 0.0) gen = std::min(gen, std::max((int)(e->est_gen + 0.999), r->n_generated + 1));
  const int final_len = std::min(r->n_prompt + gen, e->cfg.max_model_len);
  const int blocks = (final_len + e->block_size - 1) / e->block_size;
  return std::max(0, blocks - held);
}

static bool ensure_blocks(b200q_engine* e, Request* r, int n_tokens_total) {
  const int need = (n_tokens_total + e->block_size - 1) / e->block_size;
  const int have = (int)r->blocks.size();
  if (need <= have) return true;
  if ((int)e->free_blocks.size() < need - have) return false;
  for (int i = have; i < need; ++i) {
    r->blocks.push_back(e->free_blocks.back());
    e->free_blocks.pop_back();
  }
  return true;
}

extern "C" {

int b200q_engine_create(b200q_model_t model, const b200q_engine_config* cfg, b200q_engine_t* out) {
  B200Q_CHECK_ARG(model && cfg && out, "engine_create: null argument");
  const b200q_model_config& mc = model_cfg(model);
  B200Q_CHECK_ARG(cfg->max_num_seqs > 0 && cfg->max_num_seqs <= mc.max_seqs,
                  "max_num_seqs=%d must be in [1, model max_seqs=%d]", cfg->max_num_seqs,
                  mc.max_seqs);
  B200Q_CHECK_ARG(cfg->max_num_batched_tokens > 0 && cfg->max_num_batched_tokens <= mc.max_tokens,
                  "max_num_batched_tokens=%d must be in [1, model max_tokens=%d]",
                  cfg->max_num_batched_tokens, mc.max_tokens);
  B200Q_CHECK_ARG(cfg->max_model_len > 1 && cfg->max_model_len <= mc.max_pos,
                  "max_model_len=%d must be in [2, rope table rows=%d]", cfg->max_model_len,
                  mc.max_pos);
  B200Q_CHECK_ARG(model_num_blocks(model) > 0, "engine_create: bind the KV cache first");
  b200q_engine* e = new b200q_engine();
  e->model = model;
  e->cfg = *cfg;
  if (cfg->eos_token_id >= 0) e->stop_ids.push_back(cfg->eos_token_id);
  if (const char* v = getenv("B200Q_ADMIT_BATCH")) e->admit_batch = atoi(v);
  e->mcfg = mc;
  e->block_size = mc.block_size;
  e->max_blocks_per_seq = (cfg->max_model_len + mc.block_size - 1) / mc.block_size;
  e->total_blocks = (int32_t)std::min<int64_t>(model_num_blocks(model), 0x7fffffff);
  e->free_blocks.reserve(e->total_blocks);
  for (int32_t b = e->total_blocks - 1; b >= 0; --b) e->free_blocks.push_back(b);

  const int64_t T = cfg->max_num_batched_tokens, S = cfg->max_num_seqs;
  // token_ids, positions, slot_mapping [T each]; ctx_lens [S]; sample_rows [S];
  // tiles [4 * (T/16 + S)]; block table [S * max_blocks_per_seq]
  e->meta_cap = 3 * T + 2 * S + 4 * S + 8 + 4 * (T / 16 + S + 2) +
                S * (int64_t)((e->max_blocks_per_seq + 7) & ~7) + 64;
  cudaError_t ce;
  if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (ce = cudaMallocHost(&e->h_meta_buf[0], e->meta_cap * 4)) != cudaSuccess ||
      (ce = cudaMallocHost(&e->h_meta_buf[1], e->meta_cap * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&e->d_meta, e->meta_cap * 4)) != cudaSuccess ||
      (ce = cudaMallocHost(&e->h_out_buf[0], S * 4)) != cudaSuccess ||
      (ce = cudaMallocHost(&e->h_out_buf[1], S * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&e->d_out, S * 4)) != cudaSuccess ||
      (ce = cudaMemset(e->d_out, 0, S * 4)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&e->done_ev[0], cudaEventDisableTiming)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&e->done_ev[1], cudaEventDisableTiming)) != cudaSuccess) {
    set_error("engine_create: CUDA allocation failed: %s", cudaGetErrorString(ce));
    b200q_engine_destroy(e);
    return B200Q_ECUDA;
  }
  {
    const char* v = getenv("B200Q_CUDA_GRAPHS");
    e->use_graphs = !(v && v[0] == '0');
    v = getenv("B200Q_ASYNC");
    e->async_steps = !(v && v[0] == '0');
  }
  // weights / KV were produced on other streams (torch); make them visible before first use
  cudaDeviceSynchronize();
  *out = e;
  return B200Q_OK;
}

int b200q_engine_create_dryrun(const b200q_engine_config* cfg, int32_t vocab, int32_t block_size,
                               int32_t num_blocks, b200q_engine_t* out) {
  B200Q_CHECK_ARG(cfg && out && vocab > 1 && block_size > 0 && num_blocks > 0, "create_dryrun: bad argument");
  B200Q_CHECK_ARG(cfg->max_num_seqs > 0 && cfg->max_num_batched_tokens > 0 && cfg->max_model_len > 1,
                  "create_dryrun: bad engine config");
  b200q_engine* e = new b200q_engine();
  e->dry_run = true;
  e->use_graphs = false;
  e->model = nullptr;
  e->cfg = *cfg;
  if (cfg->eos_token_id >= 0) e->stop_ids.push_back(cfg->eos_token_id);
  if (const char* v = getenv("B200Q_ADMIT_BATCH")) e->admit_batch = atoi(v);
  memset(&e->mcfg, 0, sizeof(e->mcfg));
  e->mcfg.vocab = vocab;
  e->mcfg.block_size = block_size;
  e->mcfg.n_q_heads = 1;
  e->mcfg.head_dim = 1;
  e->block_size = block_size;
  e->max_blocks_per_seq = (cfg->max_model_len + block_size - 1) / block_size;
  e->total_blocks = num_blocks;
  for (int32_t b = num_blocks - 1; b >= 0; --b) e->free_blocks.push_back(b);
  const int64_t T = cfg->max_num_batched_tokens, S = cfg->max_num_seqs;
  e->meta_cap = 3 * T + 2 * S + 4 * S + 8 + 4 * (T / 16 + S + 2) +
                S * (int64_t)((e->max_blocks_per_seq + 7) & ~7) + 64;
  for (int i = 0; i < 2; ++i) {
    e->h_meta_buf[i] = (int32_t*)malloc(e->meta_cap * 4);
    e->h_out_buf[i] = (int32_t*)malloc(S * 4);
  }
  {
    const char* v = getenv("B200Q_ASYNC");
    e->async_steps = !(v && v[0] == '0');
  }
  if (!e->h_meta_buf[0] || !e->h_meta_buf[1] || !e->h_out_buf[0] || !e->h_out_buf[1]) {
    set_error("create_dryrun: out of host memory");
    b200q_engine_destroy(e);
    return B200Q_ENOMEM;
  }
  *out = e;
  return B200Q_OK;
}

int b200q_engine_destroy(b200q_engine_t e) {
  if (!e) return B200Q_OK;
  if (!e->dry_run && e->stream) cudaStreamSynchronize(e->stream);
  // requests a still-unfinished step holds but the scheduler has already let go of
  for (StepEntry& se : e->inflight.entries)
    if (se.r->dead && --se.r->refs == 0) delete se.r;
  for (auto& kv : e->by_id) delete kv.second;
  if (e->dry_run) {
    for (int i = 0; i < 2; ++i) {
      free(e->h_meta_buf[i]);
      free(e->h_out_buf[i]);
    }
    delete e;
    return B200Q_OK;
  }
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
  for (int i = 0; i < 2; ++i) {
    if (e->h_meta_buf[i]) cudaFreeHost(e->h_meta_buf[i]);
    if (e->h_out_buf[i]) cudaFreeHost(e->h_out_buf[i]);
    if (e->done_ev[i]) cudaEventDestroy(e->done_ev[i]);
  }
  if (e->d_meta) cudaFree(e->d_meta);
  if (e->d_out) cudaFree(e->d_out);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return B200Q_OK;
}

int b200q_engine_add_request(b200q_engine_t e, int64_t req_id, const int32_t* prompt_ids,
                             int32_t n_prompt, int32_t max_new_tokens, int32_t ignore_eos) {
  B200Q_CHECK_ARG(e && prompt_ids, "add_request: null argument");
  B200Q_CHECK_ARG(n_prompt >= 1, "add_request: empty prompt");
  B200Q_CHECK_ARG(n_prompt < e->cfg.max_model_len,
                  "prompt of %d tokens does not fit max_model_len=%d", n_prompt,
                  e->cfg.max_model_len);
  B200Q_CHECK_ARG(max_new_tokens >= 1, "add_request: max_new_tokens must be >= 1");
  B200Q_CHECK_ARG(e->by_id.find(req_id) == e->by_id.end(), "add_request: duplicate request id %lld",
                  (long long)req_id);
  for (int i = 0; i < n_prompt; ++i)
    B200Q_CHECK_ARG(prompt_ids[i] >= 0 && prompt_ids[i] < e->mcfg.vocab,
                    "add_request: token id %d out of range at %d", prompt_ids[i], i);
  // a request must be able to run alone: its full length has to fit the KV pool
  const int max_total = std::min(n_prompt + max_new_tokens, e->cfg.max_model_len);
  B200Q_CHECK_ARG((max_total + e->block_size - 1) / e->block_size <= e->total_blocks,
                  "request of up to %d tokens cannot fit the KV pool (%d blocks)", max_total,
                  e->total_blocks);
  Request* r = new Request();
  r->id = req_id;
  r->tokens.assign(prompt_ids, prompt_ids + n_prompt);
  r->n_prompt = n_prompt;
  r->max_new = std::min(max_new_tokens, e->cfg.max_model_len - n_prompt);
  r->ignore_eos = ignore_eos != 0;
  e->by_id[req_id] = r;
  e->waiting.push_back(r);
  return B200Q_OK;
}

int b200q_engine_add_request_sampled(b200q_engine_t e, int64_t req_id, const int32_t* prompt_ids,
                                     int32_t n_prompt, int32_t max_new_tokens, int32_t ignore_eos,
                                     float temperature, uint64_t seed) {
  B200Q_CHECK_ARG(temperature >= 0.f && temperature == temperature, "add_request: bad temperature");
  int rc = b200q_engine_add_request(e, req_id, prompt_ids, n_prompt, max_new_tokens, ignore_eos);
  if (rc) return rc;
  Request* r = e->by_id[req_id];
  r->temperature = temperature;
  r->seed = seed;
  return B200Q_OK;
}

int b200q_engine_set_stop_ids(b200q_engine_t e, const int32_t* ids, int32_t n) {
  B200Q_CHECK_ARG(e && (ids || n == 0) && n >= 0 && n <= 64, "set_stop_ids: bad argument (n=%d, at most 64 ids)", n);
  for (int i = 0; i < n; ++i)
    B200Q_CHECK_ARG(ids[i] >= 0 && ids[i] < e->mcfg.vocab, "set_stop_ids: id %d out of range", ids[i]);
  e->stop_ids.assign(ids, ids + n);
  return B200Q_OK;
}

// the scheduler lets go of r (finished or aborted): no further events, blocks back to the pool.
// Its KV blocks may be handed out again at once: whatever an in-flight step still reads or writes
// there happens before the first kernel of the step that reuses them (one stream).
static void retire_request(b200q_engine* e, Request* r) {
  auto w = std::find(e->waiting.begin(), e->waiting.end(), r);
  if (w != e->waiting.end()) e->waiting.erase(w);
  auto ru = std::find(e->running.begin(), e->running.end(), r);
  if (ru != e->running.end()) e->running.erase(ru);
  free_request_blocks(e, r);
  auto it = e->by_id.find(r->id);
  if (it != e->by_id.end() && it->second == r) e->by_id.erase(it);
  if (r->refs == 0) delete r;
  else r->dead = true;  // deleted when the last in-flight step that holds it completes
}

int b200q_engine_set_async(b200q_engine_t e, int32_t on) {
  B200Q_CHECK_ARG(e, "set_async: null engine");
  if (e->inflight.active) {
    set_error("set_async: a step is in flight; switch modes while the engine is idle");
    return B200Q_ESTATE;
  }
  e->async_steps = on != 0;
  return B200Q_OK;
}

int b200q_engine_abort(b200q_engine_t e, int64_t req_id) {
  B200Q_CHECK_ARG(e, "abort: null engine");
  auto it = e->by_id.find(req_id);
  if (it == e->by_id.end()) return B200Q_OK;
  retire_request(e, it->second);
  return B200Q_OK;
}

void* b200q_engine_stream(b200q_engine_t e) { return e ? (void*)e->stream : nullptr; }

int b200q_engine_has_work(b200q_engine_t e) {
  return e && (!e->waiting.empty() || !e->running.empty() || e->inflight.active) ? 1 : 0;
}

int b200q_engine_get_stats(b200q_engine_t e, b200q_engine_stats* out) {
  B200Q_CHECK_ARG(e && out, "get_stats: null argument");
  e->stats.running = (int32_t)e->running.size();
  e->stats.waiting = (int32_t)e->waiting.size();
  e->stats.free_blocks = (int32_t)e->free_blocks.size();
  e->stats.total_blocks = e->total_blocks;
  *out = e->stats;
  return B200Q_OK;
}

// ---- completion of an in-flight step: wait for its sampled ids, put them where the placeholders
// are, emit (request, token, flags) events and retire what finished ----
static int complete_step(b200q_engine* e, InFlightStep& st, int64_t* out_req_ids, int32_t* out_tokens,
                         int32_t* out_flags, int32_t* n_out) {
  if (!e->dry_run) {
    cudaError_t ce = cudaEventSynchronize(e->done_ev[st.buf]);
    if (ce != cudaSuccess) {
      set_error("step: forward failed: %s", cudaGetErrorString(ce));
      return B200Q_ECUDA;
    }
  }
  const int32_t* h_out = e->h_out_buf[st.buf];
  int n_ev = *n_out;
  for (StepEntry& se : st.entries) {
    Request* r = se.r;
    r->refs--;
    if (r->dead) {  // finished (stop id) or aborted after this step had been scheduled: output discarded
      if (r->refs == 0) delete r;
      continue;
    }
    if (se.sample_slot < 0) continue;
    // scheduler self-test engine: the "model" counts up from the previous token of the sequence
    const int32_t t = e->dry_run ? (r->tokens[se.token_pos - 1] + 1) % e->mcfg.vocab : h_out[se.sample_slot];
    r->tokens[se.token_pos] = t;
    if (r->pending_pos == se.token_pos) r->pending_pos = -1;
    int flags = 0;
    if (!r->ignore_eos && e->is_stop_id(t)) flags = B200Q_FLAG_FINISHED_EOS;
    else if (se.finished_len) flags = B200Q_FLAG_FINISHED_LENGTH;
    out_req_ids[n_ev] = r->id;
    out_tokens[n_ev] = t;
    out_flags[n_ev] = flags;
    ++n_ev;
    if (flags) {
      // generated-length history for growth-aware admission: mean of the first finishers, then
      // an exponential average (window ~64 requests)
      const double gen = (double)(se.token_pos + 1 - r->n_prompt);
      e->est_gen = e->est_gen < 0.0 ? gen : e->est_gen + (gen - e->est_gen) / 64.0;
      retire_request(e, r);  // a later in-flight step may still hold it: its extra token is discarded
    }
  }
  st.entries.clear();
  st.active = false;
  *n_out = n_ev;
  return B200Q_OK;
}

// ---- scheduling + launch of the next step; `step.active` stays false when there is nothing to run
static int schedule_and_launch(b200q_engine* e, InFlightStep& step) {
  int budget = e->cfg.max_num_batched_tokens;

  // ---- 1+2. pick this step's work ----
  // policy 0 (vLLM order): RUNNING requests first (decodes and unfinished prompt chunks), then
  //   FIFO admission of WAITING requests with what is left of the token budget.
  // policy 1 (prefill first, default): unfinished prompt chunks and admissions take the budget
  //   first, decodes get the rest.  Under a backlog this drives the decode batch to max_num_seqs
  //   as fast as possible and keeps it there, which is what the GEMMs (M = tokens per step) and
  //   the paged attention want on a B200; per-token latency is not this path's metric.
  std::vector<Request*> sched;
  sched.reserve(e->running.size() + 16);
  for (Request* r : e->running) r->n_sched = 0;

  auto schedule_running = [&](int kind) {  // 0: all, 1: multi-token remainders, 2: single-token
    for (size_t i = 0; i < e->running.size();) {
      Request* r = e->running[i];
      const int remaining = (int)r->tokens.size() - r->n_computed;
      const bool single = remaining == 1;
      if (r->n_sched > 0 || remaining <= 0 || (kind == 1 && single) || (kind == 2 && !single)) {
        ++i;
        continue;
      }
      const int n_new = std::min(remaining, budget);
      if (n_new <= 0) return;  // token budget exhausted
      bool ok = ensure_blocks(e, r, r->n_computed + n_new);
      bool self_preempted = false;
      while (!ok) {
        // KV pool exhausted: preempt the most recently admitted running request; it keeps its
        // tokens and is recomputed when re-admitted (preempt-by-recompute)
        Request* victim = e->running.back();
        e->running.pop_back();
        auto sit = std::find(sched.begin(), sched.end(), victim);
        if (sit != sched.end()) {
          budget += victim->n_sched;
          sched.erase(sit);
        }
        free_request_blocks(e, victim);
        victim->n_computed = 0;
        victim->n_sched = 0;
        e->waiting.push_front(victim);
        e->stats.preemptions++;
        if (victim == r) {
          self_preempted = true;
          break;
        }
        ok = ensure_blocks(e, r, r->n_computed + n_new);
      }
      if (self_preempted) return;  // r was the newest request: nothing behind it is left
      r->n_sched = n_new;
      budget -= n_new;
      sched.push_back(r);
      ++i;
    }
  };
  auto admit_waiting = [&]() {
    // policy 1 also reserves the blocks the RUNNING requests are expected to grow into (their
    // expected final length, see expected_growth_blocks): admitting on today's free blocks alone
    // over-commits the pool under a steady backlog, and every later preemption recomputes a whole
    // sequence (3 % of all tokens at 4608 x 128-in/128-out on one B200, tools/sched_sim.py).
    // Requests that outgrow the estimate are still handled by preempt-by-recompute.
    // The reservation only pays while there is a backlog to fill the batch from later: when every
    // waiting request fits the free slots (the queue is draining) a deferred request would run on
    // as a small-batch tail of its own, whereas admitting it optimistically costs at most a late
    // preemption + one recompute step (dry-run: 272 vs 382 steps for 4608 jobs in a pool 1 % short).
    const bool gate = e->cfg.policy == 1 &&
                      (int)e->waiting.size() > e->cfg.max_num_seqs - (int)e->running.size();
    int64_t growth = 0;
    if (gate)
      for (const Request* q : e->running) growth += expected_growth_blocks(e, q, (int)q->blocks.size());
    while (budget > 0 && !e->waiting.empty() && (int)e->running.size() < e->cfg.max_num_seqs) {
      Request* r = e->waiting.front();
      const int n_new = std::min((int)r->tokens.size() - r->n_computed, budget);
      // admission control: the blocks for the request's WHOLE current sequence (prompt, plus the
      // generated tokens a preempted request has to recompute) must be free, on top of a watermark of
      // ~2 steps' worth of fresh blocks for the running decodes.  Admitting on the first chunk alone
      // lets a request start a prefill it cannot finish: it preempts itself at the last block, is
      // re-admitted at once (prefill first) and starves the decodes whose completion would have
      // freed the memory — a livelock found by tests/test_scheduler_dryrun.py.
      const int seq_blocks = ((int)r->tokens.size() + e->block_size - 1) / e->block_size;
      const int need = seq_blocks - (int)r->blocks.size();
      const int reserve = e->running.empty() ? 0 : (int)e->running.size() / 8 + 1;
      if ((int)e->free_blocks.size() - need < reserve) break;
      const int r_growth = gate ? expected_growth_blocks(e, r, seq_blocks) : 0;
      if (gate && !e->running.empty() && (int64_t)e->free_blocks.size() - need - r_growth < growth)
        break;
      if (!ensure_blocks(e, r, r->n_computed + n_new)) break;
      growth += r_growth;
      e->waiting.pop_front();
      e->running.push_back(r);
      r->n_sched = n_new;
      budget -= n_new;
      sched.push_back(r);
    }
  };
  if (e->cfg.policy == 0) {
    schedule_running(0);
    admit_waiting();
  } else {
    schedule_running(1);
    // admission hysteresis: under a steady backlog a slot frees up almost every step, and
    // admitting one prompt at a time would put a tiny prefill chunk into every step (no step is
    // decode-only => no CUDA-graph replay, prefill GEMMs at M ~ 128).  Wait until a batch of
    // slots (1/16 of max_num_seqs) is free, then admit until the slots or the budget run out.
    const int free_slots = e->cfg.max_num_seqs - (int)e->running.size();
    const int thresh = e->admit_batch > 0 ? e->admit_batch : std::max(1, e->cfg.max_num_seqs / 16);
    const bool mid_prefill = !sched.empty();  // an unfinished prompt is in flight: keep the phase going
    if (e->running.empty() || mid_prefill || free_slots >= thresh || e->admitting) {
      const size_t before = e->running.size();
      admit_waiting();
      e->admitting = e->running.size() > before && !e->waiting.empty() &&
                     (int)e->running.size() < e->cfg.max_num_seqs;
    }
    schedule_running(2);
  }


  if (sched.empty()) return B200Q_OK;

  // ---- 3. order: single-token pieces (decode) first, then multi-token prefill chunks ----
  std::stable_partition(sched.begin(), sched.end(), [](Request* r) { return r->n_sched == 1; });
  int n_dec = 0, T = 0, max_blocks = 1;
  for (Request* r : sched) {
    if (r->n_sched == 1) ++n_dec;
    T += r->n_sched;
    max_blocks = std::max(max_blocks, (int)r->blocks.size());
  }
  const int n_rows = (int)sched.size();
  const int bt_stride = (max_blocks + 7) & ~7;

  const int buf = e->next_buf;
  int32_t* const h_meta = e->h_meta_buf[buf];
  int32_t* tok = h_meta;
  int32_t* pos = tok + T;
  int32_t* slot = pos + T;
  int32_t* ctx = slot + T;
  int32_t* srows = ctx + n_dec;
  int32_t* sparams = srows + n_rows;  // reserve n_rows sample slots
  sparams += (4 - ((sparams - h_meta) & 3)) & 3;
  int32_t* tiles = sparams + 4 * n_rows;  // 16-byte aligned: the prefill kernel reads tiles as int4
  int n_tiles = 0, n_sample = 0, row = 0;
  bool any_sampled = false;
  // count tiles first to place the block table after them
  for (Request* r : sched)
    if (r->n_sched > 1) n_tiles += (r->n_sched + 15) / 16;
  int32_t* btab = tiles + 4 * n_tiles;
  // keep the block table 16-byte aligned for tidy copies
  btab += (4 - ((btab - h_meta) & 3)) & 3;
  const int64_t used = (btab - h_meta) + (int64_t)n_rows * bt_stride;
  if (used > e->meta_cap) {
    set_error("step: metadata buffer overflow (%lld > %lld)", (long long)used,
              (long long)e->meta_cap);
    return B200Q_ENOMEM;
  }
  int ti = 0;
  std::vector<int32_t> slot_of(n_rows, -1);  // sample slot of row ri, -1 = a mid-prompt chunk
  for (int ri = 0; ri < n_rows; ++ri) {
    Request* r = sched[ri];
    const int bs = e->block_size;
    for (int j = 0; j < r->n_sched; ++j) {
      const int p = r->n_computed + j;
      // an unresolved token (sampled by the in-flight step, not read back yet) travels as a negative
      // id that the embedding kernel resolves against that step's out_ids on the device
      tok[row + j] = p == r->pending_pos ? -1 - r->pending_slot : r->tokens[p];
      pos[row + j] = p;
      slot[row + j] = r->blocks[p / bs] * bs + (p % bs);
    }
    if (r->n_sched == 1) {
      ctx[ri] = r->n_computed + 1;
    } else {
      for (int j = 0; j < r->n_sched; j += 16) {
        tiles[4 * ti + 0] = ri;
        tiles[4 * ti + 1] = row + j;
        tiles[4 * ti + 2] = std::min(16, r->n_sched - j);
        tiles[4 * ti + 3] = r->n_computed + j;
        ++ti;
      }
    }
    if (r->n_computed + r->n_sched == (int)r->tokens.size()) {
      slot_of[ri] = n_sample;
      float tf = r->temperature;
      int32_t tbits;
      memcpy(&tbits, &tf, 4);
      sparams[4 * n_sample + 0] = tbits;
      sparams[4 * n_sample + 1] = (int32_t)(uint32_t)(r->seed & 0xffffffffu);
      sparams[4 * n_sample + 2] = (int32_t)(uint32_t)(r->seed >> 32);
      sparams[4 * n_sample + 3] = r->n_generated;  // Philox counter: index of the token being drawn
      any_sampled |= tf > 0.f;
      srows[n_sample++] = row + r->n_sched - 1;
    }
    int32_t* brow = btab + (int64_t)ri * bt_stride;
    const int nb = (int)r->blocks.size();
    for (int k = 0; k < nb; ++k) brow[k] = r->blocks[k];
    for (int k = nb; k < bt_stride; ++k) brow[k] = 0;
    row += r->n_sched;
  }

  if (e->dry_run) {
    // ---- invariants of the metadata the kernels would consume ----
    int rc = B200Q_OK;
    auto fail = [&](const char* what) {
      set_error("scheduler self-test: %s (T=%d n_dec=%d n_tiles=%d)", what, T, n_dec, n_tiles);
      rc = B200Q_ESTATE;
    };
    if (T > e->cfg.max_num_batched_tokens) fail("token budget exceeded");
    if (n_rows > e->cfg.max_num_seqs) fail("more sequences than max_num_seqs");
    std::vector<char> slot_seen((size_t)e->total_blocks * e->block_size, 0);
    for (int t = 0; t < T && rc == B200Q_OK; ++t) {
      if (slot[t] < 0 || slot[t] >= (int)slot_seen.size()) fail("slot out of range");
      else if (slot_seen[slot[t]]++) fail("two tokens of one step share a KV slot");
      if (tok[t] >= e->mcfg.vocab || (tok[t] < 0 && -1 - tok[t] >= e->prev_n_sample))
        fail("token id out of range");
    }
    int covered = n_dec;
    for (int i = 0; i < n_tiles && rc == B200Q_OK; ++i) {
      const int32_t* tl = tiles + 4 * i;
      if (tl[1] != covered || tl[2] < 1 || tl[2] > 16 || tl[0] < n_dec || tl[0] >= n_rows) fail("bad prefill tile");
      else if (pos[tl[1]] != tl[3]) fail("tile position does not match its first token");
      covered += tl[2];
    }
    if (rc == B200Q_OK && covered != T) fail("prefill tiles do not cover the prefill rows exactly");
    for (int ri = 0; ri < n_dec && rc == B200Q_OK; ++ri)
      if (ctx[ri] != pos[ri] + 1) fail("decode context length != position + 1");
    if (rc) return rc;
    // (the "sampled" tokens are fabricated when the step completes, in sequence order)
  } else {
    cudaError_t ce = cudaMemcpyAsync(e->d_meta, h_meta, used * 4, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) {
      set_error("step: H2D metadata copy failed: %s", cudaGetErrorString(ce));
      return B200Q_ECUDA;
    }
    b200q_batch b;
    b.T = T;
    b.n_dec = n_dec;
    b.n_tiles = n_tiles;
    b.n_sample = n_sample;
    b.bt_stride = bt_stride;
    b.token_ids = e->d_meta + (tok - h_meta);
    b.positions = e->d_meta + (pos - h_meta);
    b.slot_mapping = e->d_meta + (slot - h_meta);
    b.ctx_lens = e->d_meta + (ctx - h_meta);
    b.sample_rows = e->d_meta + (srows - h_meta);
    b.tiles = e->d_meta + (tiles - h_meta);
    b.block_table = e->d_meta + (btab - h_meta);
    b.out_ids = e->d_out;
    b.prev_out_ids = e->d_out;  // the in-flight step's ids: read by this step's embedding before its sampler overwrites them
    b.sample_params = any_sampled ? e->d_meta + (sparams - h_meta) : nullptr;
    b.sum_ctx_dec = 0;
    b.prefill_flops_per_layer = 0;
    for (Request* r : sched) {
      if (r->n_sched == 1) {
        b.sum_ctx_dec += r->n_computed + 1;
      } else {
        // causal: query j of the chunk sees n_computed + j + 1 keys; QK^T and PV, 2 flops per MAC
        const double q = r->n_sched, c0 = r->n_computed;
        const double pairs = q * c0 + q * (q + 1) / 2;
        b.prefill_flops_per_layer +=
            (int64_t)(4.0 * pairs * e->mcfg.n_q_heads * e->mcfg.head_dim);
      }
    }
    int rc = B200Q_OK;
    bool launched = false;
    if (e->use_graphs && n_tiles == 0 && T == n_dec && T < (1 << 21) && bt_stride < (1 << 20) &&
        !model_is_profiling(e->model) &&
        e->stats.steps >= 2 /* first steps run eagerly: one-time attribute/occupancy/scratch setup */) {
      if (e->graph_epoch != tuning_epoch()) {  // a tuning hook changed kernel selection: rebuild
        for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
        e->graphs.clear();
        e->graph_epoch = tuning_epoch();
      }
      // everything the captured forward bakes in: T (bits 0-20), sampling rows (21-41: a decode-only
      // step can carry a single-token mid-prefill row that does not sample), block-table stride
      // (42-61), sampled flag (62).  T < 2^21 is checked by the caller's condition above.
      const uint64_t key = (uint64_t)(uint32_t)T | ((uint64_t)(uint32_t)n_sample << 21) |
                           ((uint64_t)bt_stride << 42) | ((uint64_t)(any_sampled ? 1 : 0) << 62);
      auto it = e->graphs.find(key);
      if (it == e->graphs.end()) {
        const int64_t l0 = b200q_launch_count();
        cudaGraph_t g = nullptr;
        b200q_engine::GraphEntry ge;
        if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
          rc = b200q_model_forward(e->model, &b, e->stream);
          cudaError_t ee = cudaStreamEndCapture(e->stream, &g);
          if (rc == B200Q_OK && ee == cudaSuccess && g &&
              cudaGraphInstantiate(&ge.exec, g, 0) == cudaSuccess) {
            ge.launches = b200q_launch_count() - l0;
            if (e->graphs.size() > 512) {  // bounded cache
              for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
              e->graphs.clear();
            }
            it = e->graphs.emplace(key, ge).first;
          } else {
            cudaGetLastError();
            e->use_graphs = false;  // fall back to eager launches for good
            rc = B200Q_OK;
          }
          if (g) cudaGraphDestroy(g);
        } else {
          cudaGetLastError();
          e->use_graphs = false;
        }
      } else {
        count_launch((int)it->second.launches);  // the replay runs the same kernels again
      }
      if (it != e->graphs.end() && e->use_graphs) {
        if (cudaGraphLaunch(it->second.exec, e->stream) == cudaSuccess) {
          launched = true;
        } else {
          cudaGetLastError();
          e->use_graphs = false;
        }
      }
    }
    if (!launched) rc = b200q_model_forward(e->model, &b, e->stream);
    if (rc) return rc;
    if (n_sample > 0) {
      ce = cudaMemcpyAsync(e->h_out_buf[buf], e->d_out, (size_t)n_sample * 4, cudaMemcpyDeviceToHost, e->stream);
      if (ce != cudaSuccess) {
        set_error("step: D2H copy failed: %s", cudaGetErrorString(ce));
        return B200Q_ECUDA;
      }
    }
    ce = cudaEventRecord(e->done_ev[buf], e->stream);
    if (ce != cudaSuccess) {
      set_error("step: event record failed: %s", cudaGetErrorString(ce));
      return B200Q_ECUDA;
    }
  }


  // ---- 4. the step is on its way: advance the scheduler's view of every request in it ----
  e->stats.h2d_bytes += used * 4;
  e->stats.d2h_bytes += (int64_t)n_sample * 4;
  e->stats.steps++;
  e->stats.last_step_tokens = T;
  e->stats.last_step_seqs = n_rows;
  step.active = true;
  step.buf = buf;
  step.n_sample = n_sample;
  step.entries.clear();
  step.entries.reserve(n_rows);
  e->next_buf ^= 1;
  for (int ri = 0; ri < n_rows; ++ri) {
    Request* r = sched[ri];
    const bool was_decode = r->n_computed >= r->n_prompt;
    r->n_computed += r->n_sched;
    if (was_decode) e->stats.tokens_decoded += r->n_sched;
    else e->stats.tokens_prefilled += r->n_sched;
    r->n_sched = 0;
    r->refs++;
    StepEntry se{r, slot_of[ri], -1, false};
    if (se.sample_slot >= 0) {
      // the token being sampled gets its place in the sequence now; its value follows at completion
      r->tokens.push_back(-1);
      se.token_pos = (int32_t)r->tokens.size() - 1;
      r->pending_pos = se.token_pos;
      r->pending_slot = se.sample_slot;
      r->n_generated++;
      se.finished_len = r->n_generated >= r->max_new || (int)r->tokens.size() >= e->cfg.max_model_len;
    }
    step.entries.push_back(se);
  }
  // length stops are known now: such a request takes no part in later steps and its blocks can be
  // handed out again (stream order protects the step in flight); its last event follows at completion
  for (StepEntry& se : step.entries)
    if (se.finished_len) {
      auto ru = std::find(e->running.begin(), e->running.end(), se.r);
      if (ru != e->running.end()) e->running.erase(ru);
      free_request_blocks(e, se.r);
    }
  return B200Q_OK;
}

int b200q_engine_step(b200q_engine_t e, int64_t* out_req_ids, int32_t* out_tokens,
                      int32_t* out_flags, int32_t cap, int32_t* n_out) {
  B200Q_CHECK_ARG(e && out_req_ids && out_tokens && out_flags && n_out, "step: null argument");
  B200Q_CHECK_ARG(cap >= e->cfg.max_num_seqs, "step: event capacity %d < max_num_seqs %d", cap,
                  e->cfg.max_num_seqs);
  *n_out = 0;
  e->prev_n_sample = e->inflight.active ? e->inflight.n_sample : 0;
  // 1. schedule and enqueue the next step while the previous one is still running
  InFlightStep next;
  int rc = schedule_and_launch(e, next);
  if (rc) return rc;
  const bool had_inflight = e->inflight.active;
  // 2. now read the previous step's ids back and turn them into events
  if (had_inflight && (rc = complete_step(e, e->inflight, out_req_ids, out_tokens, out_flags, n_out))) return rc;
  if (next.active) {
    if (e->async_steps) {
      e->inflight = std::move(next);
    } else if ((rc = complete_step(e, next, out_req_ids, out_tokens, out_flags, n_out))) {
      return rc;
    }
  } else if (!had_inflight && !e->waiting.empty() && e->running.empty()) {
    set_error("scheduler stalled: a waiting request cannot be admitted (KV pool too small)");
    return B200Q_ENOMEM;
  }
  return B200Q_OK;
}

}  // extern "C"
