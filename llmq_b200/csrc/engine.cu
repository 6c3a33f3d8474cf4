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
  int32_t n_sched = 0;          // tokens scheduled in the current step
  int32_t sample_slot = -1;     // index into out_ids for this step, or -1
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

  int32_t* h_meta = nullptr;  // pinned
  int32_t* d_meta = nullptr;
  int64_t meta_cap = 0;       // int32 elements
  int32_t* h_out = nullptr;   // pinned
  int32_t* d_out = nullptr;

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
      (ce = cudaMallocHost(&e->h_meta, e->meta_cap * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&e->d_meta, e->meta_cap * 4)) != cudaSuccess ||
      (ce = cudaMallocHost(&e->h_out, S * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&e->d_out, S * 4)) != cudaSuccess) {
    set_error("engine_create: CUDA allocation failed: %s", cudaGetErrorString(ce));
    b200q_engine_destroy(e);
    return B200Q_ECUDA;
  }
  {
    const char* v = getenv("B200Q_CUDA_GRAPHS");
    e->use_graphs = !(v && v[0] == '0');
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
  e->h_meta = (int32_t*)malloc(e->meta_cap * 4);
  e->h_out = (int32_t*)malloc(S * 4);
  if (!e->h_meta || !e->h_out) {
    set_error("create_dryrun: out of host memory");
    b200q_engine_destroy(e);
    return B200Q_ENOMEM;
  }
  *out = e;
  return B200Q_OK;
}

int b200q_engine_destroy(b200q_engine_t e) {
  if (!e) return B200Q_OK;
  if (e->dry_run) {
    for (auto& kv : e->by_id) delete kv.second;
    free(e->h_meta);
    free(e->h_out);
    delete e;
    return B200Q_OK;
  }
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
  for (auto& kv : e->by_id) delete kv.second;
  if (e->h_meta) cudaFreeHost(e->h_meta);
  if (e->d_meta) cudaFree(e->d_meta);
  if (e->h_out) cudaFreeHost(e->h_out);
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

int b200q_engine_abort(b200q_engine_t e, int64_t req_id) {
  B200Q_CHECK_ARG(e, "abort: null engine");
  auto it = e->by_id.find(req_id);
  if (it == e->by_id.end()) return B200Q_OK;
  Request* r = it->second;
  auto w = std::find(e->waiting.begin(), e->waiting.end(), r);
  if (w != e->waiting.end()) e->waiting.erase(w);
  auto ru = std::find(e->running.begin(), e->running.end(), r);
  if (ru != e->running.end()) e->running.erase(ru);
  free_request_blocks(e, r);
  e->by_id.erase(it);
  delete r;
  return B200Q_OK;
}

void* b200q_engine_stream(b200q_engine_t e) { return e ? (void*)e->stream : nullptr; }

int b200q_engine_has_work(b200q_engine_t e) {
  return e && (!e->waiting.empty() || !e->running.empty()) ? 1 : 0;
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

int b200q_engine_step(b200q_engine_t e, int64_t* out_req_ids, int32_t* out_tokens,
                      int32_t* out_flags, int32_t cap, int32_t* n_out) {
  B200Q_CHECK_ARG(e && out_req_ids && out_tokens && out_flags && n_out, "step: null argument");
  B200Q_CHECK_ARG(cap >= e->cfg.max_num_seqs, "step: event capacity %d < max_num_seqs %d", cap,
                  e->cfg.max_num_seqs);
  *n_out = 0;
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

  if (sched.empty()) {
    if (!e->waiting.empty() && e->running.empty()) {
      set_error("scheduler stalled: a waiting request cannot be admitted (KV pool too small)");
      return B200Q_ENOMEM;
    }
    return B200Q_OK;
  }

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

  int32_t* tok = e->h_meta;
  int32_t* pos = tok + T;
  int32_t* slot = pos + T;
  int32_t* ctx = slot + T;
  int32_t* srows = ctx + n_dec;
  int32_t* sparams = srows + n_rows;  // reserve n_rows sample slots
  sparams += (4 - ((sparams - e->h_meta) & 3)) & 3;
  int32_t* tiles = sparams + 4 * n_rows;  // 16-byte aligned: the prefill kernel reads tiles as int4
  int n_tiles = 0, n_sample = 0, row = 0;
  bool any_sampled = false;
  // count tiles first to place the block table after them
  for (Request* r : sched)
    if (r->n_sched > 1) n_tiles += (r->n_sched + 15) / 16;
  int32_t* btab = tiles + 4 * n_tiles;
  // keep the block table 16-byte aligned for tidy copies
  btab += (4 - ((btab - e->h_meta) & 3)) & 3;
  const int64_t used = (btab - e->h_meta) + (int64_t)n_rows * bt_stride;
  if (used > e->meta_cap) {
    set_error("step: metadata buffer overflow (%lld > %lld)", (long long)used,
              (long long)e->meta_cap);
    return B200Q_ENOMEM;
  }
  int ti = 0;
  for (int ri = 0; ri < n_rows; ++ri) {
    Request* r = sched[ri];
    const int bs = e->block_size;
    for (int j = 0; j < r->n_sched; ++j) {
      const int p = r->n_computed + j;
      tok[row + j] = r->tokens[p];
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
      r->sample_slot = n_sample;
      float tf = r->temperature;
      int32_t tbits;
      memcpy(&tbits, &tf, 4);
      sparams[4 * n_sample + 0] = tbits;
      sparams[4 * n_sample + 1] = (int32_t)(uint32_t)(r->seed & 0xffffffffu);
      sparams[4 * n_sample + 2] = (int32_t)(uint32_t)(r->seed >> 32);
      sparams[4 * n_sample + 3] = r->n_generated;  // Philox counter: index of the token being drawn
      any_sampled |= tf > 0.f;
      srows[n_sample++] = row + r->n_sched - 1;
    } else {
      r->sample_slot = -1;
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
      if (tok[t] < 0 || tok[t] >= e->mcfg.vocab) fail("token id out of range");
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
    for (Request* r : sched)
      if (r->sample_slot >= 0) e->h_out[r->sample_slot] = (r->tokens.back() + 1) % e->mcfg.vocab;
  } else {
    cudaError_t ce = cudaMemcpyAsync(e->d_meta, e->h_meta, used * 4, cudaMemcpyHostToDevice, e->stream);
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
    b.token_ids = e->d_meta + (tok - e->h_meta);
    b.positions = e->d_meta + (pos - e->h_meta);
    b.slot_mapping = e->d_meta + (slot - e->h_meta);
    b.ctx_lens = e->d_meta + (ctx - e->h_meta);
    b.sample_rows = e->d_meta + (srows - e->h_meta);
    b.tiles = e->d_meta + (tiles - e->h_meta);
    b.block_table = e->d_meta + (btab - e->h_meta);
    b.out_ids = e->d_out;
    b.sample_params = any_sampled ? e->d_meta + (sparams - e->h_meta) : nullptr;
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
      ce = cudaMemcpyAsync(e->h_out, e->d_out, (size_t)n_sample * 4, cudaMemcpyDeviceToHost, e->stream);
      if (ce != cudaSuccess) {
        set_error("step: D2H copy failed: %s", cudaGetErrorString(ce));
        return B200Q_ECUDA;
      }
    }
    ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) {
      set_error("step: forward failed: %s", cudaGetErrorString(ce));
      return B200Q_ECUDA;
    }
  }

  // ---- 4. update from output ----
  e->stats.h2d_bytes += used * 4;
  e->stats.d2h_bytes += (int64_t)n_sample * 4;
  e->stats.steps++;
  e->stats.last_step_tokens = T;
  e->stats.last_step_seqs = n_rows;
  int n_ev = 0;
  for (Request* r : sched) {
    const bool was_decode = r->n_computed >= r->n_prompt;
    r->n_computed += r->n_sched;
    if (was_decode) e->stats.tokens_decoded += r->n_sched;
    else e->stats.tokens_prefilled += r->n_sched;
    r->n_sched = 0;
    if (r->sample_slot < 0) continue;
    const int32_t t = e->h_out[r->sample_slot];
    r->tokens.push_back(t);
    r->n_generated++;
    int flags = 0;
    if (!r->ignore_eos && e->is_stop_id(t))
      flags = B200Q_FLAG_FINISHED_EOS;
    else if (r->n_generated >= r->max_new || (int)r->tokens.size() >= e->cfg.max_model_len)
      flags = B200Q_FLAG_FINISHED_LENGTH;
    out_req_ids[n_ev] = r->id;
    out_tokens[n_ev] = t;
    out_flags[n_ev] = flags;
    ++n_ev;
    if (flags) {
      // generated-length history for growth-aware admission: mean of the first finishers, then
      // an exponential average (window ~64 requests)
      e->est_gen = e->est_gen < 0.0 ? (double)r->n_generated
                                    : e->est_gen + ((double)r->n_generated - e->est_gen) / 64.0;
      free_request_blocks(e, r);
      e->running.erase(std::find(e->running.begin(), e->running.end(), r));
      e->by_id.erase(r->id);
      delete r;
    }
  }
  *n_out = n_ev;
  return B200Q_OK;
}

}  // extern "C"
