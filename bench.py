#!/usr/bin/env python
"""bench.py — headline benchmark of the b200 worker's hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # native arm
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json `metric`): Llama-3-8B, random-init bf16, 128 prompt tokens in / 128
generated tokens out, queue-sharded data parallel — one worker process per GPU, job i of the
canonical seeded job stream goes to rank i % N, no data-path collective.

The engine runs in steady state: `--max-num-seqs` (4608) sequences in flight and a backlog behind
them that never empties — the situation of a worker on a 100k-prompt queue.  A *step* is `--jobs`
(1152) COMPLETED jobs per GPU: the timed region is exactly K steps = K x 1152 completions per GPU
(about 4 s each on one B200), after W untimed steps that include the ramp (the cohort that fills
the empty engine has staggered output caps, so the batch holds requests of every age).
  value : output tokens/s (tokens sampled inside the timed region) with the prompts already
          tokenised and queued inside the engine (weights, KV pool, workspace resident in HBM) —
          device-side (CUDA events on the engine's stream), max over ranks.
  e2e   : the same metric through the worker-facing call, `B200Worker._process_job(Job)` — what
          the reference's `BaseWorker._process_message` awaits — behind an un-acked window of
          max_num_seqs + jobs messages (the role of VLLM_QUEUE_PREFETCH): host text prompts ->
          tokenise -> engine thread -> per-step H2D metadata / D2H sampled ids -> detokenise ->
          host text; min(K, 8) steps by wall clock, max over ranks.
  roofline : live CUDA-event timing of every kernel launch in the timed region (events recorded
          on the engine's stream by libb200q), GEMM = dominant kernel (tensor-bound), plus the
          decode-attention HBM fraction.
  cpu_baseline / --impl reference : the CPU oracle (oracle/model.py, a torch-CPU restatement
          of the vLLM/HF Llama forward; vLLM-CPU is not installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "output_tokens_per_sec"
UNIT = "tokens/s"


def ensure_llmq_importable():
    """the reference package (pip --target baseline/_ref) + the aio_pika/semhash
    stand-ins this image lacks (tests/shims) — needed to instantiate the real BaseWorker slot"""
    try:
        import aio_pika  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
    try:
        import llmq.workers.base  # noqa: F401
    except ImportError:
        for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
            if os.path.isdir(os.path.join(p, "llmq")):
                sys.path.insert(0, p)
                break
        import llmq.workers.base  # noqa: F401


def usable_host_cores() -> int:
    """threads the CPU arm may really use: scheduler affinity, capped by the cgroup CPU quota
    (os.cpu_count() reports the whole host inside a limited container and oversubscribes BLAS)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, int(os.environ.get("B200Q_CPU_THREADS", "64"))))


def ncu_traffic():
    """dram read+write bytes per launch of the dominant kernel — the fused gate_up GEMM at the
    bench shape (M=4608, N=28672, K=4096), `gemm2_bf16_kernel<256,1>` as dispatched there — from
    the newest committed `ncu --set full` capture under profiles/, or None"""
    import csv

    for fname, kernel in (("r2_ncu_full_targets.csv", "gemm2_bf16_kernel<256, 1>"),
                          ("r1_ncu_full_targets.csv", "gemm_bf16_kernel<256, 1>")):
        try:
            rows = list(csv.reader(open(os.path.join(ROOT, "profiles", fname))))
            hdr, units = rows[0], rows[1]
            ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            for r in rows[2:]:
                if kernel in r[ik]:
                    return (float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]],
                            f"ncu --set full (profiles/{fname}), {kernel} = fused gate_up GEMM at M=4608: algorithmic 405 MB")
        except Exception:
            continue
    return None, "no ncu capture found"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_tflops_burst": d["bf16_tflops"], "source": "MEASURED_PEAKS.json (sustained bf16, copy HBM)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (the oracle port), also used for the cpu_baseline key of the native arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(model_key: str, prompt_tokens: int, out_tokens: int, budget_s: float):
    """One bounded sample of the workload on the host cores: a real 128-token prefill plus as many
    greedy decode steps as fit in `budget_s`, extrapolated to a full 128-out job.  To bound host
    RAM and init time the 32 decoder layers share ONE set of random layer weights (same shapes,
    same FLOPs and bytes per layer; the outputs are not used for parity)."""
    import torch

    from llmq_b200.model import BUILTIN_SPECS

    spec = BUILTIN_SPECS[model_key]
    cores = usable_host_cores()
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(7)
    mat = lambda r, c: torch.randn(r, c, generator=g) * 0.02
    qd, kd = spec.n_q_heads * spec.head_dim, spec.n_kv_heads * spec.head_dim
    gemma = spec.arch == "gemma2"
    norm = lambda: torch.zeros(spec.hidden) if gemma else torch.ones(spec.hidden)  # (1+w) vs w
    norm_names = (["input_layernorm", "post_attention_layernorm", "pre_feedforward_layernorm",
                   "post_feedforward_layernorm"] if gemma else ["input_layernorm", "post_attention_layernorm"])
    layer = {f"{n}.weight": norm() for n in norm_names}
    layer.update({"self_attn.q_proj.weight": mat(qd, spec.hidden), "self_attn.k_proj.weight": mat(kd, spec.hidden),
                  "self_attn.v_proj.weight": mat(kd, spec.hidden), "self_attn.o_proj.weight": mat(spec.hidden, qd),
                  "mlp.gate_proj.weight": mat(spec.intermediate, spec.hidden),
                  "mlp.up_proj.weight": mat(spec.intermediate, spec.hidden),
                  "mlp.down_proj.weight": mat(spec.hidden, spec.intermediate)})
    w = {"model.embed_tokens.weight": mat(spec.vocab, spec.hidden), "model.norm.weight": norm()}
    if not spec.tie_embeddings:
        w["lm_head.weight"] = w["model.embed_tokens.weight"]  # same shape; shares storage
    for i in range(spec.n_layers):
        for k, v in layer.items():
            w[f"model.layers.{i}.{k}"] = v
    from oracle import ops as O
    if gemma:
        from oracle.gemma2 import Gemma2Dims, Gemma2Oracle
        dims = Gemma2Dims(hidden=spec.hidden, n_layers=spec.n_layers, n_q_heads=spec.n_q_heads,
                          n_kv_heads=spec.n_kv_heads, head_dim=spec.head_dim, intermediate=spec.intermediate,
                          vocab=spec.vocab, rms_eps=spec.rms_eps, rope_theta=spec.rope_theta,
                          query_pre_attn_scalar=spec.query_pre_attn_scalar or spec.head_dim,
                          attn_softcap=spec.attn_softcap or None, final_softcap=spec.final_softcap or None,
                          sliding_window=spec.sliding_window or 4096, max_pos=512)
        oracle = Gemma2Oracle.__new__(Gemma2Oracle)  # skip the per-tensor copies of __init__
        oracle.d, oracle.mode, oracle.w = dims, "fp32", w
        oracle.table = O.rope_table(512, dims.head_dim, dims.rope_theta, None, "fp32")
        oracle.scale = dims.query_pre_attn_scalar ** -0.5
    else:
        from oracle.model import LlamaDims, LlamaOracle
        dims = LlamaDims(hidden=spec.hidden, n_layers=spec.n_layers, n_q_heads=spec.n_q_heads,
                         n_kv_heads=spec.n_kv_heads, head_dim=spec.head_dim, intermediate=spec.intermediate,
                         vocab=spec.vocab, rms_eps=spec.rms_eps, rope_theta=spec.rope_theta,
                         rope_scaling=spec.rope_scaling, tie_embeddings=spec.tie_embeddings, max_pos=512)
        oracle = LlamaOracle.__new__(LlamaOracle)  # skip the per-tensor copies of __init__
        oracle.d, oracle.mode, oracle.w = dims, "fp32", w
        oracle.table = O.rope_table(512, dims.head_dim, dims.rope_theta, dims.rope_scaling, "fp32")
        oracle.scale = dims.head_dim ** -0.5
    ids = torch.randint(0, dims.vocab, (prompt_tokens,), generator=g)

    def sample():
        t0 = time.perf_counter()
        logits, kv = oracle.forward(ids, torch.arange(prompt_tokens), None, all_logits=False)
        t_prefill = time.perf_counter() - t0
        n_dec, t_dec = 0, 0.0
        tok = int(logits[-1].argmax())
        # at least 4 decode steps are always timed, so the extrapolation has a decode term
        while n_dec < out_tokens - 1 and (n_dec < 4 or (t_prefill + t_dec) < budget_s):
            t1 = time.perf_counter()
            logits, kv = oracle.forward(torch.tensor([tok]), torch.tensor([prompt_tokens + n_dec]), kv, all_logits=False)
            tok = int(logits[-1].argmax())
            t_dec += time.perf_counter() - t1
            n_dec += 1
        per_dec = t_dec / max(n_dec, 1)
        t_job = t_prefill + (out_tokens - 1) * per_dec
        return {"t_prefill_s": t_prefill, "decode_steps_timed": n_dec, "s_per_decode_token": per_dec,
                "s_per_job_extrapolated": t_job, "tokens_per_s": out_tokens / t_job, "jobs_per_s": 1.0 / t_job}

    return sample, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step is a bounded sample (a real prefill + as many decode steps as fit); the per-step
    # budget shrinks with --steps so that the whole run stays around two minutes of host time
    budget = max(2.0, min(args.cpu_budget_s, 100.0 / (args.steps + 1)))
    sample, cores = cpu_reference_sample(args.model, args.prompt_tokens, args.out_tokens, budget)
    if args.warmup > 0:
        sample()  # one warm-up pass faults in the weights; more would only burn host time
    t0 = time.perf_counter()
    res = [sample() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    tps = statistics.median(r["tokens_per_s"] for r in res)
    desc = (f"1 job per step on {cores} host threads (torch-CPU fp32 oracle, one batch row): real {args.prompt_tokens}-token prefill + "
            f"{res[0]['decode_steps_timed']} timed greedy decode steps, extrapolated to {args.out_tokens} output tokens; "
            "all decoder layers share one random layer's weights")
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / max(args.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (cpu)", "data": "synthetic",
            "config": workload_config(args),
            "jobs_per_sec": statistics.median(r["jobs_per_s"] for r in res),
            "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def builtin_weight_gb(model_key: str) -> float:
    from llmq_b200.model import BUILTIN_SPECS
    return BUILTIN_SPECS[model_key].weight_bytes_per_step() / 1e9


def workload_config(args):
    """the same dict on both arms (the reference arm runs a bounded sample of this workload)"""
    return {"workload": f"{args.model} random-init bf16, {args.prompt_tokens}-in/{args.out_tokens}-out greedy, continuous backlog "
                        f"(queue never empty), {args.max_num_seqs} concurrent sequences per GPU, step = {args.jobs} completed jobs "
                        "per GPU, queue-sharded (job i -> rank i % N)",
            "jobs_per_step_per_gpu": args.jobs,
            "max_num_seqs": args.max_num_seqs, "max_num_batched_tokens": args.max_num_batched_tokens,
            "gpu_memory_utilization": args.gpu_memory_utilization, "kv_block_size": 16, "l2": f"working set per engine step (weights {builtin_weight_gb(args.model):.0f} GB + KV) >> 126 MB L2, no flush needed",
            "parallelism": f"dp{args.gpus} (independent replicas, no collective)"}


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
class CudaTimer:
    """CUDA events on the engine's own stream (torch.cuda.Event only sees torch's current stream)"""

    def __init__(self, torch, stream_ptr):
        self.stream = torch.cuda.ExternalStream(stream_ptr)
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def start(self):
        self.e0.record(self.stream)

    def stop(self):
        self.e1.record(self.stream)
        self.e1.synchronize()

    def seconds(self):
        return self.e0.elapsed_time(self.e1) / 1e3


class HostTimer:
    def start(self):
        self.t0 = time.perf_counter()

    def stop(self):
        self.t1 = time.perf_counter()

    def seconds(self):
        return self.t1 - self.t0


def dry_run_service(args):
    """--dry-run: the real C++ scheduler / block manager without a model (tokens are fabricated) so
    that this script's own host logic (ramp, step accounting, windows, reductions) is testable on a
    machine without a GPU.  Its JSON line is marked and is never a measurement."""
    from llmq_b200.fixtures import DryRunEngine, build_tokenizer
    from llmq_b200.model import BUILTIN_SPECS
    from llmq_b200.service import GenerationService

    spec = BUILTIN_SPECS[args.model]
    eng = DryRunEngine(spec.vocab, max_num_seqs=args.max_num_seqs, max_num_batched_tokens=args.max_num_batched_tokens,
                       max_model_len=args.max_model_len, num_blocks=args.num_blocks or 4 * args.max_num_seqs * (args.max_model_len // 16),
                       eos_token_id=None)
    eng.model.spec = spec
    eng.model.set_profiling = lambda on: None
    eng.model.collect_profile = lambda reset=False: {n: {"ms": 0.0, "work": 0.0, "launches": 0} for n in
                                                     ("gemm", "decode_attn", "prefill_attn", "elementwise")}
    eng.stream_ptr = 0
    return GenerationService(eng, build_tokenizer(spec.vocab), None)


def staggered_caps(n: int, out_tokens: int):
    """generated-length caps of the first (ramp) cohort: uniform over 1..out_tokens, so that the
    sequences that fill the empty engine do not all finish on the same step — from then on the
    batch holds requests of every age, as under a long-running queue"""
    return [1 + (i * 7919) % out_tokens for i in range(n)]


def run_native(args):
    import numpy as np
    import torch

    from llmq_b200 import lib as L
    from llmq_b200.fixtures import make_jobs
    from llmq_b200.model import Engine
    from llmq_b200.service import GenerationService, build_service

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    dry = args.dry_run  # host-logic self-test of this script (tests/test_bench_host.py): no GPU, no model
    dist = None
    if not dry:
        torch.cuda.set_device(local)
        L.require_device()
    if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL's version banner goes there)
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        if dry:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red_dev = "cpu" if dry else "cuda"

    def barrier():
        if dist is not None:
            dist.barrier()
        if not dry:
            torch.cuda.synchronize()

    def reduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=red_dev)
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    peaks = load_peaks()
    if dry:
        svc = dry_run_service(args)
    else:
        svc = build_service(
            f"random:{args.model}", max_num_seqs=args.max_num_seqs, max_model_len=args.max_model_len,
            gpu_memory_utilization=args.gpu_memory_utilization, max_num_batched_tokens=args.max_num_batched_tokens,
            seed=1234 + rank, num_blocks=args.num_blocks)
    if args.gemm_mode:
        L.gemm_set_mode(args.gemm_mode)
    model, tok = svc.engine.model, svc.tokenizer
    spec = model.spec
    S, J, W, K = args.max_num_seqs, args.jobs, args.warmup, args.steps
    K_e = min(K, args.e2e_steps)
    # e2e tokens are counted on the host, per completed job: the warm-up has to flush the ramp cohort
    # (max_num_seqs jobs with staggered, i.e. shorter, outputs) before the clock starts
    W_e = 0 if W == 0 else max(2, -(-S // J))
    # canonical seeded job stream, this rank's shard (queue-sharding by job index).  Each arm needs:
    # the ramp cohort that fills the empty engine + its timed steps + a backlog that keeps the queue
    # non-empty until the clock stops (the worker is never starved: a 100k-prompt queue behind it)
    n_a = S + (W + K) * J + S + J
    n_b = S + (W_e + K_e) * J + S + J
    jobs = make_jobs((n_a + n_b) * world, spec.vocab, args.prompt_tokens - 1, start=rank, stride=world)
    jobs_a, jobs_b = jobs[:n_a], jobs[n_a:]
    caps = staggered_caps(S, args.out_tokens)

    # ---- arm A: engine-direct (prompts tokenised and queued inside the engine before the clock) ----
    eng = svc.engine
    timer = HostTimer() if dry else CudaTimer(torch, eng.stream_ptr)
    enc = svc.encode
    for i, j in enumerate(jobs_a):
        eng.add_request(i, enc(j["prompt"]), caps[i] if i < S else args.out_tokens, ignore_eos=True)
    PROFILE_EVERY = 8
    state = {"finished": 0, "tokens": 0, "k": 0, "profile": False}

    def run_until(n_finished):
        # per-launch CUDA-event profiling on every PROFILE_EVERY-th engine step only: the sampled
        # steps give the live roofline numbers, the others run exactly as in production
        # (CUDA-graph replay of decode steps included), so `value` is not perturbed
        while state["finished"] < n_finished:
            if state["profile"]:
                model.set_profiling(state["k"] % PROFILE_EVERY == 0)
            _, toks, flags = eng.step()
            state["tokens"] += len(toks)
            state["finished"] += int(np.count_nonzero(flags))
            state["k"] += 1

    run_until(W * J)                          # W untimed warm-up steps (the ramp is part of them)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    state["profile"] = True
    model.collect_profile(reset=True)
    st0, launches0, tok0, fin0 = eng.stats(), L.launch_count(), state["tokens"], state["finished"]
    barrier()
    timer.start()
    run_until(fin0 + K * J)                   # EXACTLY K steps of J completed jobs each
    timer.stop()
    barrier()
    t_a = reduce(timer.seconds(), "MAX")
    model.set_profiling(False)
    prof = model.collect_profile(reset=True)
    state["profile"] = False
    launches = L.launch_count() - launches0
    st1 = eng.stats()
    clocks = sampler.stop() if rank == 0 else None
    toks_a = state["tokens"] - tok0
    jobs_a_done = state["finished"] - fin0
    value = reduce(toks_a, "SUM") / t_a
    jobs_per_s = reduce(jobs_a_done, "SUM") / t_a
    waiting_left = st1.waiting
    eng.close()                               # drops the backlog that kept the queue non-empty

    # ---- arm B: end to end through the worker-facing call (host text in / text out) ----
    os.environ["VLLM_MAX_TOKENS"] = str(args.out_tokens)
    os.environ["B200Q_TEMPERATURE"] = "0"  # the workload is greedy (oracle mode); the worker's default is 0.7
    os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")
    ensure_llmq_importable()
    from llmq.core.models import Job

    from llmq_b200.worker import B200Worker

    if dry:
        svc.engine = eng = dry_run_service(args).engine
    else:
        svc.engine = eng = Engine(model, max_num_seqs=S, max_num_batched_tokens=args.max_num_batched_tokens,
                                  max_model_len=args.max_model_len, eos_token_id=svc.eos_token_id)
    worker = B200Worker(f"random:{args.model}", "bench-queue", tensor_parallel_size=1)
    worker.service = svc  # the engine built above (one model per process)
    for i in range(min(S, len(jobs_b))):
        jobs_b[i]["max_tokens"] = caps[i]     # ramp cohort: staggered caps through the per-job extra
    window = S + J                            # un-acked window, the role of VLLM_QUEUE_PREFETCH

    async def e2e_stream():
        # what BaseWorker does with a prefetch window: every delivered message is its own task
        # awaiting B200Worker._process_job(Job); a finished one lets the broker deliver the next
        loop = asyncio.get_running_loop()
        sem = asyncio.Semaphore(window)
        res = {"done": 0, "tok": 0, "t0": None, "t1": None, "tok0": 0, "sb0": None, "sb1": None}
        stop = asyncio.Event()
        if W_e == 0:
            res["t0"], res["sb0"] = time.perf_counter(), eng.stats()

        async def one(j):
            try:
                text = await worker._process_job(Job(**j))
            finally:
                sem.release()
            res["done"] += 1
            res["tok"] += len(text.split())   # one word per ordinary token of the synthetic vocab
            if res["done"] == W_e * J:
                res["t0"], res["tok0"], res["sb0"] = time.perf_counter(), res["tok"], eng.stats()
            elif res["done"] == (W_e + K_e) * J:
                res["t1"], res["sb1"] = time.perf_counter(), eng.stats()
                res["tok1"] = res["tok"]
                stop.set()

        tasks = []
        for j in jobs_b:
            await sem.acquire()
            if stop.is_set():
                break
            tasks.append(loop.create_task(one(j)))
        await stop.wait()
        for t in tasks:
            t.cancel()
        await asyncio.gather(*tasks, return_exceptions=True)
        # stop the engine thread while this loop is still open (it delivers results to the loop)
        await loop.run_in_executor(None, svc.stop)
        return res

    svc.start()
    barrier()
    res = asyncio.run(e2e_stream())
    svc.stop()
    if not dry:
        torch.cuda.synchronize()
    t_b = reduce(res["t1"] - res["t0"], "MAX")
    e2e_value = reduce(res["tok1"] - res["tok0"], "SUM") / t_b
    e2e_jobs = world * K_e * J / t_b
    sb0, sb1 = res["sb0"], res["sb1"]

    # ---------------- rooflines ----------------
    g = prof["gemm"]
    gemm_tflops = g["work"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
    d = prof["decode_attn"]
    dec_gbs = d["work"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
    dev_ms = sum(v["ms"] for v in prof.values())
    shares = {k: round(v["ms"] / dev_ms, 4) if dev_ms else 0 for k, v in prof.items()}
    roofline = {"bound": "tensor", "kernel": "b200q::gemm2_bf16_kernel<256,*> (tcgen05 cta_group::2, the variant dispatched at M=4608) / gemm_bf16_kernel",
                "achieved": round(gemm_tflops, 1),
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": round(gemm_tflops / peaks["bf16_tflops"], 4),
                "traffic": None, "peak_source": peaks["source"],
                "launch_avg_ms": round(g["ms"] / max(g["launches"], 1), 5), "launches": g["launches"],
                "share_of_device_time": shares}
    if args.model == "llama-3-8b":  # the committed ncu capture is of this model's fused gate_up GEMM
        roofline["traffic"], roofline["traffic_note"] = ncu_traffic()
    roofline_dec = {"bound": "hbm", "kernel": f"b200q::decode_attn_stream_kernel<{spec.head_dim},16> (B*n_kv >= 16*SMs) / decode_attn_kernel", "achieved": round(dec_gbs, 1),
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(dec_gbs / peaks["hbm_gbs"], 4),
                    "launch_avg_ms": round(d["ms"] / max(d["launches"], 1), 5), "launches": d["launches"]}
    steps_a = st1.steps - st0.steps
    dec_tokens = st1.tokens_decoded - st0.tokens_decoded
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            sample, cores = cpu_reference_sample(args.model, args.prompt_tokens, args.out_tokens, args.cpu_budget_s)
            r = sample()
            cpu = {"value": r["tokens_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                   "jobs_per_sec": r["jobs_per_s"],
                   "sample": f"1 job: real {args.prompt_tokens}-token prefill ({r['t_prefill_s']:.2f} s) + {r['decode_steps_timed']} timed decode "
                             f"steps ({r['s_per_decode_token'] * 1e3:.1f} ms/token) extrapolated to {args.out_tokens} out tokens; torch-CPU fp32 "
                             "oracle, all host threads, all decoder layers share one random layer's weights"}
        e2e_eng_steps = max(sb1.steps - sb0.steps, 1)
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": round(t_a / K * 1e3, 2), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args), "jobs_per_sec": round(jobs_per_s, 2),
                "e2e": {"value": round(e2e_value, 1), "unit": UNIT,
                        "jobs_per_sec": round(e2e_jobs, 2), "steps": K_e, "warmup": W_e,
                        "h2d_bytes_per_step": int((sb1.h2d_bytes - sb0.h2d_bytes) / K_e),
                        "d2h_bytes_per_step": int((sb1.d2h_bytes - sb0.d2h_bytes) / K_e),
                        "engine_steps_per_step": round(e2e_eng_steps / K_e, 1),
                        "note": "B200Worker._process_job(Job) on host text behind an un-acked window of max_num_seqs + J jobs "
                                "(BaseWorker's prefetch): format prompt -> tokenise -> engine thread (H2D metadata / D2H ids every "
                                "engine step) -> detokenised text"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "roofline_decode_attn": roofline_dec,
                "engine": {"engine_steps_per_bench_step": steps_a / K, "decode_tokens": int(dec_tokens),
                           "prefill_tokens": int(st1.tokens_prefilled - st0.tokens_prefilled),
                           "preemptions": int(st1.preemptions - st0.preemptions), "kv_blocks": int(st1.total_blocks),
                           "running_at_end": int(st1.running), "backlog_at_end": int(waiting_left),
                           "device_ms_profiled": round(dev_ms, 1), "profiled_engine_steps": f"1 of every {PROFILE_EVERY}",
                           "wall_ms_timed": round(t_a * 1e3, 1)},
                "cpu_baseline": cpu}
        if dry:
            line["dry_run"] = "host-logic self-test: C++ scheduler without a model, NOT a measurement"
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    model.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    # 4608 = 36 x 128 rows: with 148 SMs every projection's tile count is within 3 % of a whole
    # number of rounds of the persistent GEMM loop; the KV pool holds them all
    ap.add_argument("--jobs", type=int, default=1152,
                    help="completed jobs per step per GPU (a quarter of the 4608 concurrent sequences)")
    ap.add_argument("--e2e-steps", type=int, default=8, help="the e2e arm times min(--steps, this) steps")
    ap.add_argument("--num-blocks", type=int, default=None, help="KV pool size in 16-token blocks (default: from gpu_memory_utilization)")
    ap.add_argument("--gemm-mode", type=int, default=0, help="tuning hook: 0 auto, 1 1-CTA kernels, 2 2-CTA kernel")
    ap.add_argument("--prompt-tokens", type=int, default=128)
    ap.add_argument("--out-tokens", type=int, default=128)
    ap.add_argument("--max-num-seqs", type=int, default=4608)
    ap.add_argument("--gpu-memory-utilization", type=float, default=0.92)
    ap.add_argument("--max-num-batched-tokens", type=int, default=4608,
                    help="36 x 128 like the decode batch: A operand stays L2-resident (measured +1 %% over 9472)")
    ap.add_argument("--max-model-len", type=int, default=512)
    ap.add_argument("--cpu-budget-s", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dry-run", action="store_true", help="host-logic self-test without a GPU (not a measurement)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
