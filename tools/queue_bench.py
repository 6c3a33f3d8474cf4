"""Process-level benchmark with the reference's measurement shape (ref:performance_benchmark.py:193-394,
BASELINE.md §2): a broker process, N worker PROCESSES started through the CLI (one per GPU), the
queue pre-loaded with synthetic JSONL jobs, clock from first delivery to last result.

  config #3 (queue-sharded data parallel, ref:utils/run_llmq_benchmark.slurm:52-72):
    python tools/queue_bench.py --workers 8 --jobs 100000
  config #4 (two-stage pipeline, ref:llmq/core/broker.py:145-193, example-pipeline.yaml): N1 workers on
  stage 1, N2 on stage 2, every stage-1 result is routed to the stage-2 queue by the reference's
  unmodified `publish_pipeline_result`:
    python tools/queue_bench.py --pipeline 4+4 --model random:gemma-2-9b --jobs 50000
  config #5 (continuous-batching stress): prompt lengths ~U{32..2048}, output lengths ~U{1..512} through
  the per-job `max_tokens` extra (the reference only has the global VLLM_MAX_TOKENS):
    python tools/queue_bench.py --workers 8 --mixed --model random:gemma-2-9b --jobs 20000

Workers are `python -m llmq_b200.cli worker run MODEL QUEUE` / `... worker pipeline CFG STAGE`: the
reference's click CLI, BaseWorker and BrokerManager unmodified, the native worker in the vLLM slot.
`--worker-kind dummy` swaps in the reference's DummyWorker (CPU smoke test of this tool).
Broker: the TCP server of the aio_pika stand-in (no RabbitMQ in this image).  Output: one JSON line.
"""
import argparse
import asyncio
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "tests", "shims")
REF = next((p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference") if os.path.isdir(os.path.join(p, "llmq"))), None)
sys.path[:0] = [SHIMS, REF, ROOT]


def build_jobs(args):
    """the canonical seeded stream; --mixed draws a prompt length and an output cap per job"""
    import numpy as np

    from llmq_b200.fixtures import make_jobs, random_prompt_words

    if not args.mixed:
        return make_jobs(args.jobs, args.vocab, prompt_tokens=args.prompt_tokens - 1), args.jobs * args.prompt_tokens
    jobs, n_in = [], 0
    for i in range(args.jobs):
        rng = np.random.default_rng([20260921, 5, i])
        n_prompt = int(rng.integers(32, 2049))
        jobs.append({"id": f"mix-{i:07d}", "prompt": random_prompt_words(rng, n_prompt - 1, args.vocab),
                     "max_tokens": int(rng.integers(1, 513))})
        n_in += n_prompt
    return jobs, n_in


async def run(args, port):
    os.environ["B200Q_SHIM_BROKER"] = f"127.0.0.1:{port}"
    os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")
    from llmq.core.broker import BrokerManager
    from llmq.core.models import Job, Result

    b = BrokerManager()
    await b.connect()
    stages = []
    if args.pipeline:
        n1, n2 = (int(x) for x in args.pipeline.split("+"))
        kind = "dummy" if args.worker_kind == "dummy" else "vllm"  # `vllm` = the slot the native worker sits in
        cfg = {"name": "qbench", "stages": [
            {"name": "stage1", "worker": kind, "config": {"model": args.model}},
            {"name": "stage2", "worker": kind, "config": {"model": args.stage2_model or args.model}}]}
        cfg_path = os.path.join(args.out_dir, "qbench_pipeline.yaml")
        import yaml
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f)
        from llmq.core.pipeline import PipelineConfig
        pc = PipelineConfig(**cfg)
        await b.setup_pipeline_infrastructure(pc.name, [st.name for st in pc.stages])
        q_in, q_out = pc.get_stage_queue_name("stage1"), pc.get_pipeline_results_queue_name()
        stages = [("stage1", n1), ("stage2", n2)]
        n_workers = n1 + n2
    else:
        q_in, q_out = "qbench", "qbench.results"
        await b.setup_queue_infrastructure(q_in)
        n_workers = args.workers
    jobs, n_prompt_tokens = build_jobs(args)
    t0 = time.perf_counter()
    for j in jobs:
        await b.publish_job(q_in, Job(**j))
    t_publish = time.perf_counter() - t0
    # workers
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, REF, ROOT]), VLLM_MAX_TOKENS=str(args.out_tokens),
               VLLM_MAX_NUM_SEQS=str(args.max_num_seqs), VLLM_QUEUE_PREFETCH=str(args.prefetch),
               VLLM_MAX_MODEL_LEN=str(args.max_model_len), VLLM_GPU_MEMORY_UTILIZATION="0.92", B200Q_TEMPERATURE="0",
               B200Q_MAX_NUM_BATCHED_TOKENS=str(args.budget), LLMQ_LOG_LEVEL="WARNING")
    procs = []
    cli = [sys.executable, "-m", "llmq_b200.cli", "worker"]
    plan = []  # (gpu index, argv tail)
    if stages:
        g = 0
        for name, n in stages:
            for _ in range(n):
                plan.append((g, ["pipeline", cfg_path, name] + (["-c", str(args.prefetch)] if args.worker_kind == "dummy" else [])))
                g += 1
    elif args.worker_kind == "dummy":
        plan = [(i, ["dummy", q_in, "-c", str(args.prefetch)]) for i in range(n_workers)]
    else:
        plan = [(i, ["run", args.model, q_in, "-tp", "1"]) for i in range(n_workers)]
    for i, (gpu, tail) in enumerate(plan):
        e = dict(env, CUDA_VISIBLE_DEVICES=str(gpu))
        procs.append(subprocess.Popen(cli + tail, env=e, stderr=subprocess.STDOUT,
                                      stdout=open(os.path.join(args.out_dir, f"qbench_worker{i}.log"), "w")))
    got, toks, t_first, t_last = 0, 0, None, None
    workers_seen = set()

    async def on_res(m):
        nonlocal got, toks, t_last
        r = Result.parse_raw(m.body)
        got += 1
        toks += len(r.result.split())
        workers_seen.add(r.worker_id)
        t_last = time.perf_counter()
        await m.ack()

    q = await b.channel.declare_queue(q_out, durable=True)
    await q.consume(on_res)
    conn = b.connection
    t_start = time.perf_counter()
    stats = {}
    while got < args.jobs and time.perf_counter() - t_start < args.timeout:
        stats = (await conn._request({"op": "stats"}))["queues"]
        if t_first is None and stats.get(q_in, {}).get("delivered", 0) > 0:
            t_first = time.perf_counter()  # first delivery to a worker = the clock starts
        for p in procs:
            if p.poll() is not None:
                raise RuntimeError(f"worker exited with {p.returncode}: see {args.out_dir}/qbench_worker*.log")
        await asyncio.sleep(0.02)
    dt = (t_last - t_first) if (t_first and t_last) else float("nan")
    for p in procs:
        p.terminate()
    await b.disconnect()
    n_stages = 2 if stages else 1
    print(json.dumps({
        "kind": "queue_bench", "config": "#4 two-stage pipeline" if stages else ("#5 mixed lengths" if args.mixed else "#3 queue-sharded"),
        "workers": args.pipeline or n_workers, "jobs": got, "of": args.jobs, "seconds": dt,
        "jobs_per_s": got / dt, "final_stage_out_tokens_per_s": toks / dt,
        "model_calls_per_s": got * n_stages / dt, "prompt_tokens_submitted": n_prompt_tokens,
        "model": args.model, "max_num_seqs": args.max_num_seqs, "prefetch": args.prefetch, "budget": args.budget,
        "worker_ids": sorted(workers_seen), "startup_s": (t_first - t_start) if t_first else None,
        "publish_s": round(t_publish, 2), "queues_at_end": stats,
        "path": "broker process (TCP aio_pika stand-in) -> `python -m llmq_b200.cli worker ...` processes "
                "(reference click CLI / BaseWorker / BrokerManager unmodified) -> results queue"}))
    for p in procs:
        try:
            p.wait(20)
        except Exception:
            p.kill()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--pipeline", default="", help="N1+N2: two-stage pipeline with N1 / N2 worker processes")
    ap.add_argument("--mixed", action="store_true", help="config #5: prompts U{32..2048} tokens, outputs U{1..512}")
    ap.add_argument("--worker-kind", default="b200", choices=["b200", "dummy"])
    ap.add_argument("--jobs", type=int, default=4608)
    ap.add_argument("--model", default="random:llama-3-8b")
    ap.add_argument("--stage2-model", default="")
    ap.add_argument("--vocab", type=int, default=0, help="vocabulary the synthetic prompts are drawn from (default: the model's)")
    ap.add_argument("--prompt-tokens", type=int, default=128)
    ap.add_argument("--out-tokens", type=int, default=128, help="VLLM_MAX_TOKENS (the global cap)")
    ap.add_argument("--max-num-seqs", type=int, default=4608)
    ap.add_argument("--max-model-len", type=int, default=512)
    ap.add_argument("--prefetch", type=int, default=5000)
    ap.add_argument("--budget", type=int, default=4608)
    ap.add_argument("--timeout", type=float, default=600)
    ap.add_argument("--out-dir", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    if not args.vocab:
        from llmq_b200.model import BUILTIN_SPECS
        key = args.model.split(":", 1)[1] if args.model.startswith("random:") else ""
        args.vocab = BUILTIN_SPECS[key].vocab if key in BUILTIN_SPECS else 128256
    if args.mixed:
        args.out_tokens = max(args.out_tokens, 512)
        args.max_model_len = max(args.max_model_len, 2048 + 512 + 16)
    os.makedirs(args.out_dir, exist_ok=True)
    ready = tempfile.mktemp()
    env = dict(os.environ, PYTHONPATH=SHIMS)
    srv = subprocess.Popen([sys.executable, "-m", "aio_pika.server", "--port", "0", "--ready-file", ready], env=env)
    try:
        for _ in range(200):
            if os.path.exists(ready) and open(ready).read():
                break
            time.sleep(0.05)
        asyncio.run(run(args, open(ready).read()))
    finally:
        srv.terminate()


if __name__ == "__main__":
    main()
