"""Process-level benchmark with the reference's measurement shape (ref:performance_benchmark.py:193-394,
BASELINE.md §2): a broker process, N worker PROCESSES started through the CLI
(`python -m llmq_b200.cli worker run MODEL QUEUE`, one per GPU), the queue pre-loaded with the
synthetic JSONL jobs, clock from first delivery to last result in `<queue>.results`.

    python tools/queue_bench.py --workers 1 --jobs 4608

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


async def run(args, port):
    os.environ["B200Q_SHIM_BROKER"] = f"127.0.0.1:{port}"
    os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")
    from llmq.core.broker import BrokerManager
    from llmq.core.models import Job, Result

    from llmq_b200.fixtures import make_jobs

    q = "qbench"
    b = BrokerManager()
    await b.connect()
    await b.setup_queue_infrastructure(q)
    jobs = make_jobs(args.jobs, 128256, prompt_tokens=args.prompt_tokens - 1)
    for j in jobs:
        await b.publish_job(q, Job(**j))
    # workers
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, REF, ROOT]), VLLM_MAX_TOKENS=str(args.out_tokens),
               VLLM_MAX_NUM_SEQS=str(args.max_num_seqs), VLLM_QUEUE_PREFETCH=str(args.prefetch),
               VLLM_MAX_MODEL_LEN="512", VLLM_GPU_MEMORY_UTILIZATION="0.92", B200Q_TEMPERATURE="0",
               B200Q_MAX_NUM_BATCHED_TOKENS=str(args.budget), LLMQ_LOG_LEVEL="WARNING")
    procs = []
    for i in range(args.workers):
        e = dict(env, CUDA_VISIBLE_DEVICES=str(i))
        procs.append(subprocess.Popen([sys.executable, "-m", "llmq_b200.cli", "worker", "run", args.model, q, "-tp", "1"],
                                      env=e, stdout=open(os.path.join(args.out_dir, f"qbench_worker{i}.log"), "w"),
                                      stderr=subprocess.STDOUT))
    got, toks, t_first, t_last = 0, 0, None, None
    workers_seen = set()

    async def on_res(m):
        nonlocal got, toks, t_first, t_last
        r = Result.parse_raw(m.body)
        got += 1
        toks += len(r.result.split())
        workers_seen.add(r.worker_id)
        t_last = time.perf_counter()
        await m.ack()

    await b.consume_results(q, on_res)
    conn = b.connection
    t_start = time.perf_counter()
    while got < args.jobs and time.perf_counter() - t_start < args.timeout:
        if t_first is None:
            st = await conn._request({"op": "stats"})
            if st["queues"].get(q, {}).get("delivered", 0) > 0:
                t_first = time.perf_counter()  # first delivery to a worker = the clock starts
        for p in procs:
            if p.poll() is not None:
                raise RuntimeError(f"worker exited with {p.returncode}")
        await asyncio.sleep(0.02)
    dt = (t_last - t_first) if (t_first and t_last) else float("nan")
    for p in procs:
        p.terminate()
    await b.disconnect()
    print(json.dumps({"kind": "queue_bench", "workers": args.workers, "jobs": got, "of": args.jobs, "seconds": dt,
                      "jobs_per_s": got / dt, "out_tokens_per_s": toks / dt, "model": args.model,
                      "max_num_seqs": args.max_num_seqs, "prefetch": args.prefetch, "worker_ids": sorted(workers_seen),
                      "startup_s": (t_first - t_start) if t_first else None,
                      "path": "broker process (TCP aio_pika stand-in) -> `python -m llmq_b200.cli worker run` processes "
                              "(reference BaseWorker/BrokerManager unmodified) -> <queue>.results"}))
    for p in procs:
        try:
            p.wait(20)
        except Exception:
            p.kill()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=4608)
    ap.add_argument("--model", default="random:llama-3-8b")
    ap.add_argument("--prompt-tokens", type=int, default=128)
    ap.add_argument("--out-tokens", type=int, default=128)
    ap.add_argument("--max-num-seqs", type=int, default=4608)
    ap.add_argument("--prefetch", type=int, default=5000)
    ap.add_argument("--budget", type=int, default=4608)
    ap.add_argument("--timeout", type=float, default=400)
    ap.add_argument("--out-dir", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
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
