"""Host-path ceiling of the b200 worker, measured WITHOUT a GPU.

The C++ engine runs in its dry-run mode (real scheduler, real paged-KV bookkeeping and per-step
metadata packing, fabricated tokens, no CUDA), so what is timed is everything the host does per
job around the GPU work: prompt formatting, tokenisation, the engine thread's bookkeeping,
detokenisation, future delivery — and, with --level broker, the reference's unmodified
BaseWorker / BrokerManager / pydantic Job+Result path on the in-process aio_pika stand-in.
If this number is not far above the GPU's jobs/s the host is the bottleneck.

    python tools/host_path_bench.py --jobs 20000 --level service
    python tools/host_path_bench.py --jobs 5000 --level broker
"""
import argparse
import asyncio
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(p, "llmq")):
        sys.path.insert(0, p)
        break
os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=20000)
    ap.add_argument("--level", choices=["service", "broker"], default="service")
    ap.add_argument("--prompt-tokens", type=int, default=128)
    ap.add_argument("--out-tokens", type=int, default=128)
    ap.add_argument("--max-num-seqs", type=int, default=4608)
    ap.add_argument("--vocab", type=int, default=128256)
    ap.add_argument("--inflight", type=int, default=5000, help="jobs in flight (= VLLM_QUEUE_PREFETCH)")
    ap.add_argument("--step-ms", type=float, default=0.0,
                    help="simulated device time per engine step (0 = measure the bare host ceiling)")
    args = ap.parse_args()

    import llmq_b200.worker as W
    from llmq_b200 import service as S
    from llmq_b200.fixtures import DryRunEngine, build_tokenizer, make_jobs, special_token_ids
    from llmq.core.models import Job

    tok = build_tokenizer(args.vocab)
    eos = special_token_ids(args.vocab)["<|end_of_text|>"]
    made = {}

    def dry_build_service(model_name, **kw):
        blocks = args.max_num_seqs * ((args.prompt_tokens + args.out_tokens + 15) // 16 + 1)
        eng = DryRunEngine(args.vocab, args.max_num_seqs, 4608, 512, blocks, eos, step_ms=args.step_ms)
        made["engine"] = eng
        return S.GenerationService(eng, tok, eos)

    W.build_service = dry_build_service
    os.environ["VLLM_MAX_TOKENS"] = str(args.out_tokens)
    os.environ["VLLM_MAX_NUM_SEQS"] = str(args.max_num_seqs)
    os.environ["VLLM_QUEUE_PREFETCH"] = str(args.inflight)
    os.environ["B200Q_TEMPERATURE"] = "0"
    raw = make_jobs(args.jobs, args.vocab, args.prompt_tokens - 1)

    async def service_level():
        w = W.B200Worker("random:dry", "hq", tensor_parallel_size=1)
        await w._initialize_processor()
        jobs = [Job(**j) for j in raw]
        chars = 0
        it = iter(jobs)

        async def consumer():  # args.inflight of these = the prefetch window of a real worker
            nonlocal chars
            for j in it:
                text = await w._process_job(j)
                chars += len(text)

        t0 = time.perf_counter()
        await asyncio.gather(*[consumer() for _ in range(min(args.inflight, len(jobs)))])
        dt = time.perf_counter() - t0
        await w._cleanup_processor()
        return dt, chars

    async def broker_level():
        import aio_pika
        from llmq.core.broker import BrokerManager

        aio_pika.reset_brokers()
        w = W.B200Worker("random:dry", "hq", tensor_parallel_size=1)
        task = asyncio.create_task(w.run())
        b = BrokerManager()
        await b.connect()
        await b.setup_queue_infrastructure("hq")
        for j in raw:
            await b.publish_job("hq", Job(**j))
        done = asyncio.Event()
        got = 0
        chars = 0

        async def on_result(msg):
            nonlocal got, chars
            chars += len(json.loads(msg.body)["result"])
            got += 1
            await msg.ack()
            if got == args.jobs:
                done.set()

        t0 = time.perf_counter()
        await b.consume_results("hq", on_result)
        await done.wait()
        dt = time.perf_counter() - t0
        w.running = False
        await asyncio.wait_for(task, 30)
        await b.disconnect()
        return dt, chars

    cpu0 = time.process_time()
    dt, chars = asyncio.run(service_level() if args.level == "service" else broker_level())
    cpu = time.process_time() - cpu0  # all threads, user + system: less noisy than wall time
    eng = made["engine"]
    print(json.dumps({"level": args.level, "jobs": args.jobs, "seconds": round(dt, 3),
                      "jobs_per_sec": round(args.jobs / dt, 1),
                      "output_tokens_per_sec": round(args.jobs * args.out_tokens / dt, 1),
                      "engine_steps": eng.steps, "host_us_per_job": round(dt / args.jobs * 1e6, 1),
                      "cpu_us_per_job": round(cpu / args.jobs * 1e6, 1),
                      "result_chars": chars, "host_cores": os.cpu_count(), "step_ms": args.step_ms,
                      "device_only_jobs_per_sec": (round(args.jobs / (eng.steps * args.step_ms / 1e3), 1)
                                                   if args.step_ms else None),
                      "note": "dry-run C++ engine (no GPU work): the host path's own ceiling"}))


if __name__ == "__main__":
    main()
