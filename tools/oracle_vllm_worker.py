"""OracleVLLMWorker — the reference's own worker as the GPU oracle.

The reference worker (ref:llmq/workers/vllm_worker.py:11) is run UNMODIFIED on a B200 behind the
unmodified `BaseWorker` / `BrokerManager` (on the aio_pika stand-in: the image has no RabbitMQ);
this subclass only makes it usable as a parity oracle, without touching /root/reference:

  * :146 `await self.engine.get_tokenizer()` — vLLM >= 0.22 returns the tokenizer synchronously, so
    the engine's `get_tokenizer` is wrapped into an awaitable;
  * :161-165 `SamplingParams(temperature=0.7, ...)` — the module-level name `SamplingParams` is
    replaced by a factory that forces `temperature=0.0` (greedy, the oracle mode) and asks for the
    top-2 logprobs of every generated token (their difference is the logit margin of that step,
    which is what decides whether a divergence is a near-tie);
  * :183-193 — `engine.generate` is wrapped to capture `outputs[0].token_ids` (and the prompt ids
    vLLM tokenised) beside the text the worker returns.

  python tools/oracle_vllm_worker.py golden [case ...]      (on the B200 box)
      -> gpurun_out/vllm_worker_golden_<case>.json
  python tools/oracle_vllm_worker.py pack                   (here)
      -> tests/golden/vllm_worker_golden_<case>.json.gz     (committed fixtures)

Cases are seeded random-init checkpoints with the REAL widths of the BASELINE models
(`llmq_b200.fixtures.seeded_state_dict` reproduces them bit for bit anywhere):
  llama3_8b_w4  : Llama-3-8B hidden/intermediate/heads/vocab, 4 of the 32 layers
  llama32_1b    : Llama-3.2-1B, all 16 layers (tied embeddings, llama3 rope scaling)
Each case runs the same JSONL (64 canonical 128-token jobs + ragged lengths) through the broker
twice: vLLM as the reference configures it (compiled + CUDA graphs) and with `enforce_eager=True`,
so the file also holds how far vLLM disagrees with itself — the honest floor for "bit-exact".
"""
from __future__ import annotations

import asyncio
import gc
import hashlib
import inspect
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
for _p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(_p, "llmq")):
        sys.path.insert(0, _p)
        break
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")
OUT = os.path.join(ROOT, "gpurun_out")

from llmq_b200.model import LLAMA_32_1B, LLAMA_3_8B, ModelSpec  # noqa: E402

GOLDEN_SEED = 20260921
MAX_NEW = 128


def case_spec(name: str) -> ModelSpec:
    import dataclasses

    if name == "llama3_8b_w4":
        return dataclasses.replace(LLAMA_3_8B, n_layers=4, name="llama-3-8b-width-4layers", max_position_embeddings=2048)
    if name == "llama32_1b":
        return dataclasses.replace(LLAMA_32_1B, name="llama-3.2-1b-seeded")
    raise KeyError(name)


def golden_jobs(vocab: int):
    """64 jobs of the canonical benchmark stream (127 words -> 128 ids with BOS) + ragged lengths"""
    from llmq_b200.fixtures import make_jobs

    jobs = make_jobs(64, vocab, 127)
    for k, n in enumerate((1, 2, 15, 16, 17, 31, 33, 63, 200, 300, 511, 700)):
        j = make_jobs(1, vocab, n, seed=777 + k)[0]
        j["id"] = f"ragged-{n:04d}"
        jobs.append(j)
    return jobs


def make_oracle_worker_class(extra_engine_args: dict):
    """the subclass is built lazily: importing llmq.workers.vllm_worker imports vLLM"""
    import llmq.workers.vllm_worker as VW
    import vllm

    real_sampling_params = vllm.SamplingParams
    real_engine_args = VW.AsyncEngineArgs

    def greedy_sampling_params(**kw):
        kw["temperature"] = 0.0     # the literal 0.7 at vllm_worker.py:162
        kw["logprobs"] = 2          # top-2 logprobs: margin of every decision
        return real_sampling_params(**kw)

    def engine_args(**kw):
        kw.update(extra_engine_args)  # the reference exposes neither enforce_eager nor seed
        return real_engine_args(**kw)

    class OracleVLLMWorker(VW.VLLMWorker):
        """ref:llmq/workers/vllm_worker.py:11, instrumented (see module docstring)"""

        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            self.captured = {}

        async def _initialize_processor(self) -> None:
            VW.SamplingParams = greedy_sampling_params
            VW.AsyncEngineArgs = engine_args
            try:
                await super()._initialize_processor()
            finally:
                VW.AsyncEngineArgs = real_engine_args
            eng = self.engine
            sync_get_tokenizer = eng.get_tokenizer

            async def get_tokenizer(*a, **kw):
                t = sync_get_tokenizer(*a, **kw)
                return await t if inspect.isawaitable(t) else t

            eng.get_tokenizer = get_tokenizer
            orig_generate = eng.generate
            captured = self.captured

            async def generate(prompt, sampling_params, request_id, **kw):
                last = None
                async for out in orig_generate(prompt, sampling_params, request_id=request_id, **kw):
                    last = out
                    yield out
                if last is not None and last.outputs:
                    o = last.outputs[0]
                    margins, seconds = [], []
                    for tok, lp in zip(o.token_ids, o.logprobs or []):
                        vals = sorted((v.logprob for v in lp.values()), reverse=True)
                        margins.append(round(vals[0] - vals[1], 5) if len(vals) > 1 else None)
                        alt = [t for t in lp if t != tok]
                        seconds.append(int(alt[0]) if alt else -1)
                    captured[request_id] = {"prompt_ids": list(last.prompt_token_ids or []), "ids": list(o.token_ids),
                                            "text": o.text, "margins": margins, "runner_up": seconds,
                                            "finish_reason": o.finish_reason}

            eng.generate = generate

        async def _cleanup_processor(self) -> None:
            import llmq.workers.vllm_worker as VW2
            VW2.SamplingParams = real_sampling_params
            eng = self.engine
            await super()._cleanup_processor()
            if eng is not None and hasattr(eng, "shutdown"):
                eng.shutdown()

    return OracleVLLMWorker


async def run_through_broker(worker, queue: str, jobs, timeout_s: float = 900.0):
    """jobs in through the reference's BrokerManager, results out of `<queue>.results`"""
    from llmq.core.broker import BrokerManager
    from llmq.core.models import Job, Result

    task = asyncio.create_task(worker.run())
    b = BrokerManager()
    await b.connect()
    await b.setup_queue_infrastructure(queue)
    for j in jobs:
        await b.publish_job(queue, Job(**j))
    got = {}

    async def on_res(m):
        r = Result.parse_raw(m.body)
        got[r.id] = r.result
        await m.ack()

    await b.consume_results(queue, on_res)
    t0 = time.time()
    while len(got) < len(jobs) and time.time() - t0 < timeout_s and not task.done():
        await asyncio.sleep(0.1)
    worker.running = False
    await asyncio.wait_for(task, 120)
    return got


def golden(cases):
    import aio_pika
    import torch
    import vllm

    from llmq_b200.fixtures import write_model_dir

    os.makedirs(OUT, exist_ok=True)
    os.environ["VLLM_MAX_TOKENS"] = str(MAX_NEW)
    os.environ["VLLM_MAX_NUM_SEQS"] = "128"
    os.environ["VLLM_QUEUE_PREFETCH"] = "256"
    os.environ["VLLM_GPU_MEMORY_UTILIZATION"] = "0.5"
    os.environ["VLLM_MAX_MODEL_LEN"] = "1024"
    for name in cases:
        spec = case_spec(name)
        mdir = os.path.join("/tmp", f"b200q_wgolden_{name}")
        t0 = time.time()
        write_model_dir(mdir, spec, seed=GOLDEN_SEED, with_weights=True)
        print(f"[{name}] checkpoint written in {time.time() - t0:.0f} s", flush=True)
        jobs = golden_jobs(spec.vocab)
        res = {"case": name, "spec": spec.to_hf_config(), "weights_seed": GOLDEN_SEED, "max_new_tokens": MAX_NEW,
               "vllm": vllm.__version__, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
               "harness": "tools/oracle_vllm_worker.py: OracleVLLMWorker(VLLMWorker) behind BaseWorker/BrokerManager",
               "jobs": jobs, "runs": {}}
        for tag, extra in (("default", {"seed": 0}), ("eager", {"seed": 0, "enforce_eager": True})):
            aio_pika.reset_brokers()
            try:
                cls = make_oracle_worker_class(extra)
                w = cls(mdir, f"wg-{name}-{tag}", tensor_parallel_size=1)
                t0 = time.time()
                got = asyncio.run(run_through_broker(w, f"wg-{name}-{tag}", jobs))
                run = {}
                for n_j, j in enumerate(jobs):
                    c = w.captured.get(j["id"])
                    if c is None or j["id"] not in got:
                        continue
                    assert got[j["id"]] == c["text"], "the worker must return exactly the captured text"
                    # keep the fixture small: full text for the first 8 jobs, a digest for the rest;
                    # the prompt ids vLLM tokenised only once (they do not depend on the engine mode)
                    c["text_sha1"] = hashlib.sha1(c["text"].encode()).hexdigest()
                    if n_j >= 8:
                        del c["text"]
                    if tag != "default":
                        del c["prompt_ids"]
                    run[j["id"]] = c
                res["runs"][tag] = run
                print(f"[{name}/{tag}] {len(run)}/{len(jobs)} jobs in {time.time() - t0:.0f} s", flush=True)
                del w
            except Exception as e:  # keep going: one mode is enough to pin
                import traceback

                res["runs"][tag] = {"error": repr(e)[:500], "trace": traceback.format_exc()[-1500:]}
                print(f"[{name}/{tag}] FAILED: {e!r}", flush=True)
            gc.collect()
            torch.cuda.empty_cache()
        with open(os.path.join(OUT, f"vllm_worker_golden_{name}.json"), "w") as f:
            json.dump(res, f, separators=(",", ":"))


def pack():
    """gpurun_out/vllm_worker_golden_<case>.json -> tests/golden/vllm_worker_golden_<case>.json.gz"""
    import glob
    import gzip

    for src in sorted(glob.glob(os.path.join(OUT, "vllm_worker_golden_*.json"))):
        dst = os.path.join(ROOT, "tests", "golden", os.path.basename(src) + ".gz")
        with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
            g.write(f.read())
        print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "pack":
        pack()
    elif len(sys.argv) >= 2 and sys.argv[1] == "golden":
        golden(sys.argv[2:] or ["llama32_1b", "llama3_8b_w4"])
    else:
        raise SystemExit(__doc__)
