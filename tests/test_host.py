"""Host-side logic of the worker slot, CPU only: the reference's BaseWorker / BrokerManager /
Job / Result code runs UNMODIFIED (from baseline/_ref or /root/reference) on the aio_pika
stand-in, with B200Worker plugged in and a fake engine in place of the GPU."""
import asyncio
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(p, "llmq")):
        sys.path.insert(0, p)
        break
os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")

import aio_pika  # noqa: E402  (the stand-in)
from llmq.core.broker import BrokerManager  # noqa: E402
from llmq.core.models import Job, Result  # noqa: E402

from llmq_b200 import service as S  # noqa: E402
from llmq_b200.fixtures import build_tokenizer, make_jobs  # noqa: E402
from llmq_b200.worker import B200Worker  # noqa: E402
from tests.fake_engine import FakeEngine  # noqa: E402

VOCAB = 1024


@pytest.fixture()
def patched(monkeypatch):
    aio_pika.reset_brokers()
    tok = build_tokenizer(VOCAB)
    made = {}

    def fake_build_service(model_name, **kw):
        eng = FakeEngine(vocab=VOCAB, max_num_seqs=kw.get("max_num_seqs") or 8, max_model_len=64, eos_token_id=1)
        made["engine"] = eng
        made["kw"] = kw
        return S.GenerationService(eng, tok, 1)

    import llmq_b200.worker as W
    monkeypatch.setattr(W, "build_service", fake_build_service)
    monkeypatch.setenv("VLLM_MAX_TOKENS", "6")
    monkeypatch.setenv("VLLM_MAX_NUM_SEQS", "4")
    monkeypatch.setenv("VLLM_QUEUE_PREFETCH", "16")
    return made, tok


def test_worker_id_and_ctor_signature_match_reference(patched):
    w = B200Worker("random:llama-3-8b", "q", None, 1, 1, None, None, None, None)
    assert w.worker_id.startswith("b200-") and w.queue_name == "q" and not w.is_pipeline_worker
    wp = B200Worker("m", "pipeline.p.s1", pipeline_name="p", stage_name="s1", pipeline_stages=["s1", "s2"])
    assert wp.is_pipeline_worker


def test_rejects_tensor_parallel(patched):
    w = B200Worker("m", "q", tensor_parallel_size=2)
    with pytest.raises(ValueError):
        asyncio.run(w._initialize_processor())


def test_end_to_end_through_reference_base_worker_and_broker(patched):
    made, tok = patched

    async def main():
        w = B200Worker("random:llama-3-8b", "bq", tensor_parallel_size=1)
        task = asyncio.create_task(w.run())
        b = BrokerManager()
        await b.connect()
        await b.setup_queue_infrastructure("bq")
        jobs = [Job(id=f"j{i}", prompt="{a} w7", a=f"w{10 + i}", url=f"u{i}") for i in range(20)]
        jobs.append(Job(id="chat", messages=[{"role": "user", "content": "w5 w6"}]))
        jobs.append(Job(id="stop", prompt="w100", stop=["w103"]))
        jobs.append(Job(id="toolong", prompt=" ".join(["w9"] * 80)))  # > max_model_len: dropped
        jobs.append(Job(id="per-job-cap", prompt="w200", max_tokens=2))
        for j in jobs:
            await b.publish_job("bq", j)
        got = {}

        async def on_res(m):
            r = Result.parse_raw(m.body)
            got[r.id] = r
            await m.ack()

        await b.consume_results("bq", on_res)
        for _ in range(200):
            if len(got) >= len(jobs) - 1:
                break
            await asyncio.sleep(0.05)
        w.running = False
        await asyncio.wait_for(task, 10)
        return got, w

    got, w = asyncio.run(main())
    eng = made["engine"]
    assert made["kw"]["max_num_seqs"] == 4 and eng.max_batch_seen == 4  # VLLM_MAX_NUM_SEQS honoured
    assert "toolong" not in got  # ValueError => acked and dropped (base.py:228-235)
    assert len(got) == 23
    # fake model counts up from the last prompt token; 6 tokens (VLLM_MAX_TOKENS), text only
    # (the text is the CONTINUATION of the prompt's text, as vLLM's prompt-primed DecodeStream returns it
    # to the reference worker: this word-level vocabulary joins tokens with a space, so it starts with one
    # — pinned against the reference worker's own output in tests/test_vllm_worker_golden.py)
    assert got["j3"].result == " w8 w9 w10 w11 w12 w13"
    assert got["j3"].prompt == "w13 w7" and got["j3"].worker_id == w.worker_id
    assert got["j3"].model_dump()["url"] == "u3" and got["j3"].model_dump()["a"] == "w13"  # extras copied
    assert got["chat"].prompt == "Chat with 1 messages"
    assert got["stop"].result == " w101 w102 "  # cut right before the stop string (no stripping, as vLLM); generation aborted
    assert got["per-job-cap"].result == " w201 w202"
    # default sampling = the reference's literal temperature 0.7, unseeded (a stream per request)
    assert eng.last_sampling[0] == pytest.approx(0.7) and eng.last_sampling[1] != 0
    assert w.jobs_processed == 23


def test_chat_prompt_carries_double_bos_like_the_reference(patched):
    _, tok = patched
    w = B200Worker("m", "q", tensor_parallel_size=1)
    w.service = S.GenerationService(FakeEngine(), tok, 1)
    text = w.build_prompt(Job(id="c", messages=[{"role": "user", "content": "w5"}]))
    ids = tok(text, add_special_tokens=True).input_ids
    assert ids[0] == ids[1] == tok.bos_token_id
    assert w.stop_strings(Job(id="a", prompt="x")) is None
    assert w.stop_strings(Job(id="a", prompt="x", stop=[tok.eos_token])) is None  # inert special
    assert w.stop_strings(Job(id="a", prompt="x", stop=["END", tok.eos_token])) == ["END"]


def test_eos_finishes_and_is_not_rendered(patched):
    _, tok = patched
    eng = FakeEngine(vocab=VOCAB, eos_token_id=1)
    svc = S.GenerationService(eng, tok, 1)
    svc.start()

    async def main():
        loop = asyncio.get_running_loop()
        # last prompt token 1020 -> generates 1021, 1022/1023 (header specials), 0 (BOS), 1 (EOS): stop
        text, n = await svc.submit([0, 1020], 50, None, loop)
        return text, n

    text, n = asyncio.run(main())
    svc.stop()
    assert n == 5 and text == " w1021"


def test_job_stream_is_sharding_invariant():
    full = make_jobs(12, VOCAB, prompt_tokens=5)
    shards = [make_jobs(12, VOCAB, prompt_tokens=5, start=r, stride=3) for r in range(3)]
    merged = sorted((j for s in shards for j in s), key=lambda j: j["id"])
    assert merged == full and len({j["id"] for j in full}) == 12
    assert json.loads(json.dumps(full[0]))["prompt"].count("w") == 5


def test_config1_dummy_worker_1k_jobs_through_broker(monkeypatch):
    """BASELINE config #1: the reference's own DummyWorker (unmodified), 1k synthetic JSONL jobs
    through the broker — plumbing only.  Its 1 s sleep per job is patched to 1 ms."""
    import llmq.workers.dummy_worker as DW
    aio_pika.reset_brokers()
    real_sleep = asyncio.sleep

    async def fast_sleep(t):
        await real_sleep(0.001 if t == 1.0 else t)

    monkeypatch.setattr(DW.asyncio, "sleep", fast_sleep)

    async def main():
        w = DW.DummyWorker("dq", concurrency=250)
        task = asyncio.create_task(w.run())
        b = BrokerManager()
        await b.connect()
        await b.setup_queue_infrastructure("dq")
        for i in range(1000):
            await b.publish_job("dq", Job(id=f"job-{i:07d}", prompt="{text}", text=f"t{i}"))
        got = {}

        async def on_res(m):
            r = Result.parse_raw(m.body)
            got[r.id] = r.result
            await m.ack()

        await b.consume_results("dq", on_res)
        for _ in range(400):
            if len(got) >= 1000:
                break
            await real_sleep(0.05)
        w.running = False
        await asyncio.wait_for(task, 10)
        return got

    got = asyncio.run(main())
    assert len(got) == 1000 and got["job-0000042"] == "echo t42"


def test_pipeline_stage_routing_with_native_worker(patched):
    """config #4 plumbing: two B200Worker stages of a pipeline; stage 1's result becomes stage 2's
    prompt through the reference's publish_pipeline_result (ref:llmq/core/broker.py:145-193)"""
    made, tok = patched

    async def main():
        stages = ["translate", "format"]
        w1 = B200Worker("m", "pipeline.p.translate", pipeline_name="p", stage_name="translate", pipeline_stages=stages)
        w2 = B200Worker("m", "pipeline.p.format", pipeline_name="p", stage_name="format", pipeline_stages=stages)
        t1, t2 = asyncio.create_task(w1.run()), asyncio.create_task(w2.run())
        b = BrokerManager()
        await b.connect()
        await b.setup_pipeline_infrastructure("p", stages)
        for i in range(5):
            await b.publish_job("pipeline.p.translate", Job(id=f"p{i}", prompt=f"w{100 + i}", src=f"s{i}", temperature=0))
        got = {}

        async def on_res(m):
            r = Result.parse_raw(m.body)
            got[r.id] = r
            await m.ack()

        await b.consume_results("pipeline.p.results", on_res)
        for _ in range(200):
            if len(got) >= 5:
                break
            await asyncio.sleep(0.05)
        w1.running = w2.running = False
        await asyncio.wait_for(asyncio.gather(t1, t2), 10)
        return got

    got = asyncio.run(main())
    assert len(got) == 5
    # stage 1 counts up from w100: "w101 .. w106"; stage 2 is prompted with that text and counts up
    # from its last token w106: "w107 .. w112"; extras ride along
    assert got["p0"].result == " w107 w108 w109 w110 w111 w112"
    assert got["p0"].prompt == " w101 w102 w103 w104 w105 w106" and got["p0"].model_dump()["src"] == "s0"


def test_cli_install_routes_the_vllm_slot_to_the_native_worker():
    """`llmq worker run` lazily imports llmq.workers.vllm_worker.VLLMWorker (ref:llmq/cli/worker.py:20);
    llmq_b200.cli.install() publishes that module with the native worker and adds `worker b200`"""
    import importlib

    from llmq_b200 import cli
    cli.install()
    mod = importlib.import_module("llmq.workers.vllm_worker")
    assert mod.VLLMWorker is B200Worker
    from llmq.cli import main as M
    assert "b200" in M.worker.commands and "run" in M.worker.commands


def test_service_over_the_real_cpp_scheduler_dryrun(monkeypatch):
    """B200Worker + GenerationService on the REAL C++ engine in dry-run mode (scheduler, paged-KV
    block manager, chunked prefill, preemption — no GPU; token = previous + 1): ragged prompts,
    per-job max_tokens, a stop string and a KV pool too small for all of them at once must still
    give every job exactly its own continuation, through the reference's BaseWorker and broker."""
    from llmq_b200.fixtures import DryRunEngine

    aio_pika.reset_brokers()
    tok = build_tokenizer(VOCAB)
    made = {}

    def dry_build_service(model_name, **kw):
        eng = DryRunEngine(VOCAB, max_num_seqs=12, max_num_batched_tokens=48, max_model_len=96,
                           num_blocks=30, eos_token_id=None, policy=1)
        made["engine"] = eng
        return S.GenerationService(eng, tok, None)

    import llmq_b200.worker as W
    monkeypatch.setattr(W, "build_service", dry_build_service)
    monkeypatch.setenv("VLLM_MAX_TOKENS", "20")
    monkeypatch.setenv("VLLM_QUEUE_PREFETCH", "64")
    monkeypatch.setenv("B200Q_TEMPERATURE", "0")
    import numpy as np
    rng = np.random.default_rng(3)
    specs = []  # (id, last word, n_prompt_words, max_tokens or None)
    for i in range(60):
        n = int(rng.integers(1, 60))
        specs.append((f"d{i}", int(rng.integers(10, 900)), n, None if i % 3 else int(rng.integers(1, 20))))

    async def main():
        w = B200Worker("random:dry", "dq2", tensor_parallel_size=1)
        task = asyncio.create_task(w.run())
        b = BrokerManager()
        await b.connect()
        await b.setup_queue_infrastructure("dq2")
        for jid, last, n, cap in specs:
            words = ["w5"] * (n - 1) + [f"w{last}"]
            extra = {} if cap is None else {"max_tokens": cap}
            await b.publish_job("dq2", Job(id=jid, prompt=" ".join(words), **extra))
        await b.publish_job("dq2", Job(id="stop", prompt="w100", stop=["w104"]))
        got = {}

        async def on_res(m):
            r = Result.parse_raw(m.body)
            got[r.id] = r.result
            await m.ack()

        await b.consume_results("dq2", on_res)
        for _ in range(400):
            if len(got) == len(specs) + 1:
                break
            await asyncio.sleep(0.05)
        st = made["engine"].stats()  # before cleanup destroys the engine
        made["preemptions"], made["blocks"] = st.preemptions, (st.free_blocks, st.total_blocks)
        w.running = False
        await asyncio.wait_for(task, 10)
        return got

    got = asyncio.run(main())
    assert len(got) == len(specs) + 1
    for jid, last, n, cap in specs:
        k = 20 if cap is None else cap
        assert got[jid] == "".join(f" w{(last + 1 + j) % VOCAB}" for j in range(k)), jid
    assert got["stop"] == " w101 w102 w103 "
    # (how many preemptions the 30-block pool causes depends on how the jobs trickle in from the
    # broker; tests/test_scheduler_dryrun.py covers preemption deterministically)
    assert made["preemptions"] >= 0 and made["blocks"][0] == made["blocks"][1]


def test_token_table_matches_a_dict_of_lists_model():
    """the vectorised per-step token store of GenerationService against the obvious model, under
    random interleavings of allocate / step-append / release, growing rows and widths"""
    import numpy as np
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 2), st.integers(1, 40), st.integers(0, 10 ** 6)), min_size=1, max_size=120))
    def run(ops):
        t = S._TokenTable()
        model = {}   # slot -> [tokens]
        caps = {}
        rng = np.random.default_rng(0)
        for kind, a, b in ops:
            if kind == 0:  # allocate a request with max_new = a
                slot = t.alloc(a, has_stop=bool(b & 1))
                assert slot not in model and bool(t.has_stop[slot]) == bool(b & 1)
                model[slot], caps[slot] = [], a
            elif kind == 1 and model:  # one engine step over a random subset of live, non-full requests
                live = [s for s in model if len(model[s]) < caps[s]]
                if not live:
                    continue
                pick = rng.permutation(live)[: max(1, a % (len(live) + 1))]
                toks = rng.integers(0, 50000, size=len(pick)).astype(np.int32)
                t.append_step(np.asarray(pick, dtype=np.int64), toks)
                for s, x in zip(pick.tolist(), toks.tolist()):
                    model[s].append(x)
            elif kind == 2 and model:  # finish one
                s = sorted(model)[b % len(model)]
                assert t.tokens(s) == model[s]
                t.release(s)
                del model[s], caps[s]
            for s in model:
                assert int(t.n[s]) == len(model[s])
        for s in model:
            assert t.tokens(s) == model[s]

    run()


def test_engine_failure_rejects_and_requeues_instead_of_dropping(patched, monkeypatch):
    """SURVEY §8b error convention: only ValueError drops a job (ack).  A device / engine failure
    must surface as a non-ValueError from _process_job so that the reference's BaseWorker rejects
    the message with requeue=True (base.py:237-245) and another worker can pick it up; once the
    engine thread has died every later submit fails the same way (no silent hang)."""
    made, tok = patched

    class Boom(RuntimeError):
        pass

    async def main():
        w = B200Worker("random:llama-3-8b", "fq", tensor_parallel_size=1)
        await w._initialize_processor()
        eng = made["engine"]
        orig = eng.step
        calls = {"n": 0}

        def step():
            calls["n"] += 1
            if calls["n"] == 3:
                raise Boom("device fell off the bus")
            return orig()

        eng.step = step
        results = await asyncio.gather(*[w._process_job(Job(id=f"f{i}", prompt=f"w{50 + i}")) for i in range(6)],
                                       return_exceptions=True)
        assert all(isinstance(r, RuntimeError) and not isinstance(r, ValueError) for r in results), results
        assert "device fell off the bus" in str(results[0])
        with pytest.raises(RuntimeError, match="engine thread died"):
            await w._process_job(Job(id="late", prompt="w9"))
        # and through the reference's message envelope: rejected with requeue, never acked
        seen = {}

        class Msg:
            body = json.dumps({"id": "m1", "prompt": "w9"}).encode()

            async def ack(self):
                seen["ack"] = True

            async def reject(self, requeue=False):
                seen["reject"] = requeue

        await w._process_message(Msg())
        assert seen == {"reject": True}
        await w._cleanup_processor()

    asyncio.run(main())


def test_reference_quirks_are_replicated_not_fixed(patched):
    """behaviours of the reference worker that a drop-in must keep.
    D4 — a prompt is always passed through str.format (ref:llmq/core/models.py:46), so braces in a
    later pipeline stage's input (= the previous model's output) raise KeyError / IndexError
    (=> reject + requeue by the base class) or ValueError (=> dropped); D6 — chat jobs report the
    literal "Chat with N messages" as their prompt."""
    made, tok = patched

    async def main():
        w = B200Worker("random:llama-3-8b", "qq", tensor_parallel_size=1)
        await w._initialize_processor()
        with pytest.raises(KeyError):
            await w._process_job(Job(id="b1", prompt="w5 {not_a_field} w6"))
        with pytest.raises(IndexError):
            await w._process_job(Job(id="b2", prompt="w5 {} w6"))
        with pytest.raises(ValueError):
            await w._process_job(Job(id="b3", prompt="w5 { w6"))
        assert await w._process_job(Job(id="b4", prompt="w5 {{w6}} {x}", x="w7", temperature=0)) != ""  # escaped braces are fine
        await w._cleanup_processor()

    asyncio.run(main())


def test_absurd_max_tokens_fails_one_job_not_the_engine_thread(monkeypatch):
    """a job carrying max_tokens=3e9 used to raise MemoryError while sizing its token-table
    row, inside the engine thread, which killed generation for every request.  It is clamped to
    the engine's max_model_len now, and any per-request admission failure surfaces as ValueError
    for that request only (=> ack/drop by the base class), the others complete."""
    from llmq_b200.fixtures import DryRunEngine

    tok = build_tokenizer(VOCAB)
    eng = DryRunEngine(VOCAB, max_num_seqs=8, max_num_batched_tokens=64, max_model_len=64, num_blocks=64)
    svc = S.GenerationService(eng, tok, None)
    svc.start()

    async def main():
        loop = asyncio.get_running_loop()
        futs = [svc.submit([5, 6, 7], 3_000_000_000, None, loop),      # clamped: runs to max_model_len
                svc.submit([9], 4, None, loop),
                svc.submit([9], 0, None, loop),                        # max_tokens < 1: this job only
                svc.submit([9], 10 ** 30, None, loop)]
        return await asyncio.gather(*futs, return_exceptions=True)

    res = asyncio.run(main())
    svc.stop()
    assert svc.error is None, svc.error
    assert res[0][1] == 64 - 3 and res[1][1] == 4            # (text, n_tokens)
    assert isinstance(res[2], ValueError)
    assert res[3][1] == 63
    assert svc._table.tok.shape[1] <= 64                     # rows never sized beyond max_model_len


def test_fatal_engine_error_resolves_queued_and_in_hand_requests(monkeypatch):
    """when the engine thread dies, requests still in the inbox and the one being
    admitted must get their futures resolved too (their AMQP messages would otherwise stay un-acked
    for ever), and the worker is told to stop consuming."""
    tok = build_tokenizer(VOCAB)

    class DyingEngine(FakeEngine):
        def add_request(self, rid, ids, max_new, **kw):
            if rid >= 1:
                raise OSError("cuda context lost")   # not a per-request error
            return super().add_request(rid, ids, max_new, **kw)

    eng = DyingEngine(vocab=VOCAB, max_num_seqs=4, max_model_len=64)
    svc = S.GenerationService(eng, tok, None)
    fatal = []
    svc.on_fatal = fatal.append

    async def main():
        loop = asyncio.get_running_loop()
        futs = [svc.submit([5 + i], 4, None, loop) for i in range(6)]   # all queued before the thread starts
        svc.start()
        return await asyncio.wait_for(asyncio.gather(*futs, return_exceptions=True), 10)

    res = asyncio.run(main())
    assert all(isinstance(r, RuntimeError) and "engine failure" in str(r) for r in res), res
    assert len(fatal) == 1 and isinstance(svc.error, OSError)
    with pytest.raises(RuntimeError, match="engine thread died"):
        svc.submit([1], 1, None, None)


def test_worker_stops_consuming_when_the_engine_dies(patched):
    made, tok = patched

    async def main():
        w = B200Worker("random:llama-3-8b", "dq", tensor_parallel_size=1)
        await w._initialize_processor()
        w.running = True
        made["engine"].step = lambda: (_ for _ in ()).throw(OSError("xid 79"))
        with pytest.raises(RuntimeError):
            await w._process_job(Job(id="a", prompt="w9"))
        assert w.running is False
        await w._cleanup_processor()

    asyncio.run(main())


def test_stop_scanner_is_incremental_and_matches_whole_text_search():
    """stop strings were found by re-decoding the whole output after every
    token (O(n^2)).  _StopScanner decodes a bounded tail per token; on an 8192-token output (the
    reference's default VLLM_MAX_TOKENS) it must (a) find the same cut as a search of the full
    text, (b) call decode on short windows only."""
    import numpy as np

    tok = build_tokenizer(VOCAB)
    backend = tok.backend_tokenizer
    calls = []

    def decode(ids):
        calls.append(len(ids))
        return backend.decode(ids, skip_special_tokens=True)

    rng = np.random.default_rng(0)
    ids = rng.integers(10, 900, size=8192).tolist()
    # a stop string that spans two tokens, completed by the very last token only
    stop = [f"w{ids[-2]} w{ids[-1]}", "never-there"]
    ids_before = ids[:-2]
    while stop[0] in backend.decode(ids_before + [ids[-2]], skip_special_tokens=True):
        ids[-2] = (ids[-2] + 1) % 900 + 10  # keep the only match at the end
        stop[0] = f"w{ids[-2]} w{ids[-1]}"
    sc = S._StopScanner(stop, decode)
    cut = None
    for k, t in enumerate(ids):
        cut = sc.push(t)
        if cut is not None:
            break
    full = backend.decode(ids, skip_special_tokens=True)
    assert k == len(ids) - 1, "stop must complete on the last token"
    assert cut == full[: full.find(stop[0])]
    assert max(calls) <= 2 * S._StopScanner.PREFIX + 2, f"decode window grew to {max(calls)} ids"
    # the production path: tokenizers' DecodeStream primed with the prompt ids (what vLLM runs for the
    # reference worker) — the text is the continuation of the prompt, the cut the same search
    prompt = [5, 77]
    sc3 = S._StopScanner(stop, decode, backend, prompt)
    cut3 = None
    for k3, t in enumerate(ids):
        cut3 = sc3.push(t)
        if cut3 is not None:
            break
    cont = backend.decode(prompt + ids, skip_special_tokens=True)[len(backend.decode(prompt, skip_special_tokens=True)):]
    assert k3 == len(ids) - 1 and cut3 == cont[: cont.find(stop[0])] and cut3.startswith(" w")
    # special tokens never reach the text (skip_special_tokens) and never match
    sc2 = S._StopScanner(["<|end_of_text|>"], decode)
    from llmq_b200.fixtures import special_token_ids
    assert sc2.push(special_token_ids(VOCAB)["<|end_of_text|>"]) is None and sc2.text == ""


def test_detokenize_equals_vllms_prompt_primed_decode_stream():
    """GenerationService.detokenize (one decode of prompt tail + ids, minus the tail) against the
    reference behaviour it stands for: tokenizers.decoders.DecodeStream primed with the prompt ids and
    stepped token by token (vllm/v1/engine/detokenizer.py:181-184,215-222), special tokens skipped"""
    import numpy as np
    from tokenizers.decoders import DecodeStream

    from llmq_b200.fixtures import special_token_ids

    tok = build_tokenizer(VOCAB)
    backend = tok.backend_tokenizer
    svc = S.GenerationService(engine=None, tokenizer=tok, eos_token_id=None)
    sp = special_token_ids(VOCAB)
    bos, eos = sp["<|begin_of_text|>"], sp["<|end_of_text|>"]
    rng = np.random.default_rng(1)
    for trial in range(200):
        n_p, n_g = int(rng.integers(0, 20)), int(rng.integers(0, 30))
        prompt = [bos] + rng.integers(3, 900, size=n_p).tolist()
        if trial % 5 == 0:
            prompt += [eos, bos]          # chat-template-like prompts end in special tokens
        gen = rng.integers(3, 900, size=n_g).tolist()
        if trial % 7 == 0 and gen:
            gen[len(gen) // 2] = eos     # a special token inside the output is skipped
        stream = DecodeStream(ids=list(prompt), skip_special_tokens=True)
        want = "".join(filter(None, (stream.step(backend, t) for t in gen)))
        assert svc.detokenize(prompt[-svc.CONTEXT:], gen) == want, (prompt, gen)


def test_every_eos_id_of_the_model_ends_generation(tmp_path):
    """instruct models list several EOS ids (Llama-3.x-Instruct: 128001/128008/128009;
    gemma-2-it adds <end_of_turn> in generation_config.json); vLLM stops on all of them, so must
    the native engine — otherwise chat jobs run on to VLLM_MAX_TOKENS."""
    from llmq_b200.fixtures import DryRunEngine
    from llmq_b200.model import ModelSpec

    cfg = {"architectures": ["LlamaForCausalLM"], "hidden_size": 64, "num_hidden_layers": 1,
           "num_attention_heads": 2, "intermediate_size": 128, "vocab_size": VOCAB,
           "eos_token_id": [901, 908, 909], "bos_token_id": 900}
    spec = ModelSpec.from_hf_config(cfg)
    assert spec.eos_token_id == 901 and spec.eos_token_ids == (901, 908, 909)
    (tmp_path / "generation_config.json").write_text(json.dumps({"eos_token_id": [901, 107]}))

    class Tok:
        eos_token_id = 902

    assert S.collect_stop_ids(str(tmp_path), spec, Tok()) == [902, 901, 908, 909, 107]
    # the dry-run model counts up: prompt ending in 105 => 106, 107 (stop id) ...
    eng = DryRunEngine(VOCAB, max_num_seqs=4, max_num_batched_tokens=32, max_model_len=64, num_blocks=16,
                       eos_token_id=[902, 901, 908, 909, 107])
    eng.add_request(0, [5, 105], 30)
    eng.add_request(1, [5, 905], 30)
    eng.add_request(2, [5, 905], 30, ignore_eos=True)
    out = {0: [], 1: [], 2: []}
    while eng.has_work():
        ids, toks, flags = eng.step()
        for i, t in zip(ids.tolist(), toks.tolist()):
            out[i].append(t)
    assert out[0] == [106, 107] and out[1] == [906, 907, 908] and len(out[2]) == 30
    eng.close()


def test_default_max_model_len_is_the_models_own(monkeypatch):
    """with VLLM_MAX_MODEL_LEN unset the reference leaves the context length to vLLM
    (= max_position_embeddings), not 4096; prompts of 4096+ tokens on an 8k model must be served"""
    import llmq_b200.model as M
    seen = {}

    class StopHere(Exception):
        pass

    def fake_native_model(spec, weights, **kw):
        seen.update(kw)
        raise StopHere

    monkeypatch.setattr(M, "NativeModel", fake_native_model)
    monkeypatch.setattr(S.L, "require_device", lambda: None)
    import torch
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.delenv("B200Q_MAX_MODEL_LEN", raising=False)
    with pytest.raises(StopHere):
        S.build_service("random:llama-3-8b", max_num_seqs=4, max_model_len=None, gpu_memory_utilization=0.5)
    assert seen["max_model_len"] == 8192
    with pytest.raises(StopHere):
        S.build_service("random:llama-3-8b", max_num_seqs=4, max_model_len=1024, gpu_memory_utilization=0.5)
    assert seen["max_model_len"] == 1024


def test_prompts_of_one_loop_iteration_are_tokenised_in_one_batch_call():
    """host fast path: `encode_async` gathers every prompt submitted in the same event-loop iteration
    into ONE `encode_batch` (Rust thread pool, GIL released) — ids identical to the per-call `encode`,
    one bad text does not strand the others, and `detokenize_batch` equals the per-request `detokenize`"""
    import numpy as np

    tok = build_tokenizer(VOCAB)
    svc = S.GenerationService(engine=None, tokenizer=tok, eos_token_id=None)
    calls = []

    class CountingBackend:
        def __init__(self, be):
            self._be = be

        def encode_batch(self, texts, add_special_tokens=True):
            calls.append(len(texts))
            if any(t is None for t in texts):
                raise TypeError("not a string")
            return self._be.encode_batch(texts, add_special_tokens=add_special_tokens)

        def __getattr__(self, name):
            return getattr(self._be, name)

    svc.backend = CountingBackend(svc.backend)
    rng = np.random.default_rng(0)
    texts = [" ".join(f"w{int(i)}" for i in rng.integers(3, 900, size=int(rng.integers(1, 60)))) for _ in range(80)]

    async def main():
        a = await asyncio.gather(*[svc.encode_async(t) for t in texts])   # one iteration -> one batch
        b = []
        for t in texts[:5]:                                                 # a trickle -> batches of one
            b.append(await svc.encode_async(t))
        mixed = await asyncio.gather(svc.encode_async(texts[0]), svc.encode_async(None), svc.encode_async(texts[1]),
                                     return_exceptions=True)
        return a, b, mixed

    a, b, mixed = asyncio.run(main())
    assert a == [svc.encode(t) for t in texts] and b == a[:5]
    assert calls[0] == 80 and calls[1:6] == [1] * 5
    assert mixed[0] == a[0] and mixed[2] == a[1] and isinstance(mixed[1], Exception)
    pairs = [(a[i][-8:], rng.integers(3, 900, size=int(rng.integers(0, 40))).tolist()) for i in range(30)]
    assert svc.detokenize_batch(pairs) == [svc.detokenize(t, g) for t, g in pairs]
