"""Parity against the reference's OWN worker at BASELINE widths.

`tests/golden/vllm_worker_golden_<case>.json.gz` were produced on a B200 by
`tools/oracle_vllm_worker.py golden`: the reference's `VLLMWorker` (ref:llmq/workers/vllm_worker.py:11,
unmodified, vLLM 0.22.0) behind the unmodified `BaseWorker` / `BrokerManager`, instrumented only to
decode greedily and to record token ids and the top-2 logprob margin of every decision.  Seeded
random-init checkpoints with the REAL widths of the BASELINE models:
  llama32_1b    Llama-3.2-1B, all 16 layers          llama3_8b_w4   Llama-3-8B widths, 4 layers
76 jobs each (64 of the canonical 128-token benchmark stream + 12 ragged lengths), 128 new tokens,
twice: as the reference configures vLLM (compiled, CUDA graphs) and with enforce_eager.

What "bit-exact greedy ids" can mean here is set by vLLM itself: with V = 128256 and random-init
weights 5-10 % of all decisions have a top-2 margin of at most one bf16 ulp of the logits, and the
two vLLM modes agree on NONE of the 76 sequences over 128 tokens (first divergences at margins up to
0.094).  So the contract (the north-star's "logits within a stated tolerance") is:
  * ids are identical up to the first divergence of a sequence;
  * a first divergence may only happen where vLLM's own top-2 margin is <= MARGIN_TOL — after it the
    texts differ and are not compared;
  * over all sequences we agree with vLLM on at least as long a prefix as vLLM's two modes do with
    each other (x AGREEMENT_FLOOR).
"""
import asyncio
import gzip
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(p, "llmq")):
        sys.path.insert(0, p)
        break
os.environ.setdefault("LLMQ_LOG_LEVEL", "WARNING")

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["llama32_1b", "llama3_8b_w4"]
# logits of these models reach |l| ~ 4-8, where one bf16 ulp is 0.03125: 4 ulps.  vLLM's own two
# modes diverge at margins up to 0.094 (3 ulps) on these fixtures.
MARGIN_TOL = 0.125
AGREEMENT_FLOOR = 0.6


def load(case):
    with gzip.open(os.path.join(G, f"vllm_worker_golden_{case}.json.gz"), "rt") as f:
        d = json.load(f)
    runs = {k: v for k, v in d["runs"].items() if "error" not in v}
    assert set(runs) == {"default", "eager"}
    return d, runs


def first_divergence(a, b):
    n = min(len(a), len(b))
    return next((i for i in range(n) if a[i] != b[i]), None if len(a) == len(b) else n)


def agreed_prefix_stats(ours, run):
    """mean length of the common prefix with a vLLM run, and the list of first divergences"""
    lens, divs = [], []
    for jid, c in run.items():
        k = first_divergence(ours[jid], c["ids"])
        lens.append(len(c["ids"]) if k is None else k)
        if k is not None:
            divs.append((jid, k))
    return float(np.mean(lens)), divs


def check_divergences(ours, run, what):
    mean_len, divs = agreed_prefix_stats(ours, run)
    for jid, k in divs:
        c = run[jid]
        assert k < len(c["ids"]) and k < len(ours[jid]), f"{what} {jid}: one sequence is a strict prefix of the other"
        margin = c["margins"][k]
        assert margin is not None and margin <= MARGIN_TOL, (
            f"{what} {jid}: diverged from vLLM at token {k} ({ours[jid][k]} vs {c['ids'][k]}) where vLLM's own "
            f"top-2 margin is {margin} > {MARGIN_TOL}")
        # (the other token is usually vLLM's runner-up, but with several logits within a few ulps of each
        # other it need not be: vLLM's own two modes pick a third token at such positions)
    return mean_len, len(divs)


@pytest.mark.parametrize("case", CASES)
def test_fixture_is_the_canonical_job_stream_and_our_tokenisation(case):
    """the prompt ids vLLM's renderer produced for the reference worker's text prompt
    (add_special_tokens=True: one BOS) are exactly what GenerationService.encode produces, and the
    text the reference worker returned is what GenerationService.detokenize makes of the same ids
    (vLLM's DecodeStream is primed with the prompt: the text is the CONTINUATION of the prompt, which
    for this word-level vocabulary starts with the joining space)"""
    from oracle_vllm_worker import golden_jobs

    from llmq_b200.fixtures import build_tokenizer
    from llmq_b200.service import GenerationService

    d, runs = load(case)
    vocab = d["spec"]["vocab_size"]
    assert d["jobs"] == golden_jobs(vocab)
    svc = GenerationService(engine=None, tokenizer=build_tokenizer(vocab), eos_token_id=None)
    for n_j, j in enumerate(d["jobs"]):
        c = runs["default"][j["id"]]
        assert svc.encode(j["prompt"]) == c["prompt_ids"], j["id"]
        text = svc.detokenize(c["prompt_ids"][-svc.CONTEXT:], c["ids"])
        assert hashlib.sha1(text.encode()).hexdigest() == c["text_sha1"], j["id"]
        if n_j < 8:
            assert text == c["text"]
        assert len(c["ids"]) == len(c["margins"]) == len(c["runner_up"]) == d["max_new_tokens"]


@pytest.mark.parametrize("case", CASES)
def test_vllm_disagrees_with_itself_only_at_near_ties(case):
    """the floor: the two vLLM modes against each other under the same contract"""
    d, runs = load(case)
    a = {jid: c["ids"] for jid, c in runs["default"].items()}
    mean_len, n_div = check_divergences(a, runs["eager"], "vLLM[default] vs vLLM[eager]")
    assert n_div >= 60, "expected (and documented): vLLM's modes diverge on nearly every 128-token sequence"
    assert mean_len < 64


def test_cpu_oracle_reproduces_vllm_decisions_at_full_1b_size():
    """oracle/model.py pinned against the reference worker's ids at the real Llama-3.2-1B size
    (16 layers, tied head, llama3 rope scaling): three jobs, greedy until the first divergence"""
    from llmq_b200.fixtures import seeded_state_dict
    from llmq_b200.model import ModelSpec
    from oracle.model import LlamaDims, LlamaOracle

    d, runs = load("llama32_1b")
    spec = ModelSpec.from_hf_config(d["spec"])
    oracle = LlamaOracle(LlamaDims.from_hf_config(d["spec"]), seeded_state_dict(spec, d["weights_seed"]), "bf16",
                         max_pos=512)
    ours = {}
    picked = ["job-0000000", "job-0000001", "ragged-0017"]
    for jid in picked:
        c = runs["default"][jid]
        ours[jid] = oracle.greedy(c["prompt_ids"], 24)
    for tag in ("default", "eager"):
        run = {jid: {**runs[tag][jid], "ids": runs[tag][jid]["ids"][:24]} for jid in picked}
        check_divergences(ours, run, f"oracle vs vLLM[{tag}]")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_native_worker_matches_the_reference_worker(cuda, case, tmp_path, monkeypatch):
    """the drop-in on the same seeded checkpoint directory, the same JSONL through the same broker
    path, greedy: B200Worker's ids against the reference worker's"""
    import aio_pika
    from llmq.core.broker import BrokerManager
    from llmq.core.models import Job, Result

    from llmq_b200.fixtures import write_model_dir
    from llmq_b200.model import ModelSpec
    from llmq_b200.worker import B200Worker

    d, runs = load(case)
    spec = ModelSpec.from_hf_config(d["spec"], name=case)
    mdir = write_model_dir(str(tmp_path / case), spec, seed=d["weights_seed"], with_weights=True)
    monkeypatch.setenv("VLLM_MAX_TOKENS", str(d["max_new_tokens"]))
    monkeypatch.setenv("VLLM_MAX_NUM_SEQS", "128")
    monkeypatch.setenv("VLLM_QUEUE_PREFETCH", "256")
    monkeypatch.setenv("VLLM_MAX_MODEL_LEN", "1024")
    monkeypatch.setenv("VLLM_GPU_MEMORY_UTILIZATION", "0.5")
    monkeypatch.setenv("B200Q_TEMPERATURE", "0")
    aio_pika.reset_brokers()
    # the worker hands text back; remember which ids each text was detokenised from
    from llmq_b200.service import GenerationService
    ids_of_text = {}
    detokenize_batch = GenerationService.detokenize_batch

    def recording_detokenize(self, pairs):
        pairs = list(pairs)
        texts = detokenize_batch(self, pairs)
        for (_, ids), text in zip(pairs, texts):
            ids_of_text[text] = list(ids)
        return texts

    monkeypatch.setattr(GenerationService, "detokenize_batch", recording_detokenize)

    async def main():
        w = B200Worker(mdir, "wg", tensor_parallel_size=1)
        task = asyncio.create_task(w.run())
        b = BrokerManager()
        await b.connect()
        await b.setup_queue_infrastructure("wg")
        for j in d["jobs"]:
            await b.publish_job("wg", Job(**j))
        got = {}

        async def on_res(m):
            r = Result.parse_raw(m.body)
            got[r.id] = r.result
            await m.ack()

        await b.consume_results("wg", on_res)
        for _ in range(6000):
            if len(got) >= len(d["jobs"]):
                break
            await asyncio.sleep(0.05)
        w.running = False
        await asyncio.wait_for(task, 60)
        return got

    got = asyncio.run(main())
    assert len(got) == len(d["jobs"])
    ours = {jid: ids_of_text[text] for jid, text in got.items()}
    floor, _ = agreed_prefix_stats({jid: c["ids"] for jid, c in runs["default"].items()}, runs["eager"])
    report = {}
    for tag in ("default", "eager"):
        mean_len, n_div = check_divergences(ours, runs[tag], f"b200 worker vs vLLM[{tag}]")
        exact_text = sum(hashlib.sha1(got[jid].encode()).hexdigest() == c["text_sha1"] for jid, c in runs[tag].items())
        report[tag] = (round(mean_len, 1), n_div, exact_text)
    print(f"\n[{case}] agreed prefix (mean tokens), diverging sequences, identical texts vs vLLM: {report}; "
          f"vLLM default-vs-eager agreed prefix: {floor:.1f}")
    best = max(v[0] for v in report.values())
    assert best >= AGREEMENT_FLOOR * floor, (report, floor)
