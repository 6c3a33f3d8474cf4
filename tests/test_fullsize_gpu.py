"""Size-independent properties at BASELINE.json's full model sizes (random-init Llama-3.2-1B and
Llama-3-8B, 128-token prompts): the oracle cannot run these in seconds, so parity is carried by
(a) the op/model tests on small shapes and (b) these invariants of the same code at full size:
determinism, batch invariance (a request's tokens do not depend on its batch-mates), chunked
prefill transparency (token budget smaller than the prompt), and a logits cross-check of the
LM-head + argmax path against a torch fp32-accumulate matmul on the device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def gen(svc, prompts, max_new, **ekw):
    eng = svc.engine
    for i, p in enumerate(prompts):
        eng.add_request(i, p, max_new, ignore_eos=True)
    outs = {i: [] for i in range(len(prompts))}
    while eng.has_work():
        ids, toks, _ = eng.step()
        for i, t in zip(ids.tolist(), toks.tolist()):
            outs[i].append(t)
    return [outs[i] for i in range(len(prompts))]


@pytest.mark.parametrize("key", ["llama-3.2-1b", "llama-3-8b"])
def test_fullsize_invariants(cuda, key):
    from llmq_b200.fixtures import make_jobs
    from llmq_b200.model import Engine
    from llmq_b200.service import build_service

    svc = build_service(f"random:{key}", max_num_seqs=64, max_model_len=256, gpu_memory_utilization=0.9,
                        max_num_batched_tokens=2048, seed=7, num_blocks=2048)
    model, tok = svc.engine.model, svc.tokenizer
    jobs = make_jobs(40, model.spec.vocab, prompt_tokens=127)
    prompts = [tok(j["prompt"], add_special_tokens=True).input_ids for j in jobs]
    assert all(len(p) == 128 for p in prompts)
    from llmq_b200 import lib
    L = lib.load()
    fast = gen(svc, prompts, 6)
    assert fast == gen(svc, prompts, 6), "default (split-K on) path must be deterministic run to run"
    # exact batch invariance holds when every GEMM reduces K in one sequential pass; split-K (used
    # for decode-sized batches) changes the fp32 summation order with the batch size, so the
    # invariance checks below run in the library's batch-invariant mode (split-K off)
    lib.check(L.b200q_gemm_set_splitk(1))
    a = gen(svc, prompts, 6)
    b = gen(svc, prompts, 6)
    assert a == b, "same inputs, same engine: outputs must be identical (determinism)"
    assert all(len(o) == 6 and all(0 <= t < model.spec.vocab for t in o) for o in a)
    # batch invariance: each of the first 4 requests alone
    for i in range(4):
        assert gen(svc, [prompts[i]], 6)[0] == a[i], f"request {i} changed with its batch-mates"
    # chunked prefill (budget 48 < 128-token prompts) through a second engine on the same model
    svc.engine.close()
    eng2 = Engine(model, max_num_seqs=8, max_num_batched_tokens=48, eos_token_id=None)
    svc.engine = eng2
    c = gen(svc, prompts[:6], 6)
    # the prefill kernel tiles the same keys identically for any chunking, so ids are equal except
    # where decode-vs-prefill attention rounding meets a near-tie; require >= 5 of 6 identical
    assert sum(x == y for x, y in zip(c, a[:6])) >= 5
    # LM head + argmax cross-check on the last step's logits
    n = 6
    logits = model.logits_view(n).float()
    assert torch.isfinite(logits).all()
    # the 6 requests finish on the same (last) step, rows in request order: the ids the engine
    # returned are the lowest-index argmax of exactly these logits (V = 128256 at full width)
    assert np.array_equal(logits.argmax(-1).cpu().numpy(), np.array([o[-1] for o in c])[:n])
    lib.check(L.b200q_gemm_set_splitk(0))
    eng2.close()
    model.close()


@pytest.mark.parametrize("key", ["llama-3-8b", "llama-3.2-1b"])
def test_fullwidth_two_layer_logits_and_ids_match_the_oracle(cuda, key):
    """the oracle comparison at BASELINE widths.  Two decoder layers with the
    REAL hidden / intermediate / head / vocab sizes (K = 4096 and 14336 accumulations, V = 128256
    argmax, llama3 rope scaling and tied head for the 1B) on a seeded checkpoint; the batch is the
    bench's shape — 36 prompts x 128 tokens = 4608 prefill rows in ONE step, which dispatches the
    cta_group::2 GEMM variants and the fused SwiGLU epilogue exactly as the headline run does — and
    then one decode step (36 rows: the small-M / split-K variants and the paged decode attention).
    Logits of both steps against the CPU oracle's, within the tolerance stated below."""
    import dataclasses

    from llmq_b200.fixtures import seeded_state_dict
    from llmq_b200.model import BUILTIN_SPECS, Engine, NativeModel, fuse_hf_weights
    from oracle import ops as O
    from oracle.model import LlamaDims, LlamaOracle

    spec = dataclasses.replace(BUILTIN_SPECS[key], n_layers=2, max_position_embeddings=256)
    sd = seeded_state_dict(spec, seed=99)
    model = NativeModel(spec, fuse_hf_weights(spec, sd), max_tokens=4608, max_seqs=64, max_model_len=256,
                        num_blocks=36 * 9 + 8)
    oracle = LlamaOracle(LlamaDims.from_hf_config(spec.to_hf_config()), sd, "bf16", max_pos=256)
    g = np.random.default_rng(5)
    prompts = [g.integers(0, 128000, size=128).tolist() for _ in range(36)]
    ref = [oracle.greedy(p, 2, return_logits=True) for p in prompts]

    def run(max_new):
        eng = Engine(model, max_num_seqs=64, max_num_batched_tokens=4608, eos_token_id=None)
        for i, p in enumerate(prompts):
            eng.add_request(i, p, max_new, ignore_eos=True)
        outs = {i: [] for i in range(len(prompts))}
        steps = 0
        while eng.has_work():
            ids, toks, _ = eng.step()
            for i, t in zip(ids.tolist(), toks.tolist()):
                outs[i].append(t)
        steps = eng.stats().steps
        eng.close()
        return outs, steps, model.logits_view(len(prompts)).float().cpu()

    for max_new in (1, 2):
        outs, steps, got = run(max_new)
        assert steps == max_new, "36 x 128 tokens must go through as ONE prefill step (+ one decode step)"
        want = torch.stack([ref[i][1][max_new - 1] for i in range(len(prompts))])
        # the decode step continues from OUR first token: rows whose first token differs from the
        # oracle's (a near-tie, checked in the first pass) are a different text and are left out
        same = torch.tensor([outs[i][:max_new - 1] == ref[i][0][:max_new - 1] for i in range(len(prompts))])
        assert int(same.sum()) >= len(prompts) - 4, f"{int((~same).sum())} of {len(prompts)} rows diverged on the first token"
        got, want = got[same], want[same]
        diff = (got - want).abs()
        # stated tolerance.  Logits here have std 1.3 and reach |l| ~ 6 over 36 x 128256 entries, where
        # one bf16 ulp is 0.031: max |diff| < 0.1 (3 ulps at the largest magnitudes), mean < 0.015
        # (1 % of the logit std: two layers of bf16 activations with different fp32 summation orders);
        # measured on B200: max 0.078, mean 0.0105
        assert diff.max().item() < 0.1 and diff.mean().item() < 0.015, (max_new, diff.max().item(), diff.mean().item())
        top2 = want.topk(2, -1).values
        margin = top2[:, 0] - top2[:, 1]
        ids_got = np.array([outs[i][max_new - 1] for i in range(len(prompts))])[same.numpy()]
        ids_ref = O.argmax_first(want)
        # every decision the oracle makes with a margin above the tolerance must be reproduced
        safe = (margin > 0.1).numpy()
        assert np.array_equal(ids_got[safe], ids_ref[safe]), (ids_got, ids_ref, margin)
        assert safe.sum() >= len(ids_got) // 2, "test too weak: most margins are below the tolerance"
        assert np.array_equal(ids_got, O.argmax_first(got)), "argmax kernel vs its own logits at V=128256"
    model.close()
