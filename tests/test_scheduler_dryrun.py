"""Host logic of the C++ continuous-batching scheduler / paged-KV block manager, exercised through
the C ABI WITHOUT a GPU: `b200q_engine_create_dryrun` builds every step's metadata exactly as in
production, checks its invariants inside the library (token budget, unique KV slots, prefill tile
cover, positions) and fabricates each sampled token as (previous token + 1) mod vocab — so any
scheduling decision (chunking, admission order, preemption + recompute) must leave the per-request
token sequence unchanged.  Mirrors the behaviours listed at vllm/v1/core/sched/scheduler.py:329-340."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from llmq_b200 import lib as L

VOCAB = 1000


class DryEngine:
    def __init__(self, max_num_seqs, budget, max_model_len, num_blocks, policy, eos=-1, block_size=16,
                 async_steps=False):
        self.lib = L.load()
        cfg = L.EngineConfig(max_num_seqs=max_num_seqs, max_num_batched_tokens=budget,
                             max_model_len=max_model_len, eos_token_id=eos, policy=policy)
        h = C.c_void_p()
        L.check(self.lib.b200q_engine_create_dryrun(C.byref(cfg), VOCAB, block_size, num_blocks, C.byref(h)))
        self.h, self.cap = h, max_num_seqs
        # async stepping (the production default) returns a step's events one call later; the tests
        # that reason about WHICH step produced a token pin the synchronous mode
        L.check(self.lib.b200q_engine_set_async(h, int(async_steps)))
        self.ids = np.zeros(self.cap, np.int64)
        self.tok = np.zeros(self.cap, np.int32)
        self.flg = np.zeros(self.cap, np.int32)

    def add(self, rid, prompt, max_new, ignore_eos=True):
        arr = np.ascontiguousarray(prompt, dtype=np.int32)
        return self.lib.b200q_engine_add_request(self.h, rid, arr.ctypes.data, arr.size, max_new, int(ignore_eos))

    def step(self):
        n = C.c_int32(0)
        L.check(self.lib.b200q_engine_step(self.h, self.ids.ctypes.data, self.tok.ctypes.data,
                                           self.flg.ctypes.data, self.cap, C.byref(n)))
        k = n.value
        return self.ids[:k].tolist(), self.tok[:k].tolist(), self.flg[:k].tolist()

    def has_work(self):
        return bool(self.lib.b200q_engine_has_work(self.h))

    def stats(self):
        s = L.EngineStats()
        L.check(self.lib.b200q_engine_get_stats(self.h, C.byref(s)))
        return s

    def close(self):
        self.lib.b200q_engine_destroy(self.h)


def drain(eng, budget, max_steps=100000, outs=None, finished=None):
    outs = {} if outs is None else outs
    finished = {} if finished is None else finished
    order = []
    steps = 0
    while eng.has_work():
        ids, toks, flags = eng.step()
        s = eng.stats()
        assert s.last_step_tokens <= budget
        for i, t, f in zip(ids, toks, flags):
            outs.setdefault(i, []).append(t)
            if f:
                assert i not in finished
                finished[i] = f
                order.append(i)
        steps += 1
        assert steps < max_steps, "scheduler does not terminate"
    return outs, finished, order


def expected(prompt, n):
    return [(prompt[-1] + 1 + k) % VOCAB for k in range(n)]


@pytest.mark.parametrize("async_steps", [False, True])
@pytest.mark.parametrize("policy", [0, 1])
def test_chunked_prefill_and_budget(policy, async_steps):
    eng = DryEngine(max_num_seqs=4, budget=24, max_model_len=256, num_blocks=64, policy=policy, async_steps=async_steps)
    prompts = [list(range(3, 3 + n)) for n in (70, 5, 129, 16, 31, 1)]
    for i, p in enumerate(prompts):
        assert eng.add(i, p, 9) == 0
    outs, fin, _ = drain(eng, 24)
    for i, p in enumerate(prompts):
        assert outs[i] == expected(p, 9) and fin[i] == L.FLAG_FINISHED_LENGTH
    s = eng.stats()
    assert s.free_blocks == s.total_blocks == 64 and s.running == 0 and s.waiting == 0
    assert s.tokens_prefilled == sum(len(p) for p in prompts) and s.preemptions == 0
    eng.close()


@pytest.mark.parametrize("async_steps", [False, True])
@pytest.mark.parametrize("policy", [0, 1])
def test_no_livelock_when_two_requests_cannot_coexist(policy, async_steps):
    """found by the property test: a 114-token request and a 27-token request in a 9-block pool
    (10 blocks needed together).  The long one is preempted near its end; re-admitting it before the
    short one finishes used to starve the short one forever under the prefill-first policy."""
    eng = DryEngine(max_num_seqs=2, budget=17, max_model_len=160, num_blocks=9, policy=policy, async_steps=async_steps)
    a, b = [0], [(7 + j) % VOCAB for j in range(90)]
    assert eng.add(0, a, 26) == 0 and eng.add(1, b, 24) == 0
    outs, fin, order = drain(eng, 17, max_steps=2000)
    assert outs[0] == expected(a, 26) and outs[1] == expected(b, 24) and order == [0, 1]
    assert eng.stats().free_blocks == 9
    if policy == 1:  # (under vLLM order the short request gets ahead and the two never collide)
        assert eng.stats().preemptions >= 1
    eng.close()


@pytest.mark.parametrize("async_steps", [False, True])
@pytest.mark.parametrize("policy", [0, 1])
def test_preemption_recompute_keeps_sequences(policy, async_steps):
    # 8 prompts of exactly one block that all need a second block after their first decode token
    eng = DryEngine(max_num_seqs=8, budget=64, max_model_len=64, num_blocks=10, policy=policy, async_steps=async_steps)
    prompts = [[10 + i] * 16 for i in range(8)]
    for i, p in enumerate(prompts):
        assert eng.add(i, p, 20) == 0
    outs, fin, _ = drain(eng, 64)
    s = eng.stats()
    # (8 requests for 8 slots: the queue drains, so even the prefill-first policy admits optimistically)
    assert s.preemptions > 0 and s.free_blocks == 10
    for i, p in enumerate(prompts):
        assert outs[i] == expected(p, 20)
    eng.close()


def test_growth_aware_admission_under_a_backlog():
    """prefill-first policy with more waiting requests than free slots: admissions reserve the blocks
    the running requests are expected to grow into (max_new until a request has finished, then the
    running average of finished lengths), so the pool is not over-committed and nothing is recomputed;
    the vLLM order admits on today's free blocks and pays with preemptions.  When the estimate is too
    low (history of short requests) preempt-by-recompute still gives every request its continuation."""
    prompts = [[10 + i] * 16 for i in range(24)]

    def run(policy, history):
        eng = DryEngine(max_num_seqs=6, budget=64, max_model_len=64, num_blocks=12, policy=policy)
        if history:  # four requests that stop after one token: the length estimate becomes 1
            for i in range(4):
                assert eng.add(100 + i, [5 + i] * 4, 1) == 0
            drain(eng, 64)
        for i, p in enumerate(prompts):
            assert eng.add(i, p, 20) == 0
        outs, fin, _ = drain(eng, 64)
        s = eng.stats()
        assert s.free_blocks == 12
        for i, p in enumerate(prompts):
            assert outs[i] == expected(p, 20)
        eng.close()
        return s.preemptions

    # 3 blocks per request at full length, 12 blocks: 4 fit; 6 slots would over-commit
    assert run(0, False) > 0
    assert run(1, False) == 0
    assert run(1, True) > 0  # misleading history => optimistic => preemption path still exercised


@pytest.mark.parametrize("async_steps", [False, True])
def test_eos_and_length_and_argument_checks(async_steps):
    eng = DryEngine(max_num_seqs=4, budget=64, max_model_len=32, num_blocks=16, policy=1, eos=7, async_steps=async_steps)
    assert eng.add(1, [3, 4, 5], 50, ignore_eos=False) == 0   # generates 6, 7(EOS) -> stops
    assert eng.add(2, [3, 4, 5], 4, ignore_eos=True) == 0     # runs through the EOS id
    assert eng.add(3, [9] * 10, 1000) == 0                     # clipped by max_model_len
    assert eng.add(3, [9], 5) == -1                            # duplicate id
    assert eng.add(4, [9] * 32, 5) == -1                       # prompt fills the window
    assert eng.add(5, [VOCAB], 5) == -1                        # token id out of range
    assert eng.add(6, [], 5) == -1
    outs, fin, _ = drain(eng, 64)
    assert outs[1] == [6, 7] and fin[1] == L.FLAG_FINISHED_EOS
    assert outs[2] == [6, 7, 8, 9] and fin[2] == L.FLAG_FINISHED_LENGTH
    assert len(outs[3]) == 22 and fin[3] == L.FLAG_FINISHED_LENGTH  # 10 + 22 == max_model_len
    assert eng.lib.b200q_engine_abort(eng.h, 99) == 0
    eng.close()


def test_policies_order_work_differently_but_equivalently():
    """vLLM order serves running decodes before admitting; prefill-first fills the batch first"""
    res = {}
    for policy in (0, 1):
        eng = DryEngine(max_num_seqs=8, budget=32, max_model_len=128, num_blocks=128, policy=policy)
        for i in range(8):
            eng.add(i, [5 + i] * 30, 6)
        first_tokens_step = {}
        steps = 0
        while eng.has_work():
            ids, _, _ = eng.step()
            steps += 1
            for i in ids:
                first_tokens_step.setdefault(i, steps)
        res[policy] = (steps, first_tokens_step)
        eng.close()
    # (sync mode: a call returns the events of the step it ran)
    # prefill-first: request 0 cannot decode until every prompt is in (8 prompt steps), so its
    # 6 tokens span more steps; vLLM order finishes request 0 while later prompts still prefill
    assert res[1][1][0] == res[0][1][0] == 1
    assert res[0][0] >= res[1][0]  # prefill-first never needs more steps in total


@settings(max_examples=120, deadline=None)
@given(
    lens=st.lists(st.tuples(st.integers(1, 90), st.integers(1, 40)), min_size=1, max_size=24),
    seqs=st.integers(1, 12), budget=st.integers(16, 96), policy=st.integers(0, 1),
    pool=st.integers(9, 60), abort_one=st.booleans(), async_steps=st.booleans(), eos=st.sampled_from([-1, 3, 500]),
)
def test_scheduler_properties(lens, seqs, budget, policy, pool, abort_one, async_steps, eos):
    """for arbitrary workloads and pool sizes: every request finishes with exactly its sequence, no
    block leaks, budget and slot invariants hold every step (checked inside the library)"""
    eng = DryEngine(max_num_seqs=seqs, budget=budget, max_model_len=160, num_blocks=pool, policy=policy,
                    async_steps=async_steps, eos=eos)
    prompts = {}
    for i, (pl, mn) in enumerate(lens):
        p = [(7 * i + j) % VOCAB for j in range(pl)]
        # every third request honours the stop id: in async mode a request that samples it has
        # already been given one more token in the next step, which must be discarded silently
        if eng.add(i, p, mn, ignore_eos=i % 3 != 0) == 0:  # requests that can never fit the pool are rejected up front
            prompts[i] = (p, mn)
    outs0, fin0 = {}, {}
    if abort_one and prompts:
        victim = sorted(prompts)[0]
        ids, toks, flags = eng.step()
        for i, t, f in zip(ids, toks, flags):
            outs0.setdefault(i, []).append(t)
            if f:
                fin0[i] = f
        assert eng.lib.b200q_engine_abort(eng.h, victim) == 0
        prompts.pop(victim)
    outs, fin, _ = drain(eng, budget, outs=outs0, finished=fin0)
    for i, (p, mn) in prompts.items():
        n = min(mn, 160 - len(p))
        want = expected(p, n)
        if i % 3 == 0 and eos in want:  # stops at (and including) the first stop id
            want = want[: want.index(eos) + 1]
            assert fin.get(i) == L.FLAG_FINISHED_EOS
        assert outs.get(i, []) == want, "tokens must not depend on chunking / preemption / async stepping"
        assert i in fin
    s = eng.stats()
    assert s.free_blocks == s.total_blocks == pool and s.running == 0 and s.waiting == 0
    eng.close()
