"""GenerationService — engine thread + request table around the native engine.

Pure host logic with no llmq import, so it can be driven directly (bench.py, tests) exactly the
way ``B200Worker._process_job`` drives it: ``future = service.submit(ids, ...); await future``.

Threading model: the asyncio event-loop thread owns AMQP, tokenisation and
futures; ONE engine thread owns the CUDA context and runs the continuous-batching step loop
(`b200q_engine_step`, GIL released inside the C call).
"""
from __future__ import annotations

import asyncio
import os
import queue
import threading
import time
from typing import Dict, List, Optional

from . import lib as L


class _Req:
    __slots__ = ("slot", "prompt_ids", "max_new", "stop", "future", "loop",
                 "temperature", "seed", "detok")

    def __init__(self, prompt_ids, max_new, stop, future, loop, temperature=0.0, seed=0):
        self.prompt_ids, self.max_new, self.stop = prompt_ids, max_new, stop
        self.future, self.loop = future, loop
        self.temperature, self.seed = temperature, seed
        self.slot = -1   # engine request id = row of the token table, assigned on the engine thread
        self.detok = None  # _StopScanner, only for requests that carry stop strings


class _TokenTable:
    """Generated tokens of the live requests as one int32 matrix (row = request slot = the engine's
    request id), so that a step's (ids, tokens) arrays are stored with two vectorised numpy
    operations instead of a Python loop over the batch: with thousands of sequences per step that
    loop sits between two GPU steps and idles the device (host fast path)."""

    def __init__(self):
        import numpy as np

        self.np = np
        self.tok = np.empty((0, 0), dtype=np.int32)
        self.n = np.zeros(0, dtype=np.int64)
        self.has_stop = np.zeros(0, dtype=bool)
        self.free: List[int] = []

    def alloc(self, max_new: int, has_stop: bool) -> int:
        np = self.np
        rows, width = self.tok.shape
        if not self.free or max_new > width:
            new_rows = rows if self.free else max(64, rows * 2)
            new_width = max(width, max_new)
            tok = np.empty((new_rows, new_width), dtype=np.int32)  # untouched pages cost nothing
            tok[:rows, :width] = self.tok
            self.tok = tok
            if new_rows > rows:
                self.n = np.concatenate([self.n, np.zeros(new_rows - rows, dtype=np.int64)])
                self.has_stop = np.concatenate([self.has_stop, np.zeros(new_rows - rows, dtype=bool)])
                self.free.extend(range(new_rows - 1, rows - 1, -1))
        slot = self.free.pop()
        self.n[slot] = 0
        self.has_stop[slot] = has_stop
        return slot

    def release(self, slot: int) -> None:
        self.has_stop[slot] = False
        self.free.append(slot)

    def append_step(self, slots, toks) -> None:
        """one token for each of `slots` (unique within a step)"""
        n = self.n[slots]
        self.tok[slots, n] = toks
        self.n[slots] = n + 1

    def tokens(self, slot: int) -> List[int]:
        return self.tok[slot, : self.n[slot]].tolist()


class _StopScanner:
    """Incremental detokenisation + stop-string scan of ONE request — vLLM's detokenizer
    (vllm/v1/engine/detokenizer.py:95-165,167-247): a `tokenizers.decoders.DecodeStream` primed with
    the prompt ids yields the characters each new token adds (nothing while a multi-byte character is
    incomplete); only the new characters plus a hold-back of ``max(len(stop)) - 1`` are searched
    (:86-87) and the output is cut before the earliest stop string.  O(1) work per token, where the
    first version re-decoded the whole output after every token.

    Without a Rust tokenizer the same is emulated by decoding a bounded tail of ids
    (``PREFIX`` ids of left context) and taking what the new token adds to it."""

    PREFIX = 6  # fallback only: ids of left context that settle spacing / byte-fallback merges

    def __init__(self, stop: List[str], decode, backend=None, prompt_ids=None):
        self.stop = stop
        self.hold = max(len(s) for s in stop) - 1
        self.text = ""          # emitted text so far
        self.n = 0              # generated ids consumed
        self.stream = None
        if backend is not None:
            from tokenizers.decoders import DecodeStream

            self.backend = backend
            self.stream = DecodeStream(ids=list(prompt_ids or []), skip_special_tokens=True)
        else:
            self.decode = decode
            self.ids: List[int] = list(prompt_ids or [])[-self.PREFIX:]
            self.prefix_off = 0     # ids[prefix_off:read_off] = context already reflected in `text`
            self.read_off = len(self.ids)

    def _step(self, token: int) -> str:
        if self.stream is not None:
            return self.stream.step(self.backend, token) or ""
        self.ids.append(token)
        ids = self.ids
        prefix_text = self.decode(ids[self.prefix_off:self.read_off]) if self.read_off > self.prefix_off else ""
        full = self.decode(ids[self.prefix_off:])
        if full.endswith("\ufffd") or len(full) <= len(prefix_text):
            return ""  # an incomplete multi-byte character (or nothing new): wait for more ids
        self.prefix_off = max(self.read_off, len(ids) - self.PREFIX)
        self.read_off = len(ids)
        return full[len(prefix_text):]

    def push(self, token: int) -> Optional[str]:
        """append one token; returns the final (cut) text if a stop string completed, else None"""
        self.n += 1
        new = self._step(token)
        if not new:
            return None
        start = max(0, len(self.text) - self.hold)
        self.text += new
        best = -1
        for s in self.stop:
            i = self.text.find(s, start)
            if i >= 0 and (best < 0 or i < best):
                best = i
        return self.text[:best] if best >= 0 else None


class GenerationService:
    """Engine thread + request table.  Pure host logic around llmq_b200.model.Engine; also used
    directly by bench.py (the same public call a worker makes)."""

    def __init__(self, engine, tokenizer, eos_token_id: Optional[int]):
        self.engine = engine
        self.tokenizer = tokenizer
        self.eos_token_id = eos_token_id
        # Text <-> ids through the Rust tokenizer itself when there is one: same ids and text as the
        # transformers wrapper minus ~70 us of Python per call, and no clean_up_tokenization_spaces
        # pass on decode — which is what vLLM's FastIncrementalDetokenizer (tokenizers' DecodeStream,
        # vllm/v1/engine/detokenizer.py:167-247) produces for the reference worker.
        # tokenizers' batch calls run on its Rust thread pool with the GIL released — the per-call ones hold
        # the GIL for the whole ~150 us of a 128-token prompt, which also stalls the engine thread's Python
        os.environ.setdefault("TOKENIZERS_PARALLELISM", "true")
        self._enc_pending: Dict[object, list] = {}   # event loop -> [(text, future)] awaiting one batch encode
        backend = getattr(tokenizer, "backend_tokenizer", None)
        if backend is not None and hasattr(backend, "encode") and hasattr(backend, "decode"):
            self.backend = backend
            self.encode = lambda text: backend.encode(text, add_special_tokens=True).ids
            self.decode = lambda ids: backend.decode(ids, skip_special_tokens=True)
        else:
            self.backend = None
            self.encode = lambda text: tokenizer(text, add_special_tokens=True).input_ids
            self.decode = lambda ids: tokenizer.decode(ids, skip_special_tokens=True)
        self._inbox: "queue.SimpleQueue[_Req]" = queue.SimpleQueue()
        self._reqs: Dict[int, _Req] = {}   # by slot; touched only on the engine thread
        self._table = _TokenTable()
        self._next_id = 0
        # B200Q_SEED pins the per-request streams (reproducible runs); default: fresh entropy
        env_seed = os.environ.get("B200Q_SEED")
        self._base_seed = int(env_seed) if env_seed else int.from_bytes(os.urandom(8), "little")
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, name="b200q-engine", daemon=True)
        self.error: Optional[BaseException] = None
        self.on_fatal = None   # callable(exc): the worker stops consuming when the engine thread dies
        self.tokens_out = 0
        self.jobs_done = 0
        self.first_submit_t: Optional[float] = None
        self.last_finish_t: Optional[float] = None

    CONTEXT = 8  # prompt ids of left context for the continuation text

    def detokenize(self, prompt_tail: List[int], ids: List[int]) -> str:
        return self.detokenize_batch([(prompt_tail, ids)])[0]

    def detokenize_batch(self, pairs) -> List[str]:
        """Generated text for [(prompt tail ids, generated ids)], as the reference worker returns it: the
        CONTINUATION of the prompt's text — vLLM primes its DecodeStream with the prompt ids
        (vllm/v1/engine/detokenizer.py:181-184), so e.g. a word-level / metaspace vocabulary yields
        " w5 w6", not "w5 w6".  Fast path: decode(prompt tail + ids) minus decode(tail) — identical to
        stepping the stream unless the tail ends inside a multi-byte character, in which case the stream
        itself is stepped.  All the requests a step finished are decoded in two batch calls (Rust
        thread pool, GIL released)."""
        pairs = [(list(t), list(i)) for t, i in pairs]
        if self.backend is not None and len(pairs) > 1:
            pres = self.backend.decode_batch([t for t, _ in pairs], skip_special_tokens=True)
            fulls = self.backend.decode_batch([t + i for t, i in pairs], skip_special_tokens=True)
        else:
            pres = [self.decode(t) if t else "" for t, _ in pairs]
            fulls = [self.decode(t + i) for t, i in pairs]
        out = []
        for (tail, ids), pre, full in zip(pairs, pres, fulls):
            if pre.endswith("\ufffd") and self.backend is not None:
                from tokenizers.decoders import DecodeStream

                stream, backend = DecodeStream(ids=tail, skip_special_tokens=True), self.backend
                out.append("".join(filter(None, (stream.step(backend, t) for t in ids))))
            else:
                out.append(full[len(pre):])
        return out

    async def encode_async(self, text: str) -> List[int]:
        """`encode` for callers on an event loop: every prompt submitted in the same loop iteration (the
        broker delivers in bursts; a finished step frees a burst of prefetch slots) is tokenised in ONE
        `encode_batch` call — same ids, ~5x less loop-thread time per job, GIL released meanwhile."""
        if self.backend is None:
            return self.encode(text)
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        pending = self._enc_pending.setdefault(loop, [])
        pending.append((text, fut))
        if len(pending) == 1:
            loop.call_soon(self._flush_encodes, loop)
        return await fut

    def _flush_encodes(self, loop) -> None:
        pending = self._enc_pending.pop(loop, [])
        if not pending:
            return
        try:
            encs = self.backend.encode_batch([t for t, _ in pending], add_special_tokens=True)
        except Exception as e:  # one bad text must not strand the others: fall back to one by one
            for text, fut in pending:
                if fut.done():
                    continue
                try:
                    fut.set_result(self.encode(text))
                except Exception as e1:
                    fut.set_exception(e1)
            return
        for (_, fut), enc in zip(pending, encs):
            if not fut.done():
                fut.set_result(enc.ids)

    def start(self):
        # The engine thread re-acquires the GIL after every device step; a busy event-loop thread
        # (tokenising, pydantic) hands it over only at the interpreter's switch interval, 5 ms by
        # default — comparable to a whole step of a small model.  0.2 ms bounds that wait.
        import sys

        sys.setswitchinterval(min(sys.getswitchinterval(), 2e-4))
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread.is_alive():
            self._thread.join(timeout=30)

    # ---- called on the event-loop thread ----
    def submit(self, prompt_ids: List[int], max_new: int, stop: Optional[List[str]],
               loop: asyncio.AbstractEventLoop, temperature: float = 0.0,
               seed: Optional[int] = None) -> "asyncio.Future":
        """temperature 0 = greedy; > 0 = softmax(logits/T) sampling (the reference default is 0.7).
        seed None = a fresh stream per request (unseeded, like the reference)."""
        if self.error is not None:
            raise RuntimeError(f"engine thread died: {self.error!r}")
        import numpy as np

        fut = loop.create_future()
        # int32 array built here, on the caller's thread: the engine thread only passes its pointer on
        prompt_ids = np.ascontiguousarray(prompt_ids, dtype=np.int32)
        self._next_id += 1
        if seed is None:
            seed = (self._base_seed + 0x9E3779B97F4A7C15 * self._next_id) & 0xFFFFFFFFFFFFFFFF
        self._inbox.put(_Req(prompt_ids, max_new, stop, fut, loop, temperature, seed))
        return fut

    # ---- engine thread ----
    def _finish(self, batch, r: _Req, text: Optional[str] = None, exc: Optional[BaseException] = None):
        n_out = 0
        if r.slot >= 0:
            self._reqs.pop(r.slot, None)
            n_out = int(self._table.n[r.slot])
            if exc is None and text is None:
                # hand the ids over; detokenisation happens in _deliver on the caller's loop thread,
                # concurrently with the next device step instead of between two steps
                text = (r.prompt_ids[-self.CONTEXT:].tolist(), self._table.tokens(r.slot))
            self._table.release(r.slot)
            r.slot = -1
        self.jobs_done += 1
        self.tokens_out += n_out
        self.last_finish_t = time.perf_counter()
        batch.setdefault(r.loop, []).append((r.future, text, exc, n_out))

    def _deliver(self, items):
        """runs on the event-loop thread of the requests in `items`; `text` is either the final
        string (stop-string cut) or (prompt tail ids, generated ids) still to be detokenised"""
        todo = [(k, it[1]) for k, it in enumerate(items)
                if it[2] is None and not isinstance(it[1], str) and not it[0].done()]
        texts = {}
        if todo:
            try:
                texts = dict(zip((k for k, _ in todo), self.detokenize_batch([p for _, p in todo])))
            except Exception:
                texts = {}  # fall through: per-item decoding below reports the failing one
        for k, (fut, text, exc, n) in enumerate(items):
            if fut.done():
                continue
            if exc is not None:
                fut.set_exception(exc)
                continue
            try:
                if not isinstance(text, str):
                    text = texts[k] if k in texts else self.detokenize(*text)
                fut.set_result((text, n))
            except Exception as e:  # a detokeniser failure must not strand the waiter
                fut.set_exception(e)

    def _check_stop_strings(self, r: _Req) -> Optional[str]:
        """vLLM detokenizer semantics (vllm/v1/engine/detokenizer.py:131-143): after every new
        token look for a stop string in the newly produced text (plus the hold-back window); the
        output is cut before it.  Incremental: see _StopScanner."""
        n = int(self._table.n[r.slot])
        sc = r.detok
        cut = None
        while sc.n < n and cut is None:  # normally exactly one new token
            cut = sc.push(int(self._table.tok[r.slot, sc.n]))
        return cut

    def _admit(self, batch, r: _Req) -> None:
        """engine thread: give the request a table row and hand it to the engine.  Anything wrong
        with THIS request (oversized prompt, absurd max_tokens) fails this request only — as
        ValueError, which the base class turns into log + ack/drop (ref:llmq/workers/base.py:228-235)."""
        try:
            limit = int(getattr(self.engine, "max_model_len", 0) or 0)
            max_new = int(r.max_new)
            if max_new < 1:
                raise ValueError(f"max_tokens must be >= 1 (got {max_new})")
            if limit > 0:  # the engine clamps to max_model_len - n_prompt anyway; never size rows beyond it
                max_new = min(max_new, limit)
            r.max_new = min(max_new, 2 ** 31 - 1)
            r.slot = self._table.alloc(r.max_new, bool(r.stop))
            if r.stop:
                r.detok = _StopScanner(r.stop, self.decode, self.backend, r.prompt_ids.tolist())
            self.engine.add_request(r.slot, r.prompt_ids, r.max_new, ignore_eos=False,
                                    temperature=r.temperature, seed=r.seed)
            self._reqs[r.slot] = r
        except (ValueError, OverflowError, MemoryError) as e:  # un-servable job => dropped by the base class
            self._finish(batch, r, exc=e if isinstance(e, ValueError) else ValueError(f"un-servable request: {e!r}"))

    def _run(self):
        eng = self.engine
        try:
            if getattr(eng.model, "device", None) is not None:
                import torch

                torch.cuda.set_device(eng.model.device)  # this thread owns the CUDA context
            r = None
            batch: dict = {}
            while not self._stop.is_set():
                batch = {}
                r = None
                # admit new requests
                try:
                    block = not eng.has_work()
                    r = self._inbox.get(timeout=0.05) if block else self._inbox.get_nowait()
                    while True:
                        if self.first_submit_t is None:
                            self.first_submit_t = time.perf_counter()
                        self._admit(batch, r)
                        r = None
                        r = self._inbox.get_nowait()
                except queue.Empty:
                    pass
                if eng.has_work():
                    ids, toks, flags = eng.step()
                    table = self._table
                    table.append_step(ids, toks)  # vectorised: no per-token Python work
                    # per-request work only for the few that carry stop strings or finished this step
                    stop_hit = set()
                    if table.has_stop.any():
                        for slot, flag in zip(ids[table.has_stop[ids]].tolist(),
                                              flags[table.has_stop[ids]].tolist()):
                            r = self._reqs.get(slot)
                            if r is None:
                                continue
                            cut = self._check_stop_strings(r)
                            if cut is not None:
                                if not flag:
                                    eng.abort(slot)
                                stop_hit.add(slot)
                                self._finish(batch, r, text=cut)
                    for slot in ids[flags != 0].tolist():
                        r = self._reqs.get(slot)
                        if r is not None and slot not in stop_hit:
                            self._finish(batch, r)
                for loop, items in batch.items():
                    try:
                        loop.call_soon_threadsafe(self._deliver, items)
                    except RuntimeError:
                        pass  # that event loop has been closed: its waiters are gone, nothing to deliver to
        except BaseException as e:  # surface engine failures to every waiter
            self.error = e   # submit() refuses new work from here on
            # (`batch` may already hold results finished in this iteration: they are still delivered)
            err = RuntimeError(f"engine failure: {e!r}")
            pending = list(self._reqs.values())
            if r is not None and r.slot not in self._reqs:
                pending.append(r)   # the request in hand when the failure hit
            while True:             # ... and everything still queued behind it
                try:
                    pending.append(self._inbox.get_nowait())
                except queue.Empty:
                    break
            for q in pending:
                if not q.future.done():
                    self._finish(batch, q, exc=err)
            if self.on_fatal is not None:   # before the waiters wake up: they must see the worker stopping
                try:
                    self.on_fatal(e)
                except Exception:
                    pass
            for loop, items in batch.items():
                try:
                    loop.call_soon_threadsafe(self._deliver, items)
                except RuntimeError:   # that loop is already closed
                    pass
            raise


def collect_stop_ids(model_dir: str, spec, tokenizer) -> List[int]:
    """Every token id that ends generation for this model, as vLLM derives it for the reference
    worker: tokenizer.eos_token_id, plus every eos id of config.json, plus every eos id of
    generation_config.json (vllm/sampling_params.py update_from_generation_config adds the latter
    as stop_token_ids).  Llama-3.x-Instruct: [128001, 128008, 128009]; gemma-2-it: [1, 107]."""
    import json

    ids: List[int] = []

    def add(v):
        for x in (v if isinstance(v, (list, tuple)) else [v]):
            if x is not None and int(x) not in ids:
                ids.append(int(x))

    add(getattr(tokenizer, "eos_token_id", None))
    add(list(spec.eos_token_ids) or spec.eos_token_id)
    gpath = os.path.join(model_dir, "generation_config.json")
    if os.path.exists(gpath):
        try:
            with open(gpath) as f:
                add(json.load(f).get("eos_token_id"))
        except (OSError, ValueError):
            pass
    return [i for i in ids if 0 <= i < spec.vocab]


def build_service(model_name: str, *, max_num_seqs: Optional[int], max_model_len: Optional[int],
                  gpu_memory_utilization: float, max_num_batched_tokens: Optional[int] = None,
                  seed: int = 1234, logger=None, num_blocks: Optional[int] = None) -> GenerationService:
    """model + KV pool + engine + tokenizer for `model_name` (local dir or random:<builtin>)."""
    import torch

    from .model import Engine, NativeModel, fuse_hf_weights, load_hf_state_dict, random_engine_weights, resolve_model

    L.require_device()  # raises: this worker has no CPU path
    spec, model_dir = resolve_model(model_name)
    device = torch.device("cuda", torch.cuda.current_device())
    max_num_seqs = max_num_seqs or int(os.environ.get("B200Q_MAX_NUM_SEQS", "256"))
    # VLLM_MAX_MODEL_LEN unset: the model's own context length, like vLLM (the reference leaves it
    # to the engine, ref:llmq/workers/vllm_worker.py:119-120) — capped by what the KV pool can hold
    # (NativeModel) instead of refusing to start
    env_len = os.environ.get("B200Q_MAX_MODEL_LEN")
    max_model_len = min(max_model_len or (int(env_len) if env_len else spec.max_position_embeddings),
                        spec.max_position_embeddings)
    budget = max_num_batched_tokens or int(os.environ.get("B200Q_MAX_NUM_BATCHED_TOKENS", "4096"))
    budget = max(budget, 16)
    if model_dir is None:
        from .fixtures import build_tokenizer, special_token_ids

        weights = random_engine_weights(spec, seed, device)
        tokenizer = build_tokenizer(spec.vocab)
        stop_ids = [special_token_ids(spec.vocab)["<|end_of_text|>"]]
    else:
        from transformers import AutoTokenizer

        weights = fuse_hf_weights(spec, load_hf_state_dict(model_dir))
        tokenizer = AutoTokenizer.from_pretrained(model_dir)
        stop_ids = collect_stop_ids(model_dir, spec, tokenizer)
    model = NativeModel(spec, weights, max_tokens=budget, max_seqs=max_num_seqs,
                        max_model_len=max_model_len, gpu_memory_utilization=gpu_memory_utilization,
                        device=device, num_blocks=num_blocks)
    if logger:
        if model.max_model_len < max_model_len:
            logger.warning(f"b200q: max_model_len reduced from {max_model_len} to {model.max_model_len} "
                           f"(KV pool of {model.num_blocks} blocks / attention window); longer prompts are dropped")
        logger.info(f"b200q model {spec.name}: {model.num_blocks} KV blocks of 16 tokens, "
                    f"max_num_seqs={max_num_seqs}, token budget={budget}, max_model_len={model.max_model_len}, "
                    f"stop ids={stop_ids}")
    engine = Engine(model, max_num_seqs=max_num_seqs, max_num_batched_tokens=budget,
                    max_model_len=model.max_model_len, eos_token_id=stop_ids or None)
    eos = stop_ids[0] if stop_ids else None
    return GenerationService(engine, tokenizer, eos)
