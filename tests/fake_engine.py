"""A CPU stand-in for llmq_b200.model.Engine used by the host-logic tests (no GPU): same call
surface (add_request / abort / has_work / step / stats / close), deterministic 'model':
next token = (previous token + 1) mod vocab."""
import numpy as np


class FakeModel:
    device = None

    def close(self):
        pass


class FakeEngine:
    def __init__(self, vocab=1024, max_num_seqs=8, max_model_len=64, eos_token_id=1):
        self.vocab, self.max_num_seqs, self.max_model_len, self.eos = vocab, max_num_seqs, max_model_len, eos_token_id
        self.waiting, self.running = [], []
        self.model = FakeModel()
        self.steps = 0
        self.max_batch_seen = 0

    def add_request(self, rid, ids, max_new, ignore_eos=False, temperature=0.0, seed=0):
        self.last_sampling = (temperature, seed)
        ids = list(ids)
        if len(ids) >= self.max_model_len:
            raise ValueError(f"prompt of {len(ids)} tokens does not fit max_model_len={self.max_model_len}")
        self.waiting.append({"id": rid, "last": ids[-1], "n": 0, "max": min(max_new, self.max_model_len - len(ids)),
                             "ignore_eos": ignore_eos})

    def abort(self, rid):
        self.waiting = [r for r in self.waiting if r["id"] != rid]
        self.running = [r for r in self.running if r["id"] != rid]

    def has_work(self):
        return bool(self.waiting or self.running)

    def step(self):
        while self.waiting and len(self.running) < self.max_num_seqs:
            self.running.append(self.waiting.pop(0))
        self.steps += 1
        self.max_batch_seen = max(self.max_batch_seen, len(self.running))
        ids, toks, flags = [], [], []
        for r in list(self.running):
            t = (r["last"] + 1) % self.vocab
            r["last"], r["n"] = t, r["n"] + 1
            f = 0
            if not r["ignore_eos"] and t == self.eos:
                f = 1
            elif r["n"] >= r["max"]:
                f = 2
            ids.append(r["id"]), toks.append(t), flags.append(f)
            if f:
                self.running.remove(r)
        return np.array(ids, dtype=np.int64), np.array(toks, dtype=np.int32), np.array(flags, dtype=np.int32)

    def close(self):
        pass
