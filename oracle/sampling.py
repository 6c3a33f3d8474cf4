"""oracle/sampling.py — CPU restatement of the sampler the reference uses by default.
TEST INFRASTRUCTURE ONLY (see oracle/ops.py).

Reference behaviour: `SamplingParams(temperature=0.7, ...)` with no seed, no top-p/top-k
(ref:llmq/workers/vllm_worker.py:161-165) which vLLM executes as
    probs = softmax(logits.float() / T);  q = empty_like(probs).exponential_();
    token = probs.div_(q).argmax(-1)                (vllm/v1/sample/ops/topk_topp_sampler.py:395-416)
i.e. an exponential race / Gumbel-max draw from softmax(logits / T).  torch's generator stream
cannot be reproduced outside torch, so bit parity with vLLM is impossible by construction; what is
pinned is (a) the distribution (chi-square test against softmax(logits/T)) and (b) bit parity of
the CUDA kernel with THIS restatement, which uses the kernel's documented random stream:
q = -log(u), u = (x >> 8 + 0.5) / 2^24, x = Philox4x32-10(key = seed, counter = (v // 4, position, 0, 0))[v % 4].
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10 (Salmon et al., Random123); all inputs uint32 arrays/scalars"""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def gumbel_noise(V: int, seed: int, position: int) -> np.ndarray:
    """-log(q_v) for v in [0, V): float64"""
    groups = np.arange((V + 3) // 4, dtype=np.uint32)
    r = philox4x32_10(groups, np.uint32(position), np.uint32(0), np.uint32(0),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.stack(r, axis=1).reshape(-1)[:V]
    u = ((bits >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    return -np.log(-np.log(u))


def sample(logits: np.ndarray, temperature: float, seed: int, position: int) -> int:
    """logits: [V] values of the bf16 logits row (as float). temperature 0 => greedy first-max."""
    logits = np.asarray(logits, dtype=np.float64)
    if not temperature > 0:
        return int(np.argmax(logits))
    return int(np.argmax(logits / np.float32(temperature).astype(np.float64) + gumbel_noise(len(logits), seed, position)))


def sample_margin(logits, temperature, seed, position) -> float:
    """gap between the winning and the runner-up perturbed value (near-ties are where fp32 vs fp64
    log() may pick differently)"""
    v = np.asarray(logits, dtype=np.float64) / float(np.float32(temperature)) + gumbel_noise(len(logits), seed, position)
    top = np.partition(v, -2)[-2:]
    return float(top[1] - top[0])
