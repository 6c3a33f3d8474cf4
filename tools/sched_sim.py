"""Scheduler study without a GPU: drive the REAL C++ engine (dry-run mode: scheduler, paged-KV block
manager, chunked prefill, preemption) with a synthetic job mix and price every step with a
two-term roofline model (GEMM time from tokens per step, attention time from KV bytes read), using
the peaks measured on B200 (MEASURED_PEAKS.json) and the efficiencies measured by bench.py.
Answers "what does the scheduler leave on the table for this workload" — e.g. BASELINE config #5
(prompts ~U{32..2048}, outputs ~U{1..512}) — before spending GPU time on it.

    python tools/sched_sim.py --model gemma-2-9b --jobs 4000 --mix config5 --policy 1
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--jobs", type=int, default=4000)
    ap.add_argument("--mix", choices=["128x128", "config5"], default="config5")
    ap.add_argument("--policy", type=int, default=1)
    ap.add_argument("--max-num-seqs", type=int, default=1024)
    ap.add_argument("--budget", type=int, default=4608)
    ap.add_argument("--max-model-len", type=int, default=2560 + 16)
    ap.add_argument("--gpu-mem-gb", type=float, default=0.92 * 178.4)
    ap.add_argument("--gemm-eff", type=float, default=0.88, help="fraction of the sustained bf16 peak (bench.py)")
    ap.add_argument("--attn-eff", type=float, default=0.85, help="fraction of the HBM copy bandwidth (bench.py)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--quantise", action="store_true", help="price GEMMs by rows padded to the 128-row tile")
    args = ap.parse_args()

    from llmq_b200.fixtures import DryRunEngine
    from llmq_b200.model import BUILTIN_SPECS

    spec = BUILTIN_SPECS[args.model]
    peaks = {"bf16_tflops": 1460.0, "hbm_gbs": 6480.5}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"] = float(mp.get("hbm_copy_gbs", mp.get("hbm_gbs", peaks["hbm_gbs"])))
    except Exception:
        pass
    w_bytes = spec.weight_bytes_per_step()
    kv_tok = spec.kv_bytes_per_token()
    num_blocks = int((args.gpu_mem_gb * 1e9 - w_bytes - 4e9) // (kv_tok * 16))
    flops_per_token = w_bytes  # 2 flops per bf16 weight (2 bytes): numerically equal
    rng = np.random.default_rng(args.seed)
    if args.mix == "128x128":
        p_len = np.full(args.jobs, 128)
        o_len = np.full(args.jobs, 128)
    else:
        p_len = rng.integers(32, 2049, size=args.jobs)
        o_len = rng.integers(1, 513, size=args.jobs)

    eng = DryRunEngine(1000, args.max_num_seqs, args.budget, args.max_model_len, num_blocks, None, policy=args.policy)
    prompt = np.arange(1, 2100, dtype=np.int32) % 997 + 1
    for i in range(args.jobs):
        eng.add_request(i, prompt[: p_len[i]], int(o_len[i]), ignore_eos=True)

    # per-request context bookkeeping for the attention term (the engine reports only totals)
    t_gemm = t_attn = t_floor = 0.0
    steps = tokens = dec_rows = 0
    hist_T = []
    prev = eng.stats()
    gen = 0
    live_ctx_sum = 0.0  # running estimate: sum of contexts of decoding sequences
    ctx = {}
    while eng.has_work():
        ids, toks, flags = eng.step()
        st = eng.stats()
        T = st.last_step_tokens
        d_pre = st.tokens_prefilled - prev.tokens_prefilled
        d_dec = st.tokens_decoded - prev.tokens_decoded
        prev_pre = prev
        prev = st
        # attention bytes: every decode row reads its whole context; prefill chunks read theirs once per 16-row tile
        for rid in ids.tolist():
            ctx[rid] = ctx.get(rid, int(p_len[rid])) + 1
        dec_ctx = sum(ctx[r] for r in ids.tolist()) if d_dec else 0
        for rid, f in zip(ids.tolist(), flags.tolist()):
            if f:
                ctx.pop(rid, None)
        attn_bytes = dec_ctx * kv_tok + d_pre * 1024 * kv_tok / 16.0 * 0.5  # prefill: ~mean ctx 1024, 16 rows share a read
        # the persistent GEMMs work in 128-row tiles: a step costs what its padded row count costs
        Tq = -(-T // 128) * 128 if args.quantise else T
        g = max(Tq * flops_per_token / (peaks["bf16_tflops"] * 1e12 * args.gemm_eff), w_bytes / (peaks["hbm_gbs"] * 1e9))
        a = attn_bytes / (peaks["hbm_gbs"] * 1e9 * args.attn_eff)
        t_gemm += g
        t_attn += a
        steps += 1
        tokens += T
        gen += len(ids)
        hist_T.append(T)
    total = t_gemm + t_attn
    all_tok = int(p_len.sum() + o_len.sum())
    ideal = all_tok * flops_per_token / (peaks["bf16_tflops"] * 1e12 * args.gemm_eff)
    st = eng.stats()
    hist_T = np.array(hist_T)
    print(json.dumps({
        "model": args.model, "mix": args.mix, "policy": args.policy, "jobs": args.jobs,
        "max_num_seqs": args.max_num_seqs, "budget": args.budget, "kv_blocks": num_blocks,
        "steps": steps, "preemptions": int(st.preemptions),
        "tokens_per_step_mean": round(float(hist_T.mean()), 1),
        "steps_below_half_budget": round(float((hist_T < args.budget / 2).mean()), 3),
        "modelled_seconds": round(total, 2), "gemm_seconds": round(t_gemm, 2), "attn_seconds": round(t_attn, 2),
        "gemm_only_lower_bound_seconds": round(ideal, 2),
        "modelled_jobs_per_sec": round(args.jobs / total, 1),
        "modelled_output_tok_per_sec": round(float(o_len.sum()) / total, 1),
        "note": "roofline-priced steps of the real scheduler; not a measurement"}))


if __name__ == "__main__":
    main()
