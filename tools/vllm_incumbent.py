"""Run the incumbent — vLLM 0.22 (the engine behind the reference's VLLMWorker,
ref:llmq/workers/vllm_worker.py:105-123,161-186) — on the same B200, development/measurement tool.

  python tools/vllm_incumbent.py golden     # greedy token ids on a small seeded checkpoint
                                            #   -> gpurun_out/vllm_golden_<name>.json (copied to tests/golden/)
  python tools/vllm_incumbent.py throughput # Llama-3-8B dims, dummy weights, 128-in/128-out

Differences from the reference worker, all deliberate: temperature is forced to 0
(the reference hard-codes 0.7 unseeded), prompts are passed as token ids (same ids the reference
would get from tokenising the text), and the engine is the synchronous `LLM` front-end over the
same EngineCore instead of `AsyncLLMEngine` behind an AMQP consumer.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VLLM_ENABLE_V1_MULTIPROCESSING", "0")
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

import numpy as np  # noqa: E402

from llmq_b200.fixtures import write_model_dir  # noqa: E402
from llmq_b200.model import BUILTIN_SPECS, ModelSpec  # noqa: E402

GOLDEN_SPECS = {
    "d128": ModelSpec(hidden=512, n_layers=3, n_q_heads=8, n_kv_heads=2, head_dim=128, intermediate=1024,
                      vocab=2048, max_position_embeddings=512, name="tiny-d128"),
    "d64": ModelSpec(hidden=512, n_layers=2, n_q_heads=8, n_kv_heads=2, head_dim=64, intermediate=1024,
                     vocab=2048, max_position_embeddings=512, name="tiny-d64"),
}
GOLDEN_SEED = 4321
PROMPT_LENS = [5, 16, 17, 40, 130, 64, 33, 2, 100, 77, 128, 12]


def golden_prompts(vocab):
    g = np.random.default_rng(99)
    return [g.integers(3, vocab, size=n).tolist() for n in PROMPT_LENS]


def golden():
    from vllm import LLM, SamplingParams

    for name, spec in GOLDEN_SPECS.items():
        d = os.path.join("/tmp", f"b200q_golden_{name}")
        write_model_dir(d, spec, seed=GOLDEN_SEED, with_weights=True)
        res = {"name": name, "spec": spec.to_hf_config(), "weights_seed": GOLDEN_SEED, "max_new_tokens": 24,
               "prompts": golden_prompts(spec.vocab), "runs": {}}
        for tag, kw in (("default", {}), ("eager", {"enforce_eager": True})):
            try:
                llm = LLM(model=d, dtype="bfloat16", max_model_len=512, max_num_seqs=16, gpu_memory_utilization=0.3,
                          skip_tokenizer_init=True, disable_log_stats=True, seed=0, **kw)
                sp = SamplingParams(temperature=0.0, max_tokens=24, ignore_eos=True, detokenize=False)
                outs = llm.generate([{"prompt_token_ids": p} for p in res["prompts"]], sp, use_tqdm=False)
                res["runs"][tag] = [list(o.outputs[0].token_ids) for o in outs]
                del llm
            except Exception as e:  # keep going: one mode is enough to pin
                res["runs"][tag] = {"error": repr(e)[:500]}
            import gc

            import torch
            gc.collect()
            torch.cuda.empty_cache()
        import vllm
        res["vllm"] = vllm.__version__
        with open(os.path.join(OUT, f"vllm_golden_{name}.json"), "w") as f:
            json.dump(res, f)
        print(name, {k: (v if isinstance(v, dict) else len(v)) for k, v in res["runs"].items()})


def throughput():
    from vllm import LLM, SamplingParams

    from llmq_b200.fixtures import make_jobs, build_tokenizer

    spec = BUILTIN_SPECS["llama-3-8b"]
    d = "/tmp/b200q_llama3_8b_cfg"
    write_model_dir(d, spec, with_weights=False)
    n_jobs = int(os.environ.get("JOBS", "4608"))
    seqs = int(os.environ.get("MAX_NUM_SEQS", "1024"))
    tok = build_tokenizer(spec.vocab)
    prompts = [{"prompt_token_ids": tok(j["prompt"], add_special_tokens=True).input_ids}
               for j in make_jobs(2 * n_jobs, spec.vocab)]
    results = []
    for tag, kw in (("eager", {"enforce_eager": True}), ("default", {})):
        if os.environ.get("ONLY") and os.environ["ONLY"] != tag:
            continue
        try:
            t0 = time.time()
            llm = LLM(model=d, load_format="dummy", dtype="bfloat16", max_model_len=512, max_num_seqs=seqs,
                      gpu_memory_utilization=0.9, skip_tokenizer_init=True, disable_log_stats=True, seed=0, **kw)
            init_s = time.time() - t0
            sp = SamplingParams(temperature=0.0, max_tokens=128, ignore_eos=True, detokenize=False)
            llm.generate(prompts[:256], sp, use_tqdm=False)  # warm-up
            t0 = time.time()
            outs = llm.generate(prompts[n_jobs:2 * n_jobs], sp, use_tqdm=False)
            dt = time.time() - t0
            ntok = sum(len(o.outputs[0].token_ids) for o in outs)
            r = {"kind": "vllm_throughput", "mode": tag, "jobs": n_jobs, "max_num_seqs": seqs, "seconds": dt,
                 "out_tokens_per_s": ntok / dt, "jobs_per_s": n_jobs / dt, "init_s": init_s}
            del llm
        except Exception as e:
            r = {"kind": "vllm_throughput", "mode": tag, "error": repr(e)[:800]}
        print(json.dumps(r))
        results.append(r)
        import gc

        import torch
        gc.collect()
        torch.cuda.empty_cache()
    with open(os.path.join(OUT, "vllm_throughput.json"), "w") as f:
        json.dump(results, f)


BUDGETS = {128: 2048, 750: 4096, 4608: 4608}


def _jobs(seqs, jobs_x):
    from llmq_b200.fixtures import build_tokenizer, make_jobs

    spec = BUILTIN_SPECS["llama-3-8b"]
    tok = build_tokenizer(spec.vocab)
    ids = [tok(j["prompt"], add_special_tokens=True).input_ids for j in make_jobs(jobs_x * seqs + 256, spec.vocab)]
    return ids[:256], ids[256:]


def one():
    """one engine, one max_num_seqs, in its own process (vLLM keeps its 0.9 x 180 GB until exit):
    python tools/vllm_incumbent.py one <vllm|native> <max_num_seqs> <jobs_x>"""
    import torch

    engine, seqs, jobs_x = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    warm, ids = _jobs(seqs, jobs_x)
    n_jobs = len(ids)
    if engine == "vllm":
        from vllm import LLM, SamplingParams

        d = "/tmp/b200q_llama3_8b_cfg"
        write_model_dir(d, BUILTIN_SPECS["llama-3-8b"], with_weights=False)
        t0 = time.time()
        llm = LLM(model=d, load_format="dummy", dtype="bfloat16", max_model_len=512, max_num_seqs=seqs,
                  gpu_memory_utilization=0.9, skip_tokenizer_init=True, disable_log_stats=True, seed=0)
        init_s = time.time() - t0
        sp = SamplingParams(temperature=0.0, max_tokens=128, ignore_eos=True, detokenize=False)
        llm.generate([{"prompt_token_ids": p} for p in warm], sp, use_tqdm=False)
        torch.cuda.synchronize()
        t0 = time.time()
        outs = llm.generate([{"prompt_token_ids": p} for p in ids], sp, use_tqdm=False)
        dt = time.time() - t0
        ntok = sum(len(o.outputs[0].token_ids) for o in outs)
        r = {"engine": "vllm-0.22.0 (LLM.generate, compiled + CUDA graphs, its default token budget, dummy weights)",
             "init_s": round(init_s, 1)}
    else:
        from llmq_b200.service import build_service

        svc = build_service("random:llama-3-8b", max_num_seqs=seqs, max_model_len=512, gpu_memory_utilization=0.9,
                            max_num_batched_tokens=BUDGETS.get(seqs, 4608), seed=1234)
        eng = svc.engine
        for i, p in enumerate(warm):
            eng.add_request(i, p, 128, ignore_eos=True)
        while eng.has_work():
            eng.step()
        torch.cuda.synchronize()
        t0 = time.time()
        for i, p in enumerate(ids):
            eng.add_request(i, p, 128, ignore_eos=True)
        ntok = 0
        while eng.has_work():
            ntok += len(eng.step()[1])
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = eng.stats()
        r = {"engine": "b200q (engine-direct, random N(0,0.02) weights)", "max_num_batched_tokens": BUDGETS.get(seqs, 4608),
             "engine_steps": int(st.steps), "preemptions": int(st.preemptions)}
    r.update({"max_num_seqs": seqs, "jobs": n_jobs, "seconds": round(dt, 3), "out_tokens_per_s": round(ntok / dt, 1),
              "jobs_per_s": round(n_jobs / dt, 2)})
    print("RESULT " + json.dumps(r), flush=True)


def same_lease():
    """vLLM 0.22 and the native engine back to back on ONE lease, identical protocol: N = JOBS_X x
    max_num_seqs canonical 128-token jobs queued at once, 128 greedy tokens each (ignore_eos), wall
    clock from submission to the last token (ramp and tail included), for every max_num_seqs in
    MAX_NUM_SEQS (default 128,750,4608: the engine default the reference runs with, its production
    scripts' 750 (ref:utils/run_llmq_benchmark.slurm:32-33) and the headline configuration).
    Every run is its own process.  -> gpurun_out/same_lease_vllm_vs_native.json"""
    import subprocess

    seqs_list = [int(x) for x in os.environ.get("MAX_NUM_SEQS", "128,750,4608").split(",")]
    jobs_x = os.environ.get("JOBS_X", "4")
    results = []
    for seqs in seqs_list:
        for engine in ("vllm", "native"):
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", engine, str(seqs), jobs_x],
                               capture_output=True, text=True, timeout=1500)
            line = next((l for l in p.stdout.splitlines() if l.startswith("RESULT ")), None)
            r = json.loads(line[7:]) if line else {"engine": engine, "max_num_seqs": seqs, "error": (p.stderr or p.stdout)[-800:]}
            print(json.dumps(r), flush=True)
            results.append(r)
            with open(os.path.join(OUT, "same_lease_vllm_vs_native.json"), "w") as f:
                json.dump(results, f, indent=1)


if __name__ == "__main__":
    {"golden": golden, "throughput": throughput, "same_lease": same_lease, "one": one}[sys.argv[1]]()
