"""bench.py's own host logic, without a GPU: `--dry-run` swaps the model for the C++ scheduler's
self-test engine (fabricated tokens) and CUDA events for the host clock; everything else — the
staggered ramp cohort, the backlog that never empties, step accounting by completed jobs, the
un-acked window of the e2e arm, the max/sum reductions across ranks (gloo), the JSON contract —
is the code the GPU run executes.  The line is marked `dry_run` and is never a measurement."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--dry-run", "--steps", "3", "--warmup", "2", "--jobs", "64", "--max-num-seqs", "256",
        "--max-num-batched-tokens", "256", "--max-model-len", "512", "--no-cpu-baseline"]
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}


def _line(out: str) -> dict:
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, f"bench must print exactly one JSON line, got {len(lines)}:\n{out[-2000:]}"
    return json.loads(lines[0])


def _check(d, n):
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["dry_run"] and d["n_gpus"] == n and d["steps"] == 3 and d["warmup"] == 2
    assert d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None
    eng = d["engine"]
    # steady state: the batch is still full and the backlog non-empty when the clock stops
    assert eng["running_at_end"] > 128 and eng["backlog_at_end"] > 0, eng
    # K steps of J completed jobs per rank; 128 tokens per completed job on average in steady state
    jobs = d["jobs_per_sec"] * d["ms_per_step"] / 1e3 * d["steps"]
    assert abs(jobs - 3 * 64 * n) <= 8 * n, jobs          # a step boundary may overshoot by one engine step
    toks = d["value"] * d["ms_per_step"] / 1e3 * d["steps"]
    assert 0.6 * 128 < toks / (3 * 64 * n) < 1.6 * 128, toks
    assert d["e2e"]["steps"] == 3 and d["e2e"]["value"] > 0 and set(d["e2e"]) >= {"h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["config"]["jobs_per_step_per_gpu"] == 64 and "workload" in d["config"]


def test_bench_host_logic_single_rank():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    _check(_line(r.stdout), 1)


def test_bench_host_logic_two_ranks_gloo():
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--gpus", "2"] + ARGS, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    _check(_line(r.stdout), 2)


def test_reference_arm_prints_the_same_config_and_is_bounded():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "llama-3.2-1b",
                        "--steps", "1", "--warmup", "0", "--prompt-tokens", "8", "--out-tokens", "6", "--cpu-budget-s", "2"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    # the reference arm reports the native arm's config verbatim (same workload, bounded sample of it)
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    ns = argparse.Namespace(model="llama-3.2-1b", prompt_tokens=8, out_tokens=6, max_num_seqs=4608, jobs=1152,
                            max_num_batched_tokens=4608, gpu_memory_utilization=0.92, gpus=1)
    assert d["config"] == bench.workload_config(ns)
