"""The reference's CLI surface end to end across PROCESSES, CPU only: a TCP broker (the aio_pika
stand-in's server), two `llmq worker dummy` processes (the reference's unmodified worker + CLI),
`llmq submit` of a JSONL file and `llmq receive` of the results — queue-sharded data parallelism
exactly as llmq deploys it (competing consumers on one queue)."""
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "tests", "shims")
REF = next((p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference") if os.path.isdir(os.path.join(p, "llmq"))), None)


@pytest.mark.skipif(REF is None, reason="reference llmq package not available")
def test_two_worker_processes_share_one_queue(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, REF, ROOT]), LLMQ_LOG_LEVEL="INFO")
    ready = tmp_path / "port"
    procs = []
    try:
        srv = subprocess.Popen([sys.executable, "-m", "aio_pika.server", "--port", "0", "--ready-file", str(ready)], env=env)
        procs.append(srv)
        for _ in range(200):
            if ready.exists() and ready.read_text():
                break
            time.sleep(0.05)
        env["B200Q_SHIM_BROKER"] = f"127.0.0.1:{ready.read_text()}"
        logs = []
        for i in range(2):
            lf = open(tmp_path / f"worker{i}.log", "w")
            logs.append(lf)
            procs.append(subprocess.Popen([sys.executable, "-m", "llmq", "worker", "dummy", "mpq", "-c", "30"],
                                          env=env, stdout=lf, stderr=subprocess.STDOUT))
        # the reference's own readiness signal (ref:performance_benchmark.py:510 greps for it)
        deadline = time.time() + 60
        while time.time() < deadline:
            if all("starting to consume from queue" in open(tmp_path / f"worker{i}.log").read() for i in range(2)):
                break
            time.sleep(0.1)
        else:
            pytest.fail("workers did not start: " + open(tmp_path / "worker0.log").read()[-2000:])
        jobs = tmp_path / "jobs.jsonl"
        with open(jobs, "w") as f:
            for i in range(80):
                f.write(json.dumps({"id": f"job-{i:07d}", "prompt": "Echo {text}", "text": f"t{i}"}) + "\n")
        sub = subprocess.run([sys.executable, "-m", "llmq", "submit", "mpq", str(jobs)], env=env,
                             capture_output=True, text=True, timeout=60)
        assert sub.returncode == 0, sub.stderr[-2000:]
        rec = subprocess.run([sys.executable, "-m", "llmq", "receive", "mpq", "--timeout", "4"], env=env,
                             capture_output=True, text=True, timeout=90)
        assert rec.returncode == 0, rec.stderr[-2000:]
        results = [json.loads(l) for l in rec.stdout.splitlines() if l.startswith("{")]
        assert sorted(r["id"] for r in results) == [f"job-{i:07d}" for i in range(80)]
        assert all(r["result"] == "echo " + r["text"] and r["prompt"] == "Echo " + r["text"] for r in results)
        workers = {r["worker_id"] for r in results}
        assert len(workers) == 2, f"jobs were not shared between the two worker processes: {workers}"
    finally:
        for p in procs[::-1]:
            p.terminate()
        for p in procs:
            try:
                p.wait(5)
            except Exception:
                p.kill()


def test_queue_bench_tool_pipeline_mode_with_the_reference_dummy_worker(tmp_path):
    """tools/queue_bench.py (the process-level harness of BASELINE configs #3-#5) in its two-stage pipeline
    mode, with the reference's own DummyWorker processes standing in for the GPU workers: broker process,
    `llmq worker pipeline CFG STAGE` processes, stage-1 results routed to stage 2 by the reference's
    publish_pipeline_result, final results counted from `pipeline.<name>.results`."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "queue_bench.py"), "--worker-kind", "dummy",
                        "--pipeline", "1+1", "--jobs", "12", "--prefetch", "20", "--out-dir", str(tmp_path), "--timeout", "120"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["config"] == "#4 two-stage pipeline" and d["jobs"] == 12 and d["workers"] == "1+1"
    q = d["queues_at_end"]
    assert q["pipeline.qbench.stage1"]["delivered"] == 12 and q["pipeline.qbench.stage2"]["delivered"] == 12
