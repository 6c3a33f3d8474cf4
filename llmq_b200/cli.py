"""`llmq-b200` — the reference CLI with the native worker plugged into the vLLM worker's slot.

llmq has no plugin registry: its launchers do `from llmq.workers.vllm_worker import VLLMWorker`
lazily (ref:llmq/cli/worker.py:20,182).  `install()` therefore (a) publishes a module under that
name whose `VLLMWorker` is `B200Worker`, so `llmq worker run MODEL QUEUE` and pipeline stages with
`worker: vllm` resolve to the native worker with llmq's core and CLI files untouched, and (b) adds
an explicit `llmq worker b200 MODEL QUEUE` command.  Everything else (`submit`, `receive`,
`status`, ...) is the reference's own code.

    python -m llmq_b200.cli worker run /models/llama-3-8b translation-queue
    python -m llmq_b200.cli worker b200 random:llama-3-8b bench-queue
"""
from __future__ import annotations

import asyncio
import sys
import types


def install() -> None:
    from .worker import B200Worker

    mod = types.ModuleType("llmq.workers.vllm_worker")
    mod.VLLMWorker = B200Worker
    mod.__doc__ = "llmq_b200 shim: VLLMWorker is the native B200Worker"
    sys.modules["llmq.workers.vllm_worker"] = mod
    import llmq.workers as W

    W.VLLMWorker = B200Worker
    W.B200Worker = B200Worker

    import click
    from llmq.cli import main as M

    if "b200" not in M.worker.commands:

        @M.worker.command("b200")
        @click.argument("model_name")
        @click.argument("queue_name")
        def worker_b200(model_name: str, queue_name: str):
            """Run the native B200 worker (one replica per visible GPU process)"""
            w = B200Worker(model_name, queue_name, tensor_parallel_size=1)
            asyncio.run(w.run())


def main() -> None:
    install()
    from llmq.cli.main import cli

    cli()


if __name__ == "__main__":
    main()
