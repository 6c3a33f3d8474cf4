"""B200Worker — the native worker that drops into llmq's worker plugin slot.

It subclasses the reference's ``BaseWorker`` (ref:llmq/workers/base.py:15) and implements the
four abstract methods (:57-75) with the same constructor signature, prompt construction, stop
handling and error convention as the reference ``VLLMWorker`` (ref:llmq/workers/vllm_worker.py)
— the job envelope, Result construction, publishing and acking stay in the untouched base class.

Threading model: the asyncio event-loop thread owns AMQP, tokenisation and
futures; ONE engine thread owns the CUDA context and runs the continuous-batching step loop
(`b200q_engine_step`, GIL released inside the C call); ``_process_job`` is
``submit(request); await future``.
"""
from __future__ import annotations

import asyncio
import os
from typing import List, Optional

from llmq.core.models import Job  # the reference's wire schema (untouched)
from llmq.workers.base import BaseWorker  # the plugin slot

from .service import GenerationService, build_service


class B200Worker(BaseWorker):
    """Native B200 generation worker (drop-in for VLLMWorker: same ctor, same behaviour)."""

    def __init__(
        self,
        model_name: str,
        queue_name: str,
        worker_id: Optional[str] = None,
        tensor_parallel_size: Optional[int] = None,
        data_parallel_size: Optional[int] = None,
        concurrency: Optional[int] = None,
        pipeline_name: Optional[str] = None,
        stage_name: Optional[str] = None,
        pipeline_stages: Optional[list] = None,
    ):
        # fields used by _generate_worker_id must exist before the base ctor runs (base.py:28)
        self.model_name = model_name
        self.tensor_parallel_size = tensor_parallel_size
        self.data_parallel_size = data_parallel_size
        super().__init__(queue_name, worker_id, concurrency, pipeline_name, stage_name, pipeline_stages)
        self.service: Optional[GenerationService] = None

    def _generate_worker_id(self) -> str:
        cuda_visible = os.environ.get("CUDA_VISIBLE_DEVICES", "0")
        gpu_suffix = cuda_visible.replace(",", "-") if cuda_visible else "auto"
        return f"b200-{gpu_suffix}"

    async def _initialize_processor(self) -> None:
        tp, dp = self.tensor_parallel_size or 1, self.data_parallel_size or 1
        if tp != 1 or dp != 1:
            # the path shards by running one worker per GPU against the same queue
            raise ValueError(
                f"b200 worker runs one replica per GPU (got -tp {tp} -dp {dp}); start one worker "
                "per GPU with CUDA_VISIBLE_DEVICES=<i> instead")
        self.logger.info(f"Initializing b200q engine for model {self.model_name}")
        cfg = self.config
        self.service = build_service(
            self.model_name, max_num_seqs=cfg.vllm_max_num_seqs, max_model_len=cfg.vllm_max_model_len,
            gpu_memory_utilization=cfg.vllm_gpu_memory_utilization, logger=self.logger)
        self.service.on_fatal = self._engine_died
        self.service.start()
        self.logger.info("b200q engine initialized successfully")

    def _engine_died(self, exc: BaseException) -> None:
        """engine thread: the native engine is gone.  Stop consuming (BaseWorker.run leaves its idle
        loop when `running` is False and runs cleanup) so the un-acked jobs go back to the queue for
        the other workers instead of being rejected and redelivered to this one in a hot loop."""
        self.logger.error(f"b200q engine thread died: {exc!r}; stopping worker {self.worker_id}")
        self.running = False

    def build_prompt(self, job: Job) -> str:
        """prompt text exactly as the reference builds it (vllm_worker.py:168-180)"""
        if job.chat_mode or job.messages:
            if not job.messages:
                raise ValueError("Chat mode enabled but no messages provided")
            return self.service.tokenizer.apply_chat_template(
                conversation=job.messages, tokenize=False, add_generation_prompt=True)
        return job.get_formatted_prompt()

    def stop_strings(self, job: Job) -> Optional[List[str]]:
        """vllm_worker.py:149-158: the job's stop list replaces the default; the default is the
        EOS token *string*, which can never appear in text decoded with skip_special_tokens, so
        it is inert and the engine's EOS-id stop is what ends generation."""
        if job.stop is not None:
            specials = set(getattr(self.service.tokenizer, "all_special_tokens", []) or [])
            return [s for s in job.stop if s and s not in specials] or None
        return None

    async def _process_job(self, job: Job) -> str:
        if self.service is None:
            raise RuntimeError("b200q engine not initialized")
        prompt = str(self.build_prompt(job))
        # text prompts are tokenised with add_special_tokens=True, as vLLM's renderer does
        # (vllm/renderers/base.py:330-342) — chat prompts therefore carry a double BOS (App. D1)
        ids = await self.service.encode_async(prompt)
        extra = job.model_dump()
        max_new = int(extra.get("max_tokens") or self.config.vllm_max_tokens)
        # the reference hard-codes temperature=0.7, unseeded (vllm_worker.py:161-165); a job may
        # carry `temperature` / `seed` extras, B200Q_TEMPERATURE overrides the default (0 = greedy)
        temperature = extra.get("temperature")
        if temperature is None:
            temperature = float(os.environ.get("B200Q_TEMPERATURE", "0.7"))
        seed = extra.get("seed")
        fut = self.service.submit(ids, max_new, self.stop_strings(job), asyncio.get_running_loop(),
                                  temperature=float(temperature), seed=None if seed is None else int(seed))
        text, _n = await fut
        return text

    async def _cleanup_processor(self) -> None:
        """ref:llmq/workers/vllm_worker.py:197-201 drops the engine; here that means: stop the engine
        thread, destroy the native handles and return weights / KV pool / workspace to the driver
        (NativeModel.close empties torch's cache), so another engine can be built in this process"""
        if self.service is not None:
            svc, self.service = self.service, None
            svc.stop()
            svc.engine.close()
            svc.engine.model.close()
