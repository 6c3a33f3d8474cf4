#!/bin/bash
# GPU session r2c: full GPU suite, Gemma-2-9B bench (D=256 transposed-PV decode attention), vLLM worker goldens
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
s=$(date +%s); timeout 500 python bench.py --model gemma-2-9b --gpus 1 --steps 4 --warmup 3 --max-num-seqs 1536 --jobs 384 --max-num-batched-tokens 4608 --e2e-steps 2 > $O/gemma2_9b.json 2> $O/gemma2_9b.err; echo "gemma rc=$? wall=$(( $(date +%s)-s ))s"; tail -2 $O/gemma2_9b.err; head -c 2500 $O/gemma2_9b.json
s=$(date +%s); timeout 1500 python tools/oracle_vllm_worker.py golden llama32_1b llama3_8b_w4 > $O/vllm_golden.log 2>&1; echo "golden rc=$? wall=$(( $(date +%s)-s ))s"; grep -E "^\[|FAILED|Error" $O/vllm_golden.log | tail -12
ls -la gpurun_out/vllm_worker_golden_* 2>/dev/null
