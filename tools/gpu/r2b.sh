#!/bin/bash
# GPU session r2b: test suite at the new head, bench variants (async on/off, budget/admission), small-batch points
O=gpurun_out/r2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3"
run() { name=$1; shift; s=$(date +%s); timeout 400 env "$@" $B $EXTRA > $O/$name.json 2> $O/$name.err; echo "$name rc=$? wall=$(( $(date +%s)-s ))s $(python -c "
import json,sys
try:
  d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print('value',d['value'],'jobs/s',d['jobs_per_sec'],'e2e',d['e2e']['value'],'gemm',d['roofline']['achieved'],'dec',d['roofline_decode_attn']['achieved'],'steps/step',d['engine']['engine_steps_per_bench_step'],'preempt',d['engine']['preemptions'])
except Exception as e: print('parse failed',e)
")"; }
EXTRA="" run base_async X=1
EXTRA="" run base_sync B200Q_ASYNC=0
EXTRA="--max-num-batched-tokens 9472" run b9472_admit1 B200Q_ADMIT_BATCH=1
EXTRA="--max-num-batched-tokens 9472" run b9472_hyst X=1
EXTRA="--max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048" run s128_async X=1
EXTRA="--max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048" run s128_sync B200Q_ASYNC=0
EXTRA="--max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048" run s128_admit1 B200Q_ADMIT_BATCH=1
EXTRA="--max-num-seqs 750 --jobs 375 --max-num-batched-tokens 4096" run s750_async X=1
EXTRA="--max-num-seqs 750 --jobs 375 --max-num-batched-tokens 4096" run s750_admit1 B200Q_ADMIT_BATCH=1
