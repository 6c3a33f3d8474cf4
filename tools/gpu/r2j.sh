#!/bin/bash
# GPU session r2j (4 GPUs): BASELINE configs #3, #4, #5 at process level through the reference CLI / broker path
O=gpurun_out/r2j; mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.used --format=csv,noheader | head -8
run() { name=$1; shift; s=$(date +%s); timeout 900 python tools/queue_bench.py --out-dir $O/$name "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$? wall=$(( $(date +%s)-s ))s"; tail -1 $O/$name.json | cut -c1-900; tail -3 $O/$name.err; }
run cfg3_llama8b_4workers --workers 4 --jobs 40000 --model random:llama-3-8b --max-num-seqs 4608 --budget 4608 --prefetch 6000
run cfg4_pipeline_2p2_gemma9b --pipeline 2+2 --jobs 12000 --model random:gemma-2-9b --max-num-seqs 1536 --budget 4608 --prefetch 2500
run cfg5_mixed_gemma9b_4workers --workers 4 --mixed --jobs 3000 --model random:gemma-2-9b --max-num-seqs 256 --budget 4608 --prefetch 400
for d in $O/*/; do for f in $d/qbench_worker*.log; do echo "== $f"; tail -2 $f | cut -c1-300; done; done 2>/dev/null | tail -40
