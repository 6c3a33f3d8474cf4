#!/bin/bash
# GPU session r2f: where does a 128-sequence decode step go?  per-kernel durations inside the engine (warm caches)
O=gpurun_out/r2f; mkdir -p $O
B="python bench.py --gpus 1 --steps 2 --warmup 5 --no-cpu-baseline --e2e-steps 1 --max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__grid_size --cache-control none --clock-control none --launch-skip 150000 -c 2000 --csv --log-file $O/launches_s128.csv $B > $O/bench.json 2> $O/bench.err; echo "ncu rc=$?"; wc -l $O/launches_s128.csv
