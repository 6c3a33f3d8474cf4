#!/bin/bash
# GPU session r2d: full GPU suite, ncu evidence for the dispatched kernels, vLLM vs native on one lease
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log; grep -E "agreed prefix" $O/pytest_gpu.log
# launch list of the real bench (steady state): skip the ramp, list 3000 launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 -c 3000 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 2 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err; echo "ncu list rc=$?"; wc -l $O/launches_bench.csv
# full sets for the dominant kernels at the bench shapes (isolated process)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm|decode_attn' -c 12 -o $O/ncu_full_targets -f \
   python tools/gpu_probe.py ncu_targets > $O/ncu_targets.log 2>&1; echo "ncu full rc=$?"; ls -la $O/*.ncu-rep
ncu -i $O/ncu_full_targets.ncu-rep --page raw --csv > $O/ncu_full_targets.csv 2>/dev/null; wc -l $O/ncu_full_targets.csv
s=$(date +%s); timeout 1500 python tools/vllm_incumbent.py same_lease > $O/same_lease.log 2>&1; echo "same_lease rc=$? wall=$(( $(date +%s)-s ))s"; grep -E '^\{' $O/same_lease.log
