#!/bin/bash
# GPU session r2k: re-run of the fixed tests; Llama-3.2-1B bench (BASELINE config #2 shape: host path matters most there)
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_worker_gpu.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log; grep -E "^E  " $O/pytest_gpu.log | head -12
s=$(date +%s); timeout 600 python bench.py --model llama-3.2-1b --gpus 1 --steps 8 --warmup 5 --e2e-steps 8 > $O/bench_1b.json 2> $O/bench_1b.err; echo "1b rc=$? wall=$(( $(date +%s)-s ))s"; python -c "
import json
d=json.loads(open('$O/bench_1b.json').read().strip().splitlines()[-1]); print('value',d['value'],'jobs/s',d['jobs_per_sec'],'e2e',d['e2e']['value'],'e2e jobs/s',d['e2e']['jobs_per_sec'],'gemm',d['roofline']['achieved'],d['roofline']['share_of_device_time'],'dec',d['roofline_decode_attn']['achieved'])"
