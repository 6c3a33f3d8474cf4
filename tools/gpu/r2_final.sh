#!/bin/bash
# Final validation at the frozen head: GPU suite, smoke(), bench at the driver's settings (+ reference arm)
O=gpurun_out/r2_final; mkdir -p $O
s=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s)-s ))s"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log | cut -c1-200
s=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$? wall=$(( $(date +%s)-s ))s"
s=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s)-s ))s"; tail -2 $O/bench_n1.err
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]); print('value',d['value'],'jobs/s',d['jobs_per_sec'],'e2e',d['e2e']['value'],'gemm',d['roofline']['achieved'],d['roofline']['frac'],'dec',d['roofline_decode_attn']['achieved'],d['roofline_decode_attn']['frac'],d['roofline']['share_of_device_time'],d['clocks'])"
s=$(date +%s); timeout 400 python bench.py --gpus 1 > $O/bench_default.json 2> $O/bench_default.err; echo "default-flags bench rc=$? wall=$(( $(date +%s)-s ))s"
