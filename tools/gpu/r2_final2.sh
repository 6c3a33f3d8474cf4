#!/bin/bash
# Re-validation after the host-side teardown fix: GPU suite, smoke, bench at the driver's settings
O=gpurun_out/r2_final2; mkdir -p $O
s=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s)-s ))s"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
s=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s)-s ))s"; echo "stderr lines: $(grep -c . $O/bench_n1.err)"; grep -i "error\|Traceback" $O/bench_n1.err | head -3
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]); print('value',d['value'],'jobs/s',d['jobs_per_sec'],'e2e',d['e2e']['value'],'gemm',d['roofline']['frac'],'dec',d['roofline_decode_attn']['frac'])"
