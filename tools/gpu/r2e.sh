#!/bin/bash
# GPU session r2e: small-M GEMM with the L2 weight-prefetch window; GEMM/vLLM-golden tests; s128 bench
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_vllm_worker_golden.py tests/test_model_gpu.py -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log; grep -E "agreed prefix" $O/pytest_gpu.log
timeout 600 python tools/gpu_probe.py gemm_small_m_pf > $O/probe.log 2>&1; echo "probe rc=$?"; grep -E '^\{' $O/probe.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'],d['M'],{k:v for k,v in d.items() if k.startswith('us_') or k.startswith('ideal')})
"
B="python bench.py --gpus 1 --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 --max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048"
for pf in 0 -1; do
  s=$(date +%s); B200Q_GEMM_PREFETCH=$pf timeout 300 $B > $O/s128_pf$pf.json 2> $O/s128_pf$pf.err; echo "s128 pf=$pf rc=$? $(python -c "
import json
d=json.loads(open('$O/s128_pf$pf.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'])")"
done
