#!/bin/bash
# GPU session r2g: in-kernel split-K fix-up (no reduce kernel) + L2 promotion of small-M weight maps
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_gemma2_gpu.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python tools/gpu_probe.py gemm_small_m_pf > $O/probe.log 2>&1; echo "probe rc=$?"; grep -E '^\{' $O/probe.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'],d['M'],{k:v for k,v in d.items() if k.startswith('us_') or k.startswith('ideal')})
"
B="python bench.py --gpus 1 --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 --max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048"
for pr in 256 128 0; do
  B200Q_GEMM_SMALLM_PROMO=$pr timeout 300 $B > $O/s128_promo$pr.json 2> $O/s128_promo$pr.err; echo "s128 promo=$pr rc=$? $(python -c "
import json
d=json.loads(open('$O/s128_promo$pr.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'])")"
done
