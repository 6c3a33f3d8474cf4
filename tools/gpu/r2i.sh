#!/bin/bash
# GPU session r2i: fused split-K consumers (tests + small-batch numbers), GPU pipeline test
O=gpurun_out/r2i; mkdir -p $O
timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_worker_gpu.py tests/test_fullsize_gpu.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log; grep -E "^E  " $O/pytest_gpu.log | head -20
python tools/vllm_incumbent.py one native 128 4 2>&1 | grep RESULT
B200Q_FUSE_SPLITK=0 python tools/vllm_incumbent.py one native 128 4 2>&1 | grep RESULT
B="python bench.py --gpus 1 --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 3 --max-num-seqs 128 --jobs 128 --max-num-batched-tokens 2048"
for f in 1 0; do
  B200Q_FUSE_SPLITK=$f timeout 300 $B > $O/s128_fuse$f.json 2> $O/s128_fuse$f.err; echo "s128 fuse=$f rc=$? $(python -c "
import json
d=json.loads(open('$O/s128_fuse$f.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'])")"
done
