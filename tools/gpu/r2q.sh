#!/bin/bash
# GPU session r2q: batch tokenisation on the worker path — worker / golden tests, 1B and 8B e2e
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_worker_gpu.py tests/test_vllm_worker_golden.py tests/test_gemma2_gpu.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 400 python bench.py --model llama-3.2-1b --gpus 1 --steps 8 --warmup 5 --e2e-steps 8 --no-cpu-baseline > $O/bench_1b.json 2> $O/bench_1b.err; echo "1b rc=$?"; python -c "
import json
d=json.loads(open('$O/bench_1b.json').read().strip().splitlines()[-1]); print('1B value',d['value'],'e2e',d['e2e']['value'],'gap %.1f%%' % (100*(1-d['e2e']['value']/d['value'])))"
timeout 400 python bench.py --gpus 1 --steps 8 --warmup 5 --e2e-steps 8 --no-cpu-baseline > $O/bench_8b.json 2> $O/bench_8b.err; echo "8b rc=$?"; python -c "
import json
d=json.loads(open('$O/bench_8b.json').read().strip().splitlines()[-1]); print('8B value',d['value'],'e2e',d['e2e']['value'])"; grep -ci "error\|traceback" $O/bench_8b.err $O/bench_1b.err
