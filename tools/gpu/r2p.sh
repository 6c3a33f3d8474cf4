#!/bin/bash
# GPU session r2p: admission hysteresis at the headline configuration (decode M = 4608 - free slots: tile padding)
O=gpurun_out/r2p; mkdir -p $O
B="python bench.py --gpus 1 --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for ab in 288 144 72 36; do
  B200Q_ADMIT_BATCH=$ab timeout 300 $B > $O/admit$ab.json 2> $O/admit$ab.err; echo "admit=$ab rc=$? $(python -c "
import json
d=json.loads(open('$O/admit$ab.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'gemm',d['roofline']['achieved'],'dec',d['roofline_decode_attn']['achieved'],'steps/step',d['engine']['engine_steps_per_bench_step'])")"
done
