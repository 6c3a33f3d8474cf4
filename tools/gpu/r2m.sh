#!/bin/bash
# GPU session r2m: admission-batch / token-budget sweep at 128 and 750 sequences (engine-direct, same protocol as the same-lease table)
O=gpurun_out/r2m; mkdir -p $O
one() { tag=$1; seqs=$2; shift 2; r=$(env "$@" python tools/vllm_incumbent.py one native $seqs 4 2>&1 | grep RESULT | sed 's/RESULT //'); echo "$tag seqs=$seqs $* -> $(echo $r | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["out_tokens_per_s"], d["engine_steps"])')"; echo "{\"tag\":\"$tag\",\"env\":\"$*\",\"result\":$r}" >> $O/sweep.jsonl; }
one base 128 X=1
one admit4 128 B200Q_ADMIT_BATCH=4
one admit16 128 B200Q_ADMIT_BATCH=16
one admit32 128 B200Q_ADMIT_BATCH=32
one admit64 128 B200Q_ADMIT_BATCH=64
one nographs 128 B200Q_CUDA_GRAPHS=0
one base 750 X=1
one admit16 750 B200Q_ADMIT_BATCH=16
one admit94 750 B200Q_ADMIT_BATCH=94
one admit188 750 B200Q_ADMIT_BATCH=188
