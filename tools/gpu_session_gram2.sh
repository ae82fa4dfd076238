#!/bin/bash
# 2+ GPU session: staged vs direct block-cyclic Gram (correctness + timing sweep), then the bench line.
W=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531"
mkdir -p gpurun_out
{
echo "== correctness N=8192 staged"; GABO_GRAM_MODE=staged GABO_SWEEP=0,64,256 timeout 300 $TR tools/dist_gram_check.py 8192 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -$((3*W))
echo "== N=65536 direct"; GABO_GRAM_MODE=direct GABO_SWEEP=${DIRECT_SWEEP:-80} timeout 300 $TR tools/dist_gram_check.py 65536 2>&1 | grep "rank 0"
echo "== N=65536 staged split=1"; GABO_GRAM_MODE=staged GABO_SWEEP=${STAGED_SWEEP:-0,16,32,64} timeout 400 $TR tools/dist_gram_check.py 65536 2>&1 | grep "rank 0"
echo "== N=65536 staged split=0"; GABO_GRAM_MODE=staged GABO_GRAM_SPLIT=0 GABO_SWEEP=${STAGED_SWEEP0:-0} timeout 300 $TR tools/dist_gram_check.py 65536 2>&1 | grep "rank 0"
} > gpurun_out/r02_gram_staged_${W}gpu.log 2>&1
cat gpurun_out/r02_gram_staged_${W}gpu.log
timeout 600 $TR bench.py --gpus $W --steps 5 --warmup 3 > gpurun_out/r02_bench_${W}gpu_staged.json 2> gpurun_out/r02_bench_${W}gpu_staged.err
tail -c 600 gpurun_out/r02_bench_${W}gpu_staged.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_${W}gpu_staged.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phases_ms'], d['checks'])
P
