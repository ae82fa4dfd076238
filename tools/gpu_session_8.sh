#!/bin/bash
# 8-GPU session: three-way Gram split sweep, chain-vs-lookahead Cholesky (bitwise check + bench lines)
W=${1:-8}
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29561"
{
echo "== staged three-way (f:d)"; GABO_GRAM_MODE=staged GABO_SWEEP=${SWEEP3:-80:80,96:64,96:96,112:64,128:64} timeout 400 $TR tools/dist_gram_check.py 65536 2>&1 | grep -o "rank 0/.*"
echo "== dist_check chain"; GABO_DIST_SCHED=chain timeout 300 $TR tools/dist_check.py 16384 2>&1 | grep -o "rank [0-9]/.*" | cut -c1-200
} > gpurun_out/r02_session8_${W}gpu.log 2>&1
cat gpurun_out/r02_session8_${W}gpu.log
for sched in chain lookahead; do
  GABO_GRAM_MODE=direct GABO_DIST_SCHED=$sched timeout 600 $TR bench.py --gpus $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${W}gpu_$sched.json 2> gpurun_out/r02_bench_${W}gpu_$sched.err
  tail -c 300 gpurun_out/r02_bench_${W}gpu_$sched.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_${W}gpu_$sched.json').read().strip().splitlines()[-1])
print('$sched', d['value'], d['ms_per_step'], d['phases_ms'], d['checks'], d['loglik'], {k:(v.get('ms') if isinstance(v,dict) else None) for k,v in d.get('spd',{}).items()})
P
done
