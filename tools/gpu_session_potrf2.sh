#!/bin/bash
# N-GPU session: chain vs look-ahead distributed Cholesky -- bitwise check against the collective version, then bench lines
W=${1:-2}
export CUDA_DEVICE_MAX_CONNECTIONS=${CUDA_DEVICE_MAX_CONNECTIONS:-32}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29551"
{
for sched in chain lookahead; do
  echo "== dist_check GABO_DIST_SCHED=$sched"
  for n in ${CHECK_N:-8192 16384}; do GABO_DIST_SCHED=$sched timeout 300 $TR tools/dist_check.py $n 2>&1 | grep "rank"; done
done
} > gpurun_out/r02_potrf_chain_${W}gpu.log 2>&1
cat gpurun_out/r02_potrf_chain_${W}gpu.log | cut -c1-250
for sched in chain lookahead; do
  GABO_DIST_SCHED=$sched timeout 600 $TR bench.py --gpus $W --steps 5 --warmup 3 --no-spd --no-cpu-baseline > gpurun_out/r02_bench_${W}gpu_$sched.json 2> gpurun_out/r02_bench_${W}gpu_$sched.err
  tail -c 400 gpurun_out/r02_bench_${W}gpu_$sched.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_${W}gpu_$sched.json').read().strip().splitlines()[-1])
print('$sched', d['value'], d['ms_per_step'], d['phases_ms'], d['checks'], d['loglik'])
P
done
