#!/bin/bash
# staged Gram at N=65536: effect of CUDA_DEVICE_MAX_CONNECTIONS (hardware queues per context) on the copy-engine streams
W=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29541"
{
for mc in 32 8; do
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$mc staged"
  CUDA_DEVICE_MAX_CONNECTIONS=$mc GABO_GRAM_MODE=staged GABO_SWEEP=${STAGED_SWEEP:-48,80} timeout 300 $TR tools/dist_gram_check.py 65536 2>&1 | grep "rank 0"
done
echo "== CUDA_DEVICE_MAX_CONNECTIONS=32 direct"
CUDA_DEVICE_MAX_CONNECTIONS=32 GABO_GRAM_MODE=direct GABO_SWEEP=${DIRECT_SWEEP:-160} timeout 300 $TR tools/dist_gram_check.py 65536 2>&1 | grep "rank 0"
} > gpurun_out/r02_gram_staged_${W}gpu_maxconn.log 2>&1
cat gpurun_out/r02_gram_staged_${W}gpu_maxconn.log
