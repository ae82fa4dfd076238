"""2+ GPUs under torchrun: the NVLink-mirrored block-cyclic Gram against each rank's own full-row computation.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_gram_check.py [N]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402
from gaboflow_b200.dist import BlockCyclic, PeerBuffers, gram_block_cyclic_mirrored  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
rng = np.random.default_rng(20240)
X = rng.standard_normal((N, 51)); X /= np.linalg.norm(X, axis=1, keepdims=True)
Xd = torch.from_numpy(X).to(dev)
lay = BlockCyclic(N, 512, world, rank)
peers = PeerBuffers(lay.local_rows, N, dev)
K = peers.local
ref = torch.empty_like(K)
for p in lay.my_panels():
    s = lay.slot(p) * 512
    ops.points_gram(ops.SPHERE_GAUSSIAN, Xd, None, beta=2.0, out=ref[s:s + 512], row0=p * 512, row1=(p + 1) * 512, uplo='full')
ws = torch.empty(int(ops._lib.lib().gabo_points_gram_cyclic_workspace_bytes(N, 51)), dtype=torch.uint8, device=dev)
ok = True
for mod in [int(v) for v in os.environ.get("GABO_SWEEP", "0,64,96,128,160,256").split(",")]:   # fraction (of 256) of the cross-rank upper panel pairs recomputed by the row owner
    K.fill_(float('nan'))
    dist.barrier(); torch.cuda.synchronize()
    gram_block_cyclic_mirrored(ops.SPHERE_GAUSSIAN, Xd, lay, peers, 2.0, 1.0, workspace=ws, recompute_frac256=mod)
    err = float((K - ref).abs().max().item())
    nan = int(torch.isnan(K).sum().item())
    for _ in range(2):
        gram_block_cyclic_mirrored(ops.SPHERE_GAUSSIAN, Xd, lay, peers, 2.0, 1.0, workspace=ws, recompute_frac256=mod)
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gram_block_cyclic_mirrored(ops.SPHERE_GAUSSIAN, Xd, lay, peers, 2.0, 1.0, workspace=ws, sync=False, recompute_frac256=mod)
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    ms = e0.elapsed_time(e1) / 3
    good = nan == 0 and err == 0.0
    ok = ok and good
    print(f'[rank {rank}/{world}] N={N} recompute_frac256={mod}: max|K-ref|={err:.1e} nan={nan}  {ms:.2f} ms  {"OK" if good else "FAIL"}', flush=True)
dist.barrier()
del K, ref
peers.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
sys.exit(0 if ok else 1)
