"""2+ GPUs under torchrun: the NVLink-mirrored block-cyclic Gram against each rank's own full-row computation.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_gram_check.py [N]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from gaboflow_b200 import ops  # noqa: E402
from gaboflow_b200.dist import BlockCyclic, PeerBuffers, PanelExchange, gram_block_cyclic_mirrored, make_staged_gram_state  # noqa: E402

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
MODE = os.environ.get('GABO_GRAM_MODE', 'direct')      # direct: remote stores from the kernel; staged: local staging + copy engines
SPLIT = int(os.environ.get('GABO_GRAM_SPLIT', '1'))
DIRECT = int(os.environ.get('GABO_GRAM_DIRECT256', '0'))     # staged mode: share of the slots stored remotely by the kernel
xch = PanelExchange(lay, N, dev) if os.environ.get('GABO_GRAM_XCH', '1') == '1' and world > 1 else None
ok = True
for item in os.environ.get("GABO_SWEEP", "0,64,96,128,160,256").split(","):   # "f" or "f:d" (staged mode: d/256 of the slots by remote stores)
    mod = int(item.split(':')[0])
    if ':' in item:
        DIRECT = int(item.split(':')[1])
    # mod =   # fraction (of 256) of the cross-rank upper panel pairs recomputed by the row owner
    staged = make_staged_gram_state(N, lay, dev, recompute_frac256=mod, split=SPLIT, direct_frac256=DIRECT) if MODE == 'staged' else None
    K.fill_(float('nan'))
    dist.barrier(); torch.cuda.synchronize()
    gram_block_cyclic_mirrored(ops.SPHERE_GAUSSIAN, Xd, lay, peers, 2.0, 1.0, workspace=ws, recompute_frac256=mod, xch=xch, staged=staged)
    torch.cuda.synchronize(); dist.barrier()
    err = float((K - ref).abs().max().item())
    nan = int(torch.isnan(K).sum().item())
    for _ in range(2):
        gram_block_cyclic_mirrored(ops.SPHERE_GAUSSIAN, Xd, lay, peers, 2.0, 1.0, workspace=ws, recompute_frac256=mod, xch=xch, staged=staged)
        torch.cuda.synchronize(); dist.barrier()
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    REPS = 3
    for _ in range(REPS):   # with the flag handshake every Gram ends when all peers' blocks have landed (what bench.py times)
        gram_block_cyclic_mirrored(ops.SPHERE_GAUSSIAN, Xd, lay, peers, 2.0, 1.0, workspace=ws, sync=xch is not None, recompute_frac256=mod, xch=xch, staged=staged)
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    ms = e0.elapsed_time(e1) / REPS
    good = nan == 0 and err == 0.0
    ok = ok and good
    del staged
    print(f'[rank {rank}/{world}] N={N} mode={MODE} split={SPLIT} direct256={DIRECT} recompute_frac256={mod}: max|K-ref|={err:.1e} nan={nan}  {ms:.2f} ms  {"OK" if good else "FAIL"}', flush=True)
dist.barrier()
del K, ref
if xch is not None:
    xch.close()
peers.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
