"""Debug: can a kernel of libgabo_b200.so running on this rank's GPU store into a peer rank's buffer (CUDA IPC + peer access)?"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from gaboflow_b200 import _lib
from gaboflow_b200.dist import PeerBuffers
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
buf = torch.zeros((1024, 1024), dtype=torch.float64, device=dev)
peers = PeerBuffers(buf)
print(rank, 'ptrs', [hex(p) for p in peers.ptrs], 'devices', [str(t.device) for t in peers.tensors], flush=True)
L = _lib.lib()
other = (rank + 1) % world
# 1) torch cross-device write
peers.tensors[other][0, :4] = float(rank + 1)
torch.cuda.synchronize(); dist.barrier()
print(rank, 'after torch write, my row0:', buf[0, :4].tolist(), flush=True)
# 2) our own kernel, launched on MY device, writing 16 doubles at the peer pointer + 8 KB
st = L.gabo_kdiag_const_f64(16, float(10 + rank), peers.ptrs[other] + 8192, torch.cuda.current_stream().cuda_stream)
print(rank, 'launch status', st, flush=True)
torch.cuda.synchronize(); dist.barrier()
print(rank, 'after kernel write, my row1:', buf[1, :4].tolist(), flush=True)
dist.barrier(); dist.destroy_process_group()
