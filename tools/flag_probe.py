"""Where does a stream blocked by gabo_flag_wait_geq stall other streams?  (debugging aid)"""
import faulthandler
import sys
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops
faulthandler.dump_traceback_later(25, exit=True)
flags = torch.zeros(8, dtype=torch.int32, device='cuda')
out = torch.zeros(1, dtype=torch.float64, device='cuda')
out.add_(1.0).zero_(); ops.flag_signal([flags.data_ptr() + 28], 1); float(out.item()); flags.cpu()
torch.cuda.synchronize()
waiter, signaller = torch.cuda.Stream(), torch.cuda.Stream()
print('streams', waiter.cuda_stream, signaller.cuda_stream, flush=True)
with torch.cuda.stream(waiter):
    ops.flag_wait_geq(flags.data_ptr() + 4, 7)
    out.add_(1.0)
print('wait queued', flush=True)
with torch.cuda.stream(signaller):
    ops.flag_signal([flags.data_ptr() + 4, flags.data_ptr() + 8], 6)
print('signal 6 queued', flush=True)
signaller.synchronize()
print('signaller synced; waiter done?', waiter.query(), flush=True)
print('out', float(out.item()), flush=True)
with torch.cuda.stream(signaller):
    ops.flag_signal([flags.data_ptr() + 4], 9)
waiter.synchronize()
print('released', float(out.item()), flags.cpu().tolist(), flush=True)
