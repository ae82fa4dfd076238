"""Repeat the emulated-ranks staged Gram (tests/test_gpu_dist.py::test_cyclic_staged_gram_single_process) many times and
report where mismatches against the single-GPU mirrored Gram occur: python tools/staged_race_probe.py world frac split direct reps"""
import sys
import torch
sys.path.insert(0, '.')
from gaboflow_b200 import ops
from gaboflow_b200.dist import BlockCyclic
from oracle import gabo_oracle as orc
world, frac, split, direct, reps = [int(a) for a in sys.argv[1:6]]
n = 512 * 11
X = torch.from_numpy(orc.synth_sphere(n, 51, seed=8)).cuda()
lays = [BlockCyclic(n, 512, world, r) for r in range(world)]
bufs = [torch.full((max(l.local_rows, 1), n), float('nan'), dtype=torch.float64, device='cuda') for l in lays]
ptrs = [b.data_ptr() for b in bufs]
states = [ops.StagedGramState(n, world, r, frac, split, 'cuda', direct_frac256=direct) for r in range(world)]
full = ops.points_gram(ops.SPHERE_GAUSSIAN, X, None, beta=2.0, uplo='mirror', engine='dmma')
bad = 0
for rep in range(reps):
    for b in bufs:
        b.fill_(float('nan'))
    for r in range(world):
        ops.points_gram_cyclic_staged(ops.SPHERE_GAUSSIAN, X, ptrs, n, world, r, states[r], beta=2.0, variance=1.0)
    torch.cuda.synchronize()
    for r in range(world):
        if lays[r].local_rows == 0:
            continue
        rows = lays[r].global_rows().cuda()
        got, ref = bufs[r][:lays[r].local_rows], full[rows]
        ne = ~((got == ref) | (got.isnan() & ref.isnan()))
        if bool(ne.any()):
            bad += 1
            idx = ne.nonzero()
            nan_cnt = int(got[ne].isnan().sum())
            rr, cc = idx[:, 0], idx[:, 1]
            print(f'rep {rep} rank {r}: {int(ne.sum())} mismatches ({nan_cnt} still NaN); local rows {int(rr.min())}..{int(rr.max())} '
                  f'(global {int(rows[rr.min()])}..{int(rows[rr.max()])}), cols {int(cc.min())}..{int(cc.max())}', flush=True)
print(f'{bad} failing (rep, rank) pairs in {reps} repetitions')
