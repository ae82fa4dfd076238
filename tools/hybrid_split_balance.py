"""Balance of the hybrid Gram's hashed mirror / recompute split (Python mirror of cyclic_recompute in csrc/points_gram.cu):
share of the cross-GPU panel pairs that is recomputed, and per-rank imbalance of the tile work and of the NVLink blocks sent,
for N = 65536 (128 panels of 512 rows) and the default share of every world size.  CPU only.
    python tools/hybrid_split_balance.py"""
import sys

import numpy as np

sys.path.insert(0, '.')
from gaboflow_b200.dist import default_recompute_frac256  # noqa: E402


def recompute(PI, PJ, world, frac):
    if frac <= 0 or PI % world == PJ % world:
        return False
    h = ((PI * 0x9E3779B1) & 0xffffffff) ^ ((PJ * 0x85EBCA77) & 0xffffffff)
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & 0xffffffff
    h ^= h >> 12
    return (h & 255) < frac


npan = 128
for W in range(2, 9):
    frac = default_recompute_frac256(W)
    rec, sent, lower, cross = np.zeros(W, int), np.zeros(W, int), np.zeros(W, int), 0
    for PJ in range(npan):
        for PI in range(PJ + 1, npan):
            lower[PI % W] += 1
            if PI % W != PJ % W:
                cross += 1
                if recompute(PI, PJ, W, frac):
                    rec[PJ % W] += 1
                else:
                    sent[PI % W] += 1
    work = lower + rec
    print(f'W={W} share {frac}/256: recomputed {rec.sum() / cross:.3f} of the cross-GPU pairs; tile work max/mean '
          f'{work.max() / work.mean():.3f}; NVLink blocks sent max/mean {sent.max() / sent.mean():.3f}')
