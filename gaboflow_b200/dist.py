"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing (NCCL over NVLink on the GPU box, gloo in the
CPU tests of the host logic), the C-ABI kernels for all local arithmetic.

Layout: the Gram matrix is distributed by ROW PANELS of nb rows, 1-D block-cyclic: global panel p
lives on rank p % world in local slot p // world, each rank storing full rows [local_rows, n] (row-major, so a "column
panel" of the symmetric matrix is a row panel here).
  * Gram (gram_block_cyclic_mirrored): X is replicated (<= 27 MB); every rank computes the lower tiles of its own panels;
    an upper block is either stored into its owner's HBM over NVLink from inside the kernel or recomputed by that owner
    (hashed split, default_recompute_frac256); completion is a stream-ordered flag handshake.  gram_block_cyclic is the
    exchange-free variant (every rank produces whole rows).
  * Cholesky, right-looking over panels, two drivers with identical arithmetic (bit-identical factors):
      potrf_block_cyclic_p2p_  peer-memory exchange (PanelExchange) + one panel of look-ahead: the diagonal block is pushed
                               by copy engines, the block column lands on every rank by NVLink stores from the row-solve
                               kernel, ordering by cuStreamWaitValue32 flags -- no collective inside the loop;
      potrf_block_cyclic_      owner factors the diagonal block and broadcasts it, every rank solves its rows, the block
                               column is all-gathered (NCCL / gloo); the version the CPU tests exercise.
    Every rank then updates the lower part of its own rows (gabo_syrk_cyclic_f64, mask in global coordinates).
  * Forward substitution is pipelined over panels with an nb-vector broadcast per step and reuses the block inverses the
    factorisation produced; log-likelihood terms are summed with one all-reduce.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class BlockCyclic:
    """1-D block-cyclic distribution of n rows in panels of nb over `world` ranks."""
    n: int
    nb: int
    world: int
    rank: int

    @property
    def npanels(self) -> int:
        return (self.n + self.nb - 1) // self.nb

    def owner(self, p: int) -> int:
        return p % self.world

    def slot(self, p: int) -> int:
        return p // self.world

    def panels_of(self, rank: int) -> List[int]:
        return list(range(rank, self.npanels, self.world))

    def my_panels(self) -> List[int]:
        return self.panels_of(self.rank)

    def panel_rows(self, p: int) -> int:
        return min(self.nb, self.n - p * self.nb)

    def local_rows_of(self, rank: int) -> int:
        return sum(self.panel_rows(p) for p in self.panels_of(rank))

    @property
    def local_rows(self) -> int:
        return self.local_rows_of(self.rank)

    def slots_after(self, kp: int, rank: Optional[int] = None) -> int:
        """First local slot of `rank` whose global panel index is > kp."""
        r = self.rank if rank is None else rank
        return 0 if kp < r else (kp - r) // self.world + 1

    def global_rows(self) -> torch.Tensor:
        """Global row index of every local row (CPU int64)."""
        idx = [torch.arange(p * self.nb, p * self.nb + self.panel_rows(p)) for p in self.my_panels()]
        return torch.cat(idx) if idx else torch.empty(0, dtype=torch.int64)


class CudaLocalOps:
    """Local arithmetic through the C-ABI (the only implementation in the product)."""

    @staticmethod
    def potrf_(A, sigma2):
        from . import ops
        return ops.potrf_(A, sigma2, check=False)

    @staticmethod
    def potrf_diag_(A, sigma2, winv_all):
        from . import ops
        return ops.potrf_diag_(A, sigma2, winv_all)

    @staticmethod
    def trsm_right_(A, L, winv_all=None):
        from . import ops
        return ops.trsm_right_(A, L, winv_all)

    @staticmethod
    def syrk_cyclic_(C, A, B, row_base, col_base, nb, world, rank):
        from . import ops
        return ops.syrk_cyclic_(C, A, B, row_base, col_base, nb, world, rank)

    @staticmethod
    def trsm_left_(L, B, winv_all=None):
        from . import ops
        return ops.trsm_(L, B, trans=False, winv_all=winv_all)

    @staticmethod
    def solve_update_(Y, Lp, X):
        from . import ops
        return ops.solve_update_(Y, Lp, X)

    @staticmethod
    def loglik_block(Lblk, Vblk):
        from . import ops
        return ops.gp_loglik(Lblk, Vblk)          # the Gaussian log-density is additive over row blocks


def _world(group=None):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def gram_block_cyclic(gram_rows, layout: BlockCyclic, out: Optional[torch.Tensor] = None, dtype=torch.float64,
                      device=None) -> torch.Tensor:
    """Fill this rank's row panels: gram_rows(row0, row1, out_view) writes K[row0:row1, :] into out_view."""
    if out is None:
        out = torch.empty((layout.local_rows, layout.n), dtype=dtype, device=device)
    for p in layout.my_panels():
        s = layout.slot(p)
        r0 = p * layout.nb
        r1 = r0 + layout.panel_rows(p)
        gram_rows(r0, r1, out[s * layout.nb: s * layout.nb + (r1 - r0)])
    return out


def potrf_block_cyclic_(A_loc: torch.Tensor, layout: BlockCyclic, sigma2: float, n_factor: Optional[int] = None,
                        local=CudaLocalOps, group=None, lookahead: bool = False, winv_store: Optional[dict] = None,
                        phase_ms: Optional[dict] = None) -> int:
    """In-place distributed Cholesky of the leading n_factor x n_factor block (default: everything) of the symmetric
    matrix whose lower triangle is stored block-cyclically in A_loc [local_rows, >= n_factor].  Returns LAPACK-style info
    (0, or the order of the first non-positive-definite leading minor), identical on every rank.  If ``winv_store`` (a dict)
    is given, the owner of each panel leaves the inverses of its diagonal 128-blocks there for solve_lower_block_cyclic_.
    If ``phase_ms`` (a dict) is given (CUDA, no look-ahead), it receives the summed device time of each piece of the loop
    (diag / bcast / solve / gather / u1 / u2), from CUDA events on the main stream: a diagnostic, not a bench number.

    Right-looking over panels; the trailing update is split into U1 (next block column) and U2 (the rest).  With
    ``lookahead=True`` the compute-only parts of the next panel (diagonal factorisation on the owner, row solve on every
    rank) run on a high-priority side stream under the two halves of U2, and the two collectives are issued on the main
    stream between / after those halves.  (A first version that also ran the collectives on the side stream was slower
    than no look-ahead: NCCL's CTAs need a whole SM's registers and starve behind the 2-CTA/SM update kernel.)  On CPU
    tensors (gloo tests) the same steps run in order."""
    W, rank, nb = layout.world, layout.rank, layout.nb
    n = layout.n if n_factor is None else n_factor
    if n % nb != 0 and n != layout.n:
        raise ValueError('n_factor must be a multiple of the panel size')
    npan = (n + nb - 1) // nb
    dev = A_loc.device
    dt = A_loc.dtype
    nwb = (nb + 127) // 128
    use_streams = bool(A_loc.is_cuda and lookahead and npan > 1)
    main = torch.cuda.current_stream(dev) if A_loc.is_cuda else None
    side = torch.cuda.Stream(device=dev, priority=-1) if use_streams else main
    my = [p for p in layout.my_panels() if p < npan]
    rows_f = sum(min(nb, n - p * nb) for p in my)
    maxslots = (npan + W - 1) // W
    # double-buffered communication buffers (panel k+1 is produced while panel k is still being consumed)
    bc = [torch.empty((nb * nb + nwb * 128 * 128,), dtype=dt, device=dev) for _ in range(2)]   # factor + block inverses
    send = [torch.zeros((maxslots * nb, nb), dtype=dt, device=dev) for _ in range(2)] if W > 1 else None
    recv = [torch.empty((W, maxslots * nb, nb), dtype=dt, device=dev) for _ in range(2)] if W > 1 else None
    first_bad = torch.full((1,), 2 ** 62, dtype=torch.int64, device=dev)

    class _Ctx:
        def __init__(self, stream):
            self.stream = stream

        def __enter__(self):
            if A_loc.is_cuda:
                self.cm = torch.cuda.stream(self.stream)
                self.cm.__enter__()

        def __exit__(self, *a):
            if A_loc.is_cuda:
                self.cm.__exit__(*a)

    # ---- the panel of block column kp in three pieces -------------------------------------------------------------
    #   diag(kp)    owner factors the diagonal block (+ inverses of its 128-blocks)        compute only
    #   bcast(kp)   NCCL broadcast of factor + inverses                                     collective
    #   solve(kp)   every rank solves its rows of the block column, packs the send buffer   compute only
    #   gather(kp)  NCCL all-gather + reorder into global row order                         collective
    # With look-ahead the compute-only pieces run on a high-priority side stream while the main stream updates the trailing
    # matrix; the collectives are issued on the MAIN stream between the two halves of that update: NCCL's CTAs need a
    # whole SM's registers and starve behind the 2-CTA/SM update kernel when issued concurrently (measured).
    def diag(kp):
        nonlocal first_bad
        b = kp % 2
        k0 = kp * nb
        kb = min(nb, n - k0)
        if rank == layout.owner(kp):
            Lkk = bc[b][: nb * nb].view(nb, nb)
            Wv = bc[b][nb * nb:].view(nwb, 128, 128)
            sl = layout.slot(kp)
            blk = A_loc[sl * nb: sl * nb + kb, k0:k0 + kb]
            info = local.potrf_diag_(blk, sigma2, Wv)
            bad = info.to(torch.int64)
            first_bad = torch.minimum(first_bad, torch.where(bad > 0, bad + k0, torch.full_like(bad, 2 ** 62)))
            Lkk[:kb, :kb].copy_(blk)
            if winv_store is not None:           # the owner keeps the block inverses for the substitution that follows
                winv_store[kp] = Wv.clone()

    def bcast(kp):
        if W > 1:
            dist.broadcast(bc[kp % 2], src=layout.owner(kp), group=group)

    def solve(kp):
        b = kp % 2
        k0 = kp * nb
        kb = min(nb, n - k0)
        k1 = k0 + kb
        if k1 >= n:
            return
        Lkk = bc[b][: nb * nb].view(nb, nb)
        Wv = bc[b][nb * nb:].view(nwb, 128, 128)
        lr0 = layout.slots_after(kp) * nb
        m = rows_f - lr0
        if m > 0:
            local.trsm_right_(A_loc[lr0:rows_f, k0:k1], Lkk[:kb, :kb], Wv)
        if W > 1:
            ms = (npan - (kp + 1) + W - 1) // W
            snd = send[b][: ms * nb, :kb]
            if m > 0:
                snd[:m].copy_(A_loc[lr0:rows_f, k0:k1])
            if m < ms * nb:
                snd[m:].zero_()

    def gather(kp):
        """Returns (P, lr0, m): the block column below the diagonal block in global row order, the first local row below
        it and the number of local rows below it."""
        b = kp % 2
        k0 = kp * nb
        kb = min(nb, n - k0)
        k1 = k0 + kb
        if k1 >= n:
            return None, rows_f, 0
        lr0 = layout.slots_after(kp) * nb
        m = rows_f - lr0
        if W == 1:
            return A_loc[lr0:rows_f, k0:k1], lr0, m
        ms = (npan - (kp + 1) + W - 1) // W
        snd = send[b][: ms * nb, :kb]
        rcv = recv[b].view(-1)[: W * ms * nb * kb].view(W, ms * nb, kb)
        dist.all_gather_into_tensor(rcv.view(-1, kb), snd if snd.is_contiguous() else snd.contiguous(), group=group)
        shift = (kp + 1) % W                      # panel kp+1+t sits on rank (kp+1+t) % W, relative slot t // W
        order = [(shift + i) % W for i in range(W)]
        P = rcv[order].view(W, ms, nb, kb).permute(1, 0, 2, 3).reshape(-1, kb)[: n - k1]
        return P.contiguous(), lr0, m

    ev = [torch.cuda.Event() for _ in range(4)] if use_streams else None
    marks = []                                     # (phase that just ended, event) when phase_ms is requested

    def tick(name):
        if phase_ms is not None and A_loc.is_cuda and not use_streams:
            e = torch.cuda.Event(enable_timing=True)
            e.record(main)
            marks.append((name, e))

    tick(None)
    diag(0); tick('diag')
    bcast(0); tick('bcast')
    solve(0); tick('solve')
    cur = gather(0); tick('gather')
    for kp in range(npan - 1):
        k0 = kp * nb
        k1 = k0 + nb
        k2 = min(k1 + nb, n)
        P, lr0, m = cur
        rb, cb = (lr0 // nb) * W * nb, None
        # U1: the next block column only (lower part, mask in global coordinates)
        if m > 0:
            local.syrk_cyclic_(A_loc[lr0:rows_f, k1:k2], A_loc[lr0:rows_f, k0:k1], P[: k2 - k1], rb, k1, nb, W, rank)
        tick('u1')
        # U2 on rows [lr0, rows_f) x columns [k2, n), in two halves when look-ahead is on
        def update(r_lo, r_hi):
            if r_hi > r_lo and k2 < n:
                local.syrk_cyclic_(A_loc[r_lo:r_hi, k2:n], A_loc[r_lo:r_hi, k0:k1], P[k2 - k1:],
                                   (r_lo // nb) * W * nb, k2, nb, W, rank)
        if not use_streams:
            update(lr0, rows_f); tick('u2')
            diag(kp + 1); tick('diag')
            bcast(kp + 1); tick('bcast')
            solve(kp + 1); tick('solve')
            cur = gather(kp + 1); tick('gather')
            continue
        # split point: about a third of the local rows (whole panels), enough to cover the diagonal factorisation
        nslots = (rows_f - lr0 + nb - 1) // nb
        mid = lr0 + min(rows_f - lr0, max(1, nslots // 3) * nb)
        ev[0].record(main)
        side.wait_event(ev[0])
        with _Ctx(side):
            diag(kp + 1)
            ev[1].record(side)
        update(lr0, mid)                           # U2a   || diag(kp+1)
        main.wait_event(ev[1])
        bcast(kp + 1)                              # collective on the main stream: the SMs are free between U2a and U2b
        ev[2].record(main)
        side.wait_event(ev[2])
        with _Ctx(side):
            solve(kp + 1)
            ev[3].record(side)
        update(mid, rows_f)                        # U2b   || solve(kp+1)
        main.wait_event(ev[3])
        cur = gather(kp + 1)                       # collective on the main stream after U2b
    if use_streams:
        main.wait_stream(side)
    if W > 1:
        dist.all_reduce(first_bad, op=dist.ReduceOp.MIN, group=group)
    fb = int(first_bad.item())
    if marks:
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            phase_ms[name] = phase_ms.get(name, 0.0) + e0.elapsed_time(e1)
    return 0 if fb >= 2 ** 62 else fb


def solve_lower_block_cyclic_(A_loc: torch.Tensor, B_loc: torch.Tensor, layout: BlockCyclic,
                              n_factor: Optional[int] = None, local=CudaLocalOps, group=None,
                              winv_store: Optional[dict] = None) -> torch.Tensor:
    """B_loc <- L^-1 B for the block-cyclically stored factor; B_loc [local_rows(, of the factored block), nrhs]."""
    W, rank, nb = layout.world, layout.rank, layout.nb
    n = layout.n if n_factor is None else n_factor
    npan = (n + nb - 1) // nb
    nrhs = int(B_loc.shape[1])
    my = [p for p in layout.my_panels() if p < npan]
    rows_f = sum(min(nb, n - p * nb) for p in my)
    xk = torch.empty((nb, nrhs), dtype=B_loc.dtype, device=B_loc.device)
    for kp in range(npan):
        k0 = kp * nb
        kb = min(nb, n - k0)
        owner = layout.owner(kp)
        if rank == owner:
            s = layout.slot(kp)
            bk = B_loc[s * nb: s * nb + kb]
            local.trsm_left_(A_loc[s * nb: s * nb + kb, k0:k0 + kb], bk, None if winv_store is None else winv_store.get(kp))
            xk[:kb].copy_(bk)
        if W > 1:
            dist.broadcast(xk, src=owner, group=group)
        lr0 = layout.slots_after(kp) * nb
        if rows_f - lr0 > 0:
            local.solve_update_(B_loc[lr0:rows_f], A_loc[lr0:rows_f, k0:k0 + kb], xk[:kb])
    return B_loc


def solve_lower_block_cyclic_p2p_(A_loc: torch.Tensor, B_loc: torch.Tensor, layout: BlockCyclic, xch: 'PanelExchange',
                                  n_factor: Optional[int] = None, winv_store: Optional[dict] = None) -> torch.Tensor:
    """``solve_lower_block_cyclic_`` without a collective: the owner of panel kp publishes its solved block x_kp into every
    peer's memory together with a flag (one launch, ``gabo_push_signal_f64``), the peers' streams wait for the flag
    (``cuStreamWaitValue32``).  The dependency chain is x_kp -> update of the rows of panel kp+1 -> solve of panel kp+1 -> ...,
    so the owner of panel kp+1 updates THOSE rows first, solves and publishes x_kp+1, and only then updates the rest of its
    rows with x_kp -- the bulk updates are off the critical path.  (Round 1: 64 NCCL broadcasts of a 512-vector, 6.8 ms on 8
    GPUs against 4.0 ms on one.)  x blocks land in the (idle) diagonal-block buffer 0 of the exchange area, one slot per
    panel, so no slot is reused within a solve; callers separate successive solves by a rank-wide synchronisation (the
    log-likelihood all-reduce does)."""
    from . import ops
    W, rank, nb = layout.world, layout.rank, layout.nb
    n = layout.n if n_factor is None else n_factor
    npan = (n + nb - 1) // nb
    nrhs = int(B_loc.shape[1])
    if W == 1 or npan * nb * nrhs > xch.bc_elems or not B_loc.is_contiguous():
        return solve_lower_block_cyclic_(A_loc, B_loc, layout, n_factor=n_factor, winv_store=winv_store)
    my = [p for p in layout.my_panels() if p < npan]
    rows_f = sum(min(nb, n - p * nb) for p in my)
    me = xch.peers.rank
    others = [r for r in range(W) if r != rank]
    base = xch.solve_epoch
    xch.solve_epoch += npan + 1
    FLAG = 24
    if winv_store is not None and all(p in winv_store for p in my):
        # the whole sequence as ONE host call (the Python loop below costs ~100 us of host time per step: host-bound)
        winv_ptrs = [winv_store[p].data_ptr() if p in winv_store else 0 for p in range(npan)]
        ops.solve_cyclic_p2p_(A_loc, B_loc, n, nb, W, rank, winv_ptrs, [xch.bc_ptr(r, 0) for r in range(W)],
                              [xch.flag_ptr(r, FLAG) for r in range(W)], base)
        return B_loc
    xloc = xch.bc_local(0)[: npan * nb * nrhs].view(npan, nb, nrhs)       # x_kp as received from its owner

    def solve_and_publish(kp):
        k0 = kp * nb
        kb = min(nb, n - k0)
        s = layout.slot(kp)
        bk = B_loc[s * nb: s * nb + kb]
        ops.trsm_(A_loc[s * nb: s * nb + kb, k0:k0 + kb], bk, trans=False, winv_all=None if winv_store is None else winv_store.get(kp))
        if kp + 1 < npan:                                               # the last block has no consumer
            off = 8 * kp * nb * nrhs
            ops.push_signal(bk, [xch.bc_ptr(r, 0) + off for r in others], [xch.flag_ptr(r, FLAG) for r in others], base + kp + 1)
        return bk

    xk = solve_and_publish(0) if rank == layout.owner(0) else None
    for kp in range(npan - 1):
        k0 = kp * nb
        kb = min(nb, n - k0)
        if rank != layout.owner(kp):
            ops.flag_wait_geq(xch.flag_ptr(me, FLAG), base + kp + 1)
            xk = xloc[kp, :kb]
        lr0 = layout.slots_after(kp) * nb
        xnext = None
        if rank == layout.owner(kp + 1):                                # critical path first
            kb1 = min(nb, n - (kp + 1) * nb)
            ops.solve_update_(B_loc[lr0: lr0 + kb1], A_loc[lr0: lr0 + kb1, k0:k0 + kb], xk)
            xnext = solve_and_publish(kp + 1)
            lr0 += kb1
        if rows_f - lr0 > 0:
            ops.solve_update_(B_loc[lr0:rows_f], A_loc[lr0:rows_f, k0:k0 + kb], xk)
        if xnext is not None:
            xk = xnext
    return B_loc


def gp_loglik_block_cyclic(A_loc: torch.Tensor, V_loc: torch.Tensor, layout: BlockCyclic, n_factor: Optional[int] = None,
                           local=CudaLocalOps, group=None) -> float:
    """-0.5 n p log(2 pi) - p sum log L_ii - 0.5 ||V||^2 from the distributed factor and V = L^-1 (y - m)."""
    W, nb = layout.world, layout.nb
    n = layout.n if n_factor is None else n_factor
    npan = (n + nb - 1) // nb
    acc = torch.zeros((1,), dtype=torch.float64, device=A_loc.device)
    for p in layout.my_panels():
        if p >= npan:
            continue
        s = layout.slot(p)
        kb = min(nb, n - p * nb)
        acc += local.loglik_block(A_loc[s * nb: s * nb + kb, p * nb: p * nb + kb], V_loc[s * nb: s * nb + kb])
    if W > 1:
        dist.all_reduce(acc, group=group)
    return float(acc.item())


# ---- peer-memory plumbing for the NVLink-mirrored Gram -----------------------------------------------------------------
class _RawCudaArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it without copying."""

    def __init__(self, ptr, shape, typestr='<f8'):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


class PeerBuffers:
    """One [rows, cols] FP64 buffer per rank, allocated with cudaMalloc through the C-ABI, exported with CUDA IPC and opened
    by every other rank of the node on its own device with lazy peer access.  `local` is this rank's buffer as a torch
    tensor (no copy); `ptrs[r]` is the device pointer of rank r's buffer in THIS process."""

    def __init__(self, rows: int, cols: int, device, group=None):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self._L = L
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        self._opened = []
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            _lib.check('gabo_shared_alloc', L.gabo_shared_alloc(max(rows * cols * 8, 8), C.byref(p)))
            self._base = p.value
            self.local = torch.as_tensor(_RawCudaArray(self._base, (rows, cols)), device=self.device)
            self.ptrs = [None] * self.world
            self.ptrs[self.rank] = self._base
            if self.world > 1:
                hbuf = C.create_string_buffer(64)
                _lib.check('gabo_ipc_export', L.gabo_ipc_export(self._base, hbuf))
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(hbuf.raw), group=group)
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    q = C.c_void_p()
                    _lib.check('gabo_ipc_open', L.gabo_ipc_open(h, C.byref(q)))
                    self.ptrs[r] = q.value
                    self._opened.append(q.value)
                torch.cuda.synchronize(self.device)
                dist.barrier(group=group)

    def close(self, group=None):
        """Unmap the peers' buffers, then free the local one (collective: every rank must have unmapped first)."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for q in self._opened:
                self._L.gabo_ipc_close(q)
            self._opened = []
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=group)
            self.local = None
            if self._base:
                self._L.gabo_shared_free(self._base)
                self._base = None


def default_recompute_frac256(world: int) -> int:
    """Share (of 256) of the cross-GPU upper panel pairs that the hybrid Gram recomputes instead of mirroring; measured on
    B200 / NVSwitch at N = 65536, D = 51 (tools/dist_gram_check.py): 2 GPUs 80 (8.4 ms; mirror-all 11.1, recompute-all 10.1),
    4 GPUs 144-160 (5.0 ms; 112: 5.8, 128: 5.35), 8 GPUs 160 (2.7 ms; mirror-all 4.2, recompute-all 3.07, 192: 2.8)."""
    return {1: 0, 2: 80, 3: 128}.get(int(world), 152 if int(world) == 4 else 160)


def default_staged_frac256(world: int) -> int:
    """Recomputed share (of 256) for the STAGED Gram (copy-engine mirror blocks): the SMs no longer stall on remote stores, so
    the balance point is FP64 pipe vs NVLink bandwidth alone -- lower-tile time (1 + 0.88 f) against (1 - f) x cross-rank bytes
    per rank / the copy rate.  Measured on B200 / NVSwitch at N = 65536, D = 51 (tools/dist_gram_check.py, GABO_GRAM_MODE=staged;
    profiles/r02_gram_staged_*): 2 GPUs: share 0 -> 7.75 ms (direct stores: 8.49), the copy engines sustain >= 555 GB/s; 8 GPUs:
    the copy engines sustain only ~300-350 GB/s per GPU in the all-to-all pattern (32: 5.3 ms, 64: 4.65, 112: 3.72; with 64/256 of
    the slots as direct stores on top: 128: 2.78 ms = the direct-store optimum 2.79), so staging is the default at 2 GPUs only."""
    return {1: 0, 2: 0, 3: 64, 4: 96}.get(int(world), 128)


def make_staged_gram_state(n: int, layout: BlockCyclic, device, recompute_frac256: Optional[int] = None, split: int = 1,
                           direct_frac256: int = 0):
    """Scratch of the staged Gram for this rank (ops.StagedGramState): allocate once, reuse for every Gram of that size.
    Of every (row panel, destination rank) block list, recompute_frac256 / 256 is recomputed by the destination, the next
    direct_frac256 / 256 stored remotely by the kernel, the rest staged locally and moved by copy engine."""
    from . import ops
    f = default_staged_frac256(layout.world) if recompute_frac256 is None else int(recompute_frac256)
    return ops.StagedGramState(n, layout.world, layout.rank, f, split, device, direct_frac256=int(direct_frac256))


def gram_block_cyclic_mirrored(kind: int, X: torch.Tensor, layout: BlockCyclic, peers: PeerBuffers, beta: float,
                               variance: float, workspace=None, group=None, sync: bool = True,
                               recompute_frac256: int = 0, xch: Optional['PanelExchange'] = None, staged=None) -> torch.Tensor:
    """Symmetric Gram K(X,X) into the block-cyclic row panels of all ranks: each rank computes the lower tiles of its own
    panels once and writes their mirror images into the owning peers' memory from inside the kernel.  With
    ``recompute_frac256 = f``, f/256 of the panel pairs owned by different ranks are recomputed by the row owner instead of
    mirrored, which balances the FP64 pipe against NVLink.  ``staged`` (``make_staged_gram_state``): the mirror blocks for other
    ranks are staged locally and moved by copy engine instead of being stored remotely by the SMs.

    Contract between successive Grams into the same buffers: a rank may start writing the next K into a peer while that peer
    still reads the previous one -- callers must end every use of K with a rank-wide synchronisation (bench.py: the all-reduce
    and host read of the log-likelihood)."""
    from . import ops
    if layout.nb != 512:
        raise ValueError('the mirrored Gram uses 512-row panels')
    if staged is not None:
        # mirror blocks for other ranks leave through a local staging buffer + copy engines (``recompute_frac256`` is the one
        # the state was built with); on return the stream already waits for this rank's outgoing copies
        ops.points_gram_cyclic_staged(kind, X, peers.ptrs, int(peers.local.stride(0)), layout.world, layout.rank, staged, beta,
                                      variance, workspace)
    else:
        ops.points_gram_cyclic(kind, X, peers.ptrs, int(peers.local.stride(0)), layout.world, layout.rank, beta, variance,
                               workspace, recompute_frac256=recompute_frac256)
    if sync and xch is not None and layout.world > 1:
        # stream-ordered handshake instead of a host barrier: tell every peer that this rank's mirror stores are out, and
        # make the stream wait for the same message from every peer (flags 16 + source rank of the exchange area).  The
        # caller guarantees that no peer still READS its panels of the previous matrix (in the bench: the all-reduce and
        # host read of the log-likelihood end every step).
        xch.gram_epoch += 1
        me = layout.rank
        others = [r for r in range(layout.world) if r != me]
        ops.flag_signal([xch.flag_ptr(r, 16 + me) for r in others], xch.gram_epoch)
        for r in others:
            ops.flag_wait_geq(xch.flag_ptr(me, 16 + r), xch.gram_epoch)
    elif sync:
        torch.cuda.synchronize(X.device)
        if layout.world > 1:
            dist.barrier(group=group)      # remote mirror stores of every rank have landed
    return peers.local


def gram_block_cyclic_rows_i8(kind: int, X: torch.Tensor, layout: BlockCyclic, K_local: torch.Tensor, beta: float, variance: float,
                              workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Block-cyclic sphere Gram with NO exchange: this rank computes the full rows of its 512-row panels with the int8-tensor-core
    engine (ops.points_gram_cyclic_rows_i8), bit-identical to the rows of the single-GPU int8 Gram.  Symmetry is not exploited
    across GPUs (every entry is computed twice in the job), but nothing crosses NVLink and no handshake is needed: K_local is
    private to the rank.  Measured at N = 65536, D = 51 (one GPU: full rows 19.9 ms): 10.0 / 5.0 / 2.5 ms at 2 / 4 / 8 GPUs against
    7.8 / 5.2 / 2.87 ms for the mirrored DMMA variants (gram_block_cyclic_mirrored) -- the default from 4 GPUs on.
    Requires n % 512 == 0 and dim <= 64."""
    from . import ops
    if layout.n % layout.nb != 0 or layout.nb != 512:
        raise ValueError('gram_block_cyclic_rows_i8 needs 512-row panels and n % 512 == 0')
    return ops.points_gram_cyclic_rows_i8(kind, X, K_local, layout.world, layout.rank, beta=beta, variance=variance, workspace=workspace)


def default_gram_mode(world: int, n: int, dim: int) -> str:
    """'staged' (copy-engine mirror blocks) at 2 GPUs, 'i8rows' (exchange-free int8 rows) from 4 GPUs on where it applies,
    otherwise 'direct' (hybrid mirror / recompute with NVLink stores from the kernel)."""
    if int(world) == 2:
        return 'staged'
    if int(world) >= 4 and n % 512 == 0 and 1 <= dim <= 64:
        return 'i8rows'
    return 'direct'


def spd_gram_block_cyclic(kind: int, prep, layout: BlockCyclic, peers: PeerBuffers, beta: float, variance: float,
                          group=None, sync: bool = True, xch: Optional['PanelExchange'] = None) -> torch.Tensor:
    """Symmetric SPD Gram (affine-invariant Gaussian / Laplace, Stein) into the block-cyclic row panels of all ranks: the
    lower 32 x 128 blocks are dealt round-robin to the ranks (balanced triangular work), and every block and its mirror image
    are stored from inside the kernel into the HBM of the ranks owning those rows (``gabo_spd_gram_cyclic_f64``).  The pair
    kernel is compute-bound (8 B per ~2.5 kflop), so the NVLink stores ride for free.  Completion: the stream-ordered flag
    handshake of ``gram_block_cyclic_mirrored`` when ``xch`` is given, else stream sync + barrier."""
    from . import ops
    if layout.nb != 512:
        raise ValueError('the cyclic SPD Gram uses 512-row panels')
    ops.spd_gram_cyclic(kind, prep, peers.ptrs, int(peers.local.stride(0)), layout.world, layout.rank, beta, variance)
    if sync and xch is not None and layout.world > 1:
        xch.gram_epoch += 1
        me = layout.rank
        others = [r for r in range(layout.world) if r != me]
        ops.flag_signal([xch.flag_ptr(r, 16 + me) for r in others], xch.gram_epoch)
        for r in others:
            ops.flag_wait_geq(xch.flag_ptr(me, 16 + r), xch.gram_epoch)
    elif sync:
        torch.cuda.synchronize(peers.local.device)
        if layout.world > 1:
            dist.barrier(group=group)
    return peers.local


class PanelExchange:
    """Peer-memory exchange area of the distributed factorisation (one NVSwitch node, world <= 8).

    Every rank allocates, exports (CUDA IPC) and maps on all peers:
      * three block-column buffers P[b] [n][nb]: the block column below panel kp, in GLOBAL row order, lands in P[kp % 3]
        of EVERY rank by remote stores from the ranks that solve its rows (gabo_trsm_right_scatter_f64);
      * three buffers bc[b] for the factored diagonal block + the inverses of its 128-blocks, pushed by the owner with
        copy-engine transfers;
      * 32-bit flags: [0] "bc of panel kp has arrived", [1 + r] "rank r's rows of block column kp have arrived".
    Flags carry epoch + kp + 1 and are waited for with cuStreamWaitValue32(>=): no collective kernel, no SM taken by a
    rank that waits.  Three buffers are what makes one-panel look-ahead safe: a peer can only start writing buffer b for
    panel kp + 3 after this rank has finished the trailing update of panel kp (see potrf_block_cyclic_p2p_)."""

    NBUF = 3

    def __init__(self, layout: BlockCyclic, n: int, device, group=None):
        self.layout = layout
        self.n, self.nb = int(n), int(layout.nb)
        if layout.world > 8:
            raise ValueError('PanelExchange: at most 8 ranks (one NVSwitch node)')
        if self.nb % 2:
            raise ValueError('PanelExchange: even panel width required')
        self.nwb = (self.nb + 127) // 128
        self.p_elems = self.n * self.nb
        self.bc_elems = self.nb * self.nb + self.nwb * 128 * 128
        # the critical-chain schedule lets the owners run up to `world` panels ahead of a rank that owns none of them, so the
        # buffers rotate over world + 1 (the look-ahead schedule uses the first three)
        self.NBUF = max(3, layout.world + 1)
        self.flag_off = self.NBUF * (self.p_elems + self.bc_elems)          # in doubles
        total = self.flag_off + 32
        self.peers = PeerBuffers(1, total, device, group=group)
        self.device = torch.device(device)
        self.epoch = 0
        self.gram_epoch = 0
        self.solve_epoch = 0
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self._cstate = None
        import ctypes as _C
        from . import _lib as _l
        e1, e2 = _C.c_void_p(), _C.c_void_p()
        with torch.cuda.device(self.device):
            _l.check('gabo_event_create', _l.lib().gabo_event_create(_C.byref(e1)))
            _l.check('gabo_event_create', _l.lib().gabo_event_create(_C.byref(e2)))
        self.ev_u1, self.ev_side = e1.value, e2.value
        self.peers.local[0, self.flag_off:].zero_()
        self._warm_up()
        torch.cuda.synchronize(self.device)
        if layout.world > 1:
            dist.barrier(group=group)

    def _warm_up(self):
        """Run every kernel of the factorisation once, single-rank, with nothing pending.  The first launch of a kernel
        loads its module, which synchronises the whole context: were that to happen on the side stream while the main
        stream is stalled on a peer's flag, the host would block and the ranks could wait for each other forever."""
        from . import ops
        nb = self.nb
        nw = min(self.n, 3 * nb)
        if nw < 1:
            return
        A = torch.zeros((nw, nw), dtype=torch.float64, device=self.device)
        A.diagonal().fill_(2.0)
        potrf_block_cyclic_p2p_(A, BlockCyclic(nw, nb, 1, 0), 0.0, self, winv_store={})
        scratch = self.flag_ptr(self.peers.rank, 31)
        ops.flag_signal([scratch], 1)
        ops.flag_wait_geq(scratch, 1)
        # ... and the kernels of the substitution (solve_lower_block_cyclic_p2p_): one-right-hand-side solve, row update, push
        v = torch.ones((nw, 1), dtype=torch.float64, device=self.device)
        ops.trsm_(A[:nb, :nb] if nw >= nb else A, v[: min(nb, nw)], trans=False, winv_all=None)
        if nw > nb:
            ops.solve_update_(v[nb:], A[nb:, :nb], v[:nb])
        ops.push_signal(v[: min(nb, nw)].contiguous(), [self.bc_ptr(self.peers.rank, 1)], [scratch], 2)

    # raw pointers in THIS process of rank r's areas
    def p_ptr(self, r: int, b: int) -> int:
        return self.peers.ptrs[r] + 8 * b * self.p_elems

    def bc_ptr(self, r: int, b: int) -> int:
        return self.peers.ptrs[r] + 8 * (self.NBUF * self.p_elems + b * self.bc_elems)

    def flag_ptr(self, r: int, idx: int) -> int:
        return self.peers.ptrs[r] + 8 * self.flag_off + 4 * idx

    # torch views of the local areas
    def p_local(self, b: int) -> torch.Tensor:
        o = b * self.p_elems
        return self.peers.local[0, o: o + self.p_elems].view(self.n, self.nb)

    def bc_local(self, b: int) -> torch.Tensor:
        o = self.NBUF * self.p_elems + b * self.bc_elems
        return self.peers.local[0, o: o + self.bc_elems]

    def c_state(self, npan: int, nslots: int, rows: int):
        """Scratch of the one-call C drivers, allocated once per exchange (no allocation while streams wait on peers' flags)."""
        from . import _lib
        L = _lib.lib()
        st = self._cstate
        if st is None or st['npan'] < npan or st['nslots'] < nslots or st['rows'] < rows:
            with torch.cuda.device(self.device):
                st = {'npan': npan, 'nslots': nslots, 'rows': rows,
                      'info': torch.zeros((npan,), dtype=torch.int32, device=self.device),
                      'winv': torch.empty((nslots, self.nwb, 128, 128), dtype=torch.float64, device=self.device),
                      'ws_diag': torch.empty(int(max(L.gabo_potrf_workspace_bytes(self.nb), L.gabo_trsm_right_workspace_bytes(self.nb, self.nb))),
                                             dtype=torch.uint8, device=self.device),
                      'ws_trsm': torch.empty(int(L.gabo_trsm_right_workspace_bytes(rows, self.nb)), dtype=torch.uint8, device=self.device)}
            self._cstate = st
        return st

    def close(self, group=None):
        from . import _lib
        if getattr(self, 'ev_u1', None):
            torch.cuda.synchronize(self.device)
            _lib.lib().gabo_event_destroy(self.ev_u1)
            _lib.lib().gabo_event_destroy(self.ev_side)
            self.ev_u1 = self.ev_side = None
        self.peers.close(group=group)


def potrf_block_cyclic_p2p_(A_loc: torch.Tensor, layout: BlockCyclic, sigma2: float, xch: PanelExchange,
                            n_factor: Optional[int] = None, group=None, winv_store: Optional[dict] = None) -> int:
    """Distributed Cholesky with the same arithmetic as potrf_block_cyclic_ but with the panel exchange done through peer
    memory and one panel of look-ahead:

        main stream                                   side stream (high priority)
        wait P(kp) flags of all peers
        U1(kp): next block column                 ->  owner: factor diagonal block kp+1, push it to the peers (copy engines),
        U2(kp): rest of the trailing matrix               raise their bc flags
                                                      all:   wait bc flag, solve own rows of block column kp+1, storing the
                                                             result into P[(kp+1) % 3] of EVERY rank (NVLink stores),
                                                             raise the peers' P flags

    so that the whole panel chain of step kp+1 is hidden under U2(kp), and nothing on the side stream needs an SM that the
    update kernel holds for long: the panel kernels fit beside a resident update CTA, the transfers use copy engines or
    ride on the solve kernel's stores, and waiting is a stream memory operation.  Returns LAPACK-style info."""
    from . import ops
    W, rank, nb = layout.world, layout.rank, layout.nb
    n = layout.n if n_factor is None else n_factor
    if n % nb != 0 and n != layout.n:
        raise ValueError('n_factor must be a multiple of the panel size')
    if n > xch.n or nb != xch.nb:
        raise ValueError('PanelExchange too small for this factorisation')
    npan = (n + nb - 1) // nb
    dev = A_loc.device
    nwb = xch.nwb
    main = torch.cuda.current_stream(dev)
    side = xch.side
    base = xch.epoch
    xch.epoch += npan + 1
    others = [r for r in range(W) if r != rank]
    me = xch.peers.rank                               # index of this process in the exchange area (== rank unless W == 1)
    if W > 1 and (W != xch.peers.world or me != rank):
        raise ValueError('layout and PanelExchange disagree on world / rank')
    my = [p for p in layout.my_panels() if p < npan]
    rows_f = sum(min(nb, n - p * nb) for p in my)
    if os.environ.get('GABO_DIST_PYLOOP', '0') != '1' and nb % 128 == 0 and 128 <= nb <= 512:
        # the whole loop as ONE C call (gabo_potrf_cyclic_p2p_f64): the Python loop below -- kept as the readable specification --
        # spends 0.2-0.3 ms of host time per panel, as much as the device needs for a panel step at 8 GPUs
        import ctypes as C
        from . import _lib
        st = xch.c_state(npan, max(len(my), 1), max(rows_f, 1))
        if os.environ.get('GABO_DIST_SCHED', 'chain') == 'chain' and nb == 512:
            # critical-chain schedule (gabo_potrf_cyclic_chain_f64): next owner first, U1 released by one 512-block, deeper buffers
            nbuf = xch.NBUF
            Ph = (C.c_void_p * (nbuf * W))(*[xch.p_ptr(r if W > 1 else me, b) for r in range(W) for b in range(nbuf)])
            Bh = (C.c_void_p * (nbuf * W))(*[xch.bc_ptr(r if W > 1 else me, b) for r in range(W) for b in range(nbuf)])
            Fh = (C.c_void_p * W)(*[xch.flag_ptr(r if W > 1 else me, 0) for r in range(W)])
            rc = _lib.lib().gabo_potrf_cyclic_chain_f64(n, nb, W, rank, A_loc.data_ptr(), max(int(A_loc.stride(0)), int(A_loc.shape[1])),
                                                       float(sigma2), nbuf, C.cast(Ph, C.c_void_p), C.cast(Bh, C.c_void_p),
                                                       C.cast(Fh, C.c_void_p), base, st['info'].data_ptr(),
                                                       st['winv'].data_ptr() if winv_store is not None else None,
                                                       st['ws_diag'].data_ptr(), st['ws_diag'].numel(), st['ws_trsm'].data_ptr(),
                                                       st['ws_trsm'].numel(), main.cuda_stream, side.cuda_stream)
            _lib.check('gabo_potrf_cyclic_chain_f64', rc)
        else:
            rc = None
        Ph = (C.c_void_p * (3 * W))(*[xch.p_ptr(r if W > 1 else me, b) for r in range(W) for b in range(3)])
        Bh = (C.c_void_p * (3 * W))(*[xch.bc_ptr(r if W > 1 else me, b) for r in range(W) for b in range(3)])
        Fh = (C.c_void_p * W)(*[xch.flag_ptr(r if W > 1 else me, 0) for r in range(W)])
        if rc is None:
            rc = _lib.lib().gabo_potrf_cyclic_p2p_f64(n, nb, W, rank, A_loc.data_ptr(), max(int(A_loc.stride(0)), int(A_loc.shape[1])),
                                                     float(sigma2), C.cast(Ph, C.c_void_p), C.cast(Bh, C.c_void_p), C.cast(Fh, C.c_void_p),
                                                     base, st['info'].data_ptr(), st['winv'].data_ptr() if winv_store is not None else None,
                                                     st['ws_diag'].data_ptr(), st['ws_diag'].numel(), st['ws_trsm'].data_ptr(),
                                                     st['ws_trsm'].numel(), main.cuda_stream, side.cuda_stream, xch.ev_u1, xch.ev_side)
            _lib.check('gabo_potrf_cyclic_p2p_f64', rc)
        if winv_store is not None:
            for p_ in my:                                   # views: valid until the next factorisation through this exchange
                winv_store[p_] = st['winv'][layout.slot(p_)]
        info = st['info'][:npan].to(torch.int64)
        k0s = torch.arange(npan, device=dev, dtype=torch.int64) * nb
        first_bad = torch.where(info > 0, info + k0s, torch.full_like(info, 2 ** 62)).min().reshape(1)
        if W > 1:
            dist.all_reduce(first_bad, op=dist.ReduceOp.MIN, group=group)
        fb = int(first_bad.item())
        return 0 if fb >= 2 ** 62 else fb
    first_bad = torch.full((1,), 2 ** 62, dtype=torch.int64, device=dev)
    bc_bytes = 8 * xch.bc_elems

    def parts(kp):
        bcl = xch.bc_local(kp % xch.NBUF)
        return bcl[: nb * nb].view(nb, nb), bcl[nb * nb:].view(nwb, 128, 128)

    def diag_and_push(kp):
        """Owner only: factor the diagonal block of panel kp, hand factor + block inverses to every peer."""
        nonlocal first_bad
        b = kp % xch.NBUF
        k0 = kp * nb
        kb = min(nb, n - k0)
        Lkk, Wv = parts(kp)
        sl = layout.slot(kp)
        blk = A_loc[sl * nb: sl * nb + kb, k0:k0 + kb]
        info = ops.potrf_diag_(blk, sigma2, Wv)
        bad = info.to(torch.int64)
        first_bad = torch.minimum(first_bad, torch.where(bad > 0, bad + k0, torch.full_like(bad, 2 ** 62)))
        Lkk[:kb, :kb].copy_(blk)
        if winv_store is not None:
            winv_store[kp] = Wv.clone()
        if k0 + kb >= n:
            return                                   # last panel: nobody solves against it
        for r in others:
            ops.copy_async(xch.bc_ptr(r, b), xch.bc_ptr(me, b), bc_bytes)
        ops.flag_signal([xch.flag_ptr(r, 0) for r in others], base + kp + 1)

    def solve_and_scatter(kp):
        """Every rank: rows of block column kp below the diagonal block, delivered to all ranks."""
        b = kp % xch.NBUF
        k0 = kp * nb
        kb = min(nb, n - k0)
        k1 = k0 + kb
        if k1 >= n:
            return
        if rank != layout.owner(kp):
            ops.flag_wait_geq(xch.flag_ptr(me, 0), base + kp + 1)
        Lkk, Wv = parts(kp)
        s0 = layout.slots_after(kp)
        lr0 = s0 * nb
        m = rows_f - lr0
        if m > 0:
            ops.trsm_right_scatter_(A_loc[lr0:rows_f, k0:k1], Lkk[:kb, :kb], Wv, [xch.p_ptr(r if W > 1 else me, b) for r in range(W)],
                                    nb, (s0 * W + rank) * nb - k1, nb, W)
        ops.flag_signal([xch.flag_ptr(r, 1 + rank) for r in others], base + kp + 1)

    ev_u1 = torch.cuda.Event()
    ev_side = torch.cuda.Event()
    if rank == layout.owner(0):
        diag_and_push(0)
    solve_and_scatter(0)
    for kp in range(npan - 1):
        k0 = kp * nb
        k1 = k0 + nb
        k2 = min(k1 + nb, n)
        lr0 = layout.slots_after(kp) * nb
        m = rows_f - lr0
        for r in others:
            ops.flag_wait_geq(xch.flag_ptr(me, 1 + r), base + kp + 1)
        if kp > 0:
            main.wait_event(ev_side)
        P = xch.p_local(kp % xch.NBUF)[: n - k1]
        rb = (lr0 // nb) * W * nb
        if m > 0:
            ops.syrk_cyclic_(A_loc[lr0:rows_f, k1:k2], A_loc[lr0:rows_f, k0:k1], P[: k2 - k1], rb, k1, nb, W, rank)
        ev_u1.record(main)
        side.wait_event(ev_u1)
        with torch.cuda.stream(side):
            if rank == layout.owner(kp + 1):
                diag_and_push(kp + 1)
            solve_and_scatter(kp + 1)
            ev_side.record(side)
        if m > 0 and k2 < n:
            ops.syrk_cyclic_(A_loc[lr0:rows_f, k2:n], A_loc[lr0:rows_f, k0:k1], P[k2 - k1:], rb, k2, nb, W, rank)
    main.wait_stream(side)
    if W > 1:
        dist.all_reduce(first_bad, op=dist.ReduceOp.MIN, group=group)
    fb = int(first_bad.item())
    return 0 if fb >= 2 ** 62 else fb
