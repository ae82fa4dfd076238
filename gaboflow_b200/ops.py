"""Operator layer: torch CUDA tensors in, C-ABI calls (include/gabo_b200.h) on the current stream, torch tensors out.

PyTorch is only the allocator / stream owner here; all arithmetic happens in libgabo_b200.so.  Every function
raises if the input is not a CUDA tensor -- there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib

SPHERE_GAUSSIAN, SPHERE_LAPLACE, SQDIST_GAUSSIAN = 0, 1, 2
SPD_AI_GAUSSIAN, SPD_AI_LAPLACE, SPD_STEIN = 0, 1, 2
FULL, LOWER, SYMM_MIRROR = 0, 1, 2
_UPLO = {'full': FULL, 'lower': LOWER, 'mirror': SYMM_MIRROR, FULL: FULL, LOWER: LOWER, SYMM_MIRROR: SYMM_MIRROR}
ROW_BLOCK = 128   # row-sharding granularity of the Gram kernels


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str, dtype=torch.float64) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f'{name} must be a CUDA tensor (gaboflow_b200 has no CPU path)')
    if t.dtype != dtype:
        raise TypeError(f'{name} must be {dtype}, got {t.dtype}')
    return t


def _rowmajor2d(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError('expected a 2-D tensor')
    if t.stride(1) != 1 and t.shape[1] > 1:
        t = t.contiguous()
    if t.shape[1] == 1 and t.stride(0) < 1:
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] > 1 else int(t.shape[1])


def device_info():
    import ctypes as C
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _lib.check('gabo_device_info', _lib.lib().gabo_device_info(C.byref(a), C.byref(b), C.byref(c)))
    return {'sm_count': a.value, 'cc': (b.value, c.value)}


# --------------------------------------------------------------------------------------------------
I8_MIN_ENTRIES = 1 << 22     # below this the Gram is launch-bound and the extra slicing pass of the int8 engine does not pay


def gram_engine(kind: int, dtype, dim: int, n1: int, n2: int, engine: Optional[str] = None) -> str:
    """Which kernel forms an FP64 Gram: 'i8' = inner products on the integer tensor cores (csrc/points_gram_i8.cu: sphere kinds,
    dim <= 64), 'dmma' = FP64 tensor instruction (csrc/points_gram.cu: every kind and size).  ``engine`` or the environment
    variable GABO_GRAM_ENGINE force one; the default is 'i8' wherever it applies and the problem is large enough."""
    import os
    choice = engine or os.environ.get('GABO_GRAM_ENGINE', 'auto')
    ok = dtype == torch.float64 and kind in (SPHERE_GAUSSIAN, SPHERE_LAPLACE) and 1 <= dim <= 64
    if choice == 'i8':
        if not ok:
            raise ValueError('the int8 Gram engine covers the FP64 sphere kernels with dim <= 64')
        return 'i8'
    if choice == 'dmma' or not ok:
        return 'dmma'
    return 'i8' if n1 * n2 >= I8_MIN_ENTRIES else 'dmma'


def points_gram(kind: int, X: torch.Tensor, X2: Optional[torch.Tensor] = None, beta: float = 1.0,
                variance: float = 1.0, out: Optional[torch.Tensor] = None, row0: int = 0,
                row1: Optional[int] = None, uplo='full', workspace: Optional[torch.Tensor] = None,
                engine: Optional[str] = None) -> torch.Tensor:
    """K[row0:row1, :] of the sphere Gaussian / Laplace or squared-distance Gaussian Gram (FP64 or FP32 by X.dtype)."""
    L = _lib.lib()
    dtype = X.dtype
    if dtype not in (torch.float64, torch.float32):
        raise TypeError('X must be float64 or float32')
    X = _rowmajor2d(_req(X, 'X', dtype))
    n1, dim = int(X.shape[0]), int(X.shape[1])
    if X2 is not None:
        X2 = _rowmajor2d(_req(X2, 'X2', dtype))
        if X2.shape[1] != dim:
            raise ValueError('X and X2 must have the same number of columns')
        n2 = int(X2.shape[0])
    else:
        n2 = n1
    if row1 is None:
        row1 = n1
    rows = row1 - row0
    if out is None:
        out = torch.empty((rows, n2), dtype=dtype, device=X.device)
    else:
        _req(out, 'out', dtype)
        if out.dim() != 2 or out.shape[0] < rows or out.shape[1] != n2 or (n2 > 1 and out.stride(1) != 1):
            raise ValueError('out must be a row-major [row1-row0, n2] tensor')
    if gram_engine(kind, dtype, dim, rows, n2, engine) == 'i8':
        nbytes = int(L.gabo_points_gram_i8_workspace_bytes(n1, n2))
        if workspace is None or workspace.numel() < nbytes or workspace.data_ptr() % 128:
            workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=X.device)
        st = L.gabo_points_gram_i8_f64(int(kind), X.data_ptr(), n1, _ld(X), X2.data_ptr() if X2 is not None else None, n2,
                                       _ld(X2) if X2 is not None else 0, dim, float(beta), float(variance), out.data_ptr(),
                                       _ld(out) if rows > 1 else max(n2, int(out.stride(0))), int(row0), int(row1), _UPLO[uplo], 0, 0,
                                       workspace.data_ptr(), workspace.numel(), _stream())
        _lib.check('gabo_points_gram_i8_f64', st)
        return out
    nbytes = int(L.gabo_points_gram_workspace_bytes(n1, n2, dim))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=X.device)
    fn_name = 'gabo_points_gram_f64' if dtype == torch.float64 else 'gabo_points_gram_f32'
    st = getattr(L, fn_name)(int(kind), X.data_ptr(), n1, _ld(X),
                             X2.data_ptr() if X2 is not None else None, n2, _ld(X2) if X2 is not None else 0,
                             dim, float(beta), float(variance), out.data_ptr(), _ld(out) if rows > 1 else max(n2, int(out.stride(0))),
                             int(row0), int(row1), _UPLO[uplo], workspace.data_ptr(), workspace.numel(), _stream())
    _lib.check(fn_name, st)
    return out


def kdiag_const(n: int, variance: float, device) -> torch.Tensor:
    out = torch.empty((n,), dtype=torch.float64, device=device)
    _lib.check('gabo_kdiag_const_f64', _lib.lib().gabo_kdiag_const_f64(n, float(variance), out.data_ptr(), _stream()))
    return out


# --------------------------------------------------------------------------------------------------
@dataclass
class SpdPrep:
    """Per-point quantities of a set of SPD matrices (see gabo_spd_prep_f64)."""
    d: int
    n: int
    S: Optional[torch.Tensor]        # [n, d*d]
    W: Optional[torch.Tensor]        # [n, d*d]  inverse lower Cholesky factor
    logdet: Optional[torch.Tensor]   # [n]
    logm: Optional[torch.Tensor]     # [n, m]    Mandel(logm S)
    info: torch.Tensor               # [n] int32
    Lt: Optional[torch.Tensor] = None   # [m, ldt] packed lower Cholesky factors, transposed (column operand of the AI kernel)

    def check(self) -> 'SpdPrep':
        """Raise ValueError naming the first input matrix that is not SPD (one host sync)."""
        if self.n and bool(self.info.any().item()):
            bad = int(torch.nonzero(self.info)[0].item())
            raise ValueError(f'input matrix {bad} is not symmetric positive definite '
                             f'(Cholesky pivot {int(self.info[bad].item())} is not positive)')
        return self


def spd_prep(Xm: torch.Tensor, d: int, want_S=True, want_W=True, want_logdet=True, want_logm=False, want_Lt=None) -> SpdPrep:
    Xm = _rowmajor2d(_req(Xm, 'Xmandel'))
    n, m = int(Xm.shape[0]), int(Xm.shape[1])
    if m != d * (d + 1) // 2:
        raise ValueError(f'Mandel length {m} does not match matrix_dim {d}')
    if want_Lt is None:
        want_Lt = want_W                      # the affine-invariant kinds need both halves of the pair operand
    dev = Xm.device
    S = torch.empty((n, d * d), dtype=torch.float64, device=dev) if want_S else None
    W = torch.empty((n, d * d), dtype=torch.float64, device=dev) if want_W else None
    ld = torch.empty((n,), dtype=torch.float64, device=dev) if want_logdet else None
    lm = torch.empty((n, m), dtype=torch.float64, device=dev) if want_logm else None
    ldt = ((n + 31) // 32) * 32               # 256 B aligned rows
    Lt = torch.empty((m, max(ldt, 1)), dtype=torch.float64, device=dev) if want_Lt else None
    info = torch.zeros((max(n, 1),), dtype=torch.int32, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    st = _lib.lib().gabo_spd_prep_f64(Xm.data_ptr(), n, _ld(Xm), d, ptr(S), ptr(W), ptr(ld), ptr(lm), ptr(Lt), max(ldt, 1),
                                      info.data_ptr(), _stream())
    _lib.check('gabo_spd_prep_f64', st)
    return SpdPrep(d, n, S, W, ld, lm, info[:n], Lt)


def spd_gram(kind: int, p1: SpdPrep, p2: Optional[SpdPrep] = None, beta: float = 1.0, variance: float = 1.0,
             out: Optional[torch.Tensor] = None, row0: int = 0, row1: Optional[int] = None, uplo='full') -> torch.Tensor:
    symmetric = p2 is None
    if symmetric:
        p2 = p1
    if p1.d != p2.d:
        raise ValueError('matrix_dim mismatch')
    n1, n2 = p1.n, p2.n
    if row1 is None:
        row1 = n1
    rows = row1 - row0
    dev = p1.info.device
    if out is None:
        out = torch.empty((rows, n2), dtype=torch.float64, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    if kind != SPD_STEIN and (p1.W is None or p2.Lt is None):
        raise ValueError('the affine-invariant kinds need W of the row points and Lt of the column points (spd_prep want_W / want_Lt)')
    st = _lib.lib().gabo_spd_gram_f64(int(kind), p1.d, ptr(p1.W), ptr(p1.S), ptr(p1.logdet), n1,
                                      ptr(p2.S), ptr(p2.Lt), int(p2.Lt.stride(0)) if p2.Lt is not None else 0,
                                      ptr(p2.logdet), n2, 1 if symmetric else 0,
                                      float(beta), float(variance), out.data_ptr(),
                                      _ld(out) if rows > 1 else max(n2, int(out.stride(0))),
                                      int(row0), int(row1), _UPLO[uplo], _stream())
    _lib.check('gabo_spd_gram_f64', st)
    return out


def spd_gram_cyclic(kind: int, p: SpdPrep, bases, ldk: int, world: int, rank: int, beta: float = 1.0,
                    variance: float = 1.0) -> None:
    """Block-cyclic symmetric SPD Gram with peer-memory stores (gabo_spd_gram_cyclic_f64).  `bases`: list of `world` device
    pointers (ints), entry r = base of rank r's local [local_rows, ldk] buffer as mapped in this process."""
    import ctypes as C
    if kind != SPD_STEIN and (p.W is None or p.Lt is None):
        raise ValueError('the affine-invariant kinds need W and Lt (spd_prep want_W / want_Lt)')
    ptr = lambda t: t.data_ptr() if t is not None else None
    arr = (C.c_void_p * world)(*[C.c_void_p(int(b)) for b in bases])
    st = _lib.lib().gabo_spd_gram_cyclic_f64(int(kind), p.d, ptr(p.W), ptr(p.S), ptr(p.logdet), ptr(p.Lt),
                                             int(p.Lt.stride(0)) if p.Lt is not None else 0, p.n, float(beta), float(variance),
                                             C.cast(arr, C.c_void_p), int(ldk), int(world), int(rank), _stream())
    _lib.check('gabo_spd_gram_cyclic_f64', st)


def spd_stein_kdiag(p: SpdPrep, beta: float, variance: float) -> torch.Tensor:
    out = torch.empty((p.n,), dtype=torch.float64, device=p.info.device)
    st = _lib.lib().gabo_spd_stein_kdiag_f64(p.logdet.data_ptr(), p.n, float(beta), float(variance), out.data_ptr(), _stream())
    _lib.check('gabo_spd_stein_kdiag_f64', st)
    return out


def _spd_map(fn_name: str, Y: torch.Tensor, base: torch.Tensor, d: int, mandel: bool, want_info: bool):
    Y = _req(Y, 'input')
    base = _req(base, 'base')
    w = d * (d + 1) // 2 if mandel else d * d
    single = (Y.dim() == 1) if mandel else (Y.dim() == 2)
    Y2 = Y.reshape(1, w) if single else Y.reshape(Y.shape[0], w)
    Y2 = _rowmajor2d(Y2)
    n = int(Y2.shape[0])
    b2 = base.reshape(-1, w).contiguous()
    if b2.shape[0] not in (1, n):
        raise ValueError('base must be one matrix or one per input')
    out = torch.empty((n, w), dtype=torch.float64, device=Y.device)
    info = torch.zeros((max(n, 1),), dtype=torch.int32, device=Y.device)
    st = getattr(_lib.lib(), fn_name)(int(d), Y2.data_ptr(), _ld(Y2), b2.data_ptr(), 0 if b2.shape[0] == 1 else w, n,
                                     1 if mandel else 0, out.data_ptr(), w, info.data_ptr(), _stream())
    _lib.check(fn_name, st)
    out = out.reshape(Y.shape)
    return (out, info[:n]) if want_info else out


def spd_expmap(U: torch.Tensor, S: torch.Tensor, d: int, mandel: bool = False, want_info: bool = False):
    """Expmap_S(U) = S expm(S^-1 U), batched (gabo_spd_expmap_f64).  U [n,d,d] / [d,d] matrices, or Mandel vectors [n,m] / [m]
    with mandel=True; S one base point (same layout) or one per input."""
    return _spd_map('gabo_spd_expmap_f64', U, S, d, mandel, want_info)


def spd_logmap(X: torch.Tensor, S: torch.Tensor, d: int, mandel: bool = False, want_info: bool = False):
    """Logmap_S(X) = S logm(S^-1 X), batched (gabo_spd_logmap_f64); layouts as spd_expmap."""
    return _spd_map('gabo_spd_logmap_f64', X, S, d, mandel, want_info)


def sym_to_mandel(S: torch.Tensor, d: int) -> torch.Tensor:
    """[n, d, d] (or [d, d]) symmetric matrices -> Mandel vectors [n, m] (or [m]) (gabo_sym_to_mandel_f64)."""
    S = _req(S, 'S')
    single = S.dim() == 2
    S2 = _rowmajor2d(S.reshape(-1, d * d))
    n, m = int(S2.shape[0]), d * (d + 1) // 2
    out = torch.empty((n, m), dtype=torch.float64, device=S.device)
    st = _lib.lib().gabo_sym_to_mandel_f64(int(d), S2.data_ptr(), _ld(S2), n, out.data_ptr(), m, _stream())
    _lib.check('gabo_sym_to_mandel_f64', st)
    return out[0] if single else out


def log_gram_(K: torch.Tensor, variance: float = 1.0, scale: float = 1.0) -> torch.Tensor:
    """In place: K <- scale * min(0, log(K / variance)) (gabo_log_gram_f64); scale=1: the shared operand of lanczos_sweep,
    scale=-1: the (squared) distance matrix behind a beta = 1 Gram."""
    K = _req(K, 'K')
    if K.dim() != 2 or (K.stride(1) != 1 and K.shape[1] > 1):
        raise ValueError('K must be a row-major 2-D tensor')
    st = _lib.lib().gabo_log_gram_f64(K.data_ptr(), _ld(K), K.data_ptr(), _ld(K), int(K.shape[0]), int(K.shape[1]), float(variance),
                                      float(scale), _stream())
    _lib.check('gabo_log_gram_f64', st)
    return K


def lanczos_sweep(G: torch.Tensor, ratios: torch.Tensor, variance: float = 1.0, m_max: Optional[int] = None, tol: float = 0.0,
                  seed: int = 1234, want_tridiag: bool = False):
    """Extreme eigenvalues of variance*exp(ratio*G) for every ratio and every log-Gram in G ([n,n] or [n_mat,n,n], contiguous).
    Returns a [n_mat, n_ratios, 4] tensor: lambda_min, lambda_max, residual estimate, steps (gabo_lanczos_sweep_f64)."""
    G = _req(G, 'G')
    G3 = G.reshape(1, *G.shape) if G.dim() == 2 else G
    if G3.dim() != 3 or G3.shape[1] != G3.shape[2] or not G3.is_contiguous():
        raise ValueError('G must be a contiguous [n, n] or [n_mat, n, n] tensor')
    ratios = _req(ratios, 'ratios').contiguous().reshape(-1)
    n_mat, n, nb = int(G3.shape[0]), int(G3.shape[1]), int(ratios.numel())
    m_max = min(n, 2048) if m_max is None else min(int(m_max), n, 2048)
    lib = _lib.lib()
    need = int(lib.gabo_lanczos_sweep_workspace_bytes(n, n_mat, nb, m_max))
    ws = torch.empty((need,), dtype=torch.uint8, device=G.device)
    out = torch.empty((n_mat, nb, 4), dtype=torch.float64, device=G.device)
    al = torch.empty((n_mat, nb, m_max), dtype=torch.float64, device=G.device) if want_tridiag else None
    be = torch.empty((n_mat, nb, m_max), dtype=torch.float64, device=G.device) if want_tridiag else None
    st = lib.gabo_lanczos_sweep_f64(G3.data_ptr(), n, n * n, n, n_mat, ratios.data_ptr(), nb, float(variance), m_max, float(tol),
                                    int(seed), out.data_ptr(), al.data_ptr() if al is not None else None,
                                    be.data_ptr() if be is not None else None, ws.data_ptr(), need, _stream())
    _lib.check('gabo_lanczos_sweep_f64', st)
    return (out, al, be) if want_tridiag else out


# --------------------------------------------------------------------------------------------------
_ws_cache = {}


def potrf_(A: torch.Tensor, sigma2: float = 0.0, check: bool = True) -> torch.Tensor:
    """In-place lower Cholesky of A + sigma2*I (row-major, lower triangle).  Returns the device info word;
    with check=True a non-zero info raises numpy.linalg.LinAlgError (one host sync)."""
    _req(A, 'A')
    if A.dim() != 2 or A.shape[0] != A.shape[1] or (A.shape[1] > 1 and A.stride(1) != 1):
        raise ValueError('A must be a square row-major matrix')
    n = int(A.shape[0])
    L = _lib.lib()
    nbytes = int(L.gabo_potrf_workspace_bytes(n))
    key = (A.device.index, 'potrf', torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=A.device)
        _ws_cache[key] = ws
    info = torch.zeros((1,), dtype=torch.int32, device=A.device)
    st = L.gabo_potrf_f64(n, A.data_ptr(), _ld(A) if n > 1 else 1, float(sigma2), info.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check('gabo_potrf_f64', st)
    if check:
        k = int(info.item())
        if k != 0:
            raise np.linalg.LinAlgError(f'Cholesky failed: leading minor of order {k} is not positive definite')
    return info


def trsm_(L_: torch.Tensor, B: torch.Tensor, trans: bool = False, winv_all: Optional[torch.Tensor] = None) -> torch.Tensor:
    """B <- L^-1 B (or L^-T B), in place; B is [n, nrhs] row-major.  winv_all: inverses of L's 128 x 128 diagonal blocks
    when already known (from potrf_diag_), which skips the inversion kernel."""
    _req(L_, 'L'); _req(B, 'B')
    if B.dim() == 1:
        B = B.unsqueeze(1)
    n = int(L_.shape[0])
    if B.shape[0] != n or (B.shape[1] > 1 and B.stride(1) != 1):
        raise ValueError('B must be a row-major [n, nrhs] tensor')
    Lh = _lib.lib()
    if winv_all is not None:
        _req(winv_all, 'winv_all')
        key = (L_.device.index, 'trsm_winv', torch.cuda.current_stream().cuda_stream)
        ws = _ws_cache.get(key)
        if ws is None:
            ws = torch.empty(256, dtype=torch.uint8, device=L_.device)
            _ws_cache[key] = ws
        st = Lh.gabo_trsm_winv_f64(1 if trans else 0, n, int(B.shape[1]), L_.data_ptr(), _ld(L_) if n > 1 else 1,
                                   winv_all.data_ptr(), B.data_ptr(),
                                   max(int(B.stride(0)), int(B.shape[1])) if n > 1 else int(B.shape[1]),
                                   ws.data_ptr(), ws.numel(), _stream())
        _lib.check('gabo_trsm_winv_f64', st)
        return B
    nbytes = int(Lh.gabo_trsm_workspace_bytes(n, int(B.shape[1])))
    key = (L_.device.index, 'trsm', torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=L_.device)
        _ws_cache[key] = ws
    st = Lh.gabo_trsm_f64(1 if trans else 0, n, int(B.shape[1]), L_.data_ptr(), _ld(L_) if n > 1 else 1,
                          B.data_ptr(), max(int(B.stride(0)), int(B.shape[1])) if n > 1 else int(B.shape[1]),
                          ws.data_ptr(), ws.numel(), _stream())
    _lib.check('gabo_trsm_f64', st)
    return B


def logdet_half(L_: torch.Tensor) -> torch.Tensor:
    """sum_i log L_ii as a device scalar tensor."""
    _req(L_, 'L')
    n = int(L_.shape[0])
    out = torch.empty((1,), dtype=torch.float64, device=L_.device)
    st = _lib.lib().gabo_logdet_f64(n, L_.data_ptr(), _ld(L_) if n > 1 else 1, out.data_ptr(), _stream())
    _lib.check('gabo_logdet_f64', st)
    return out


def gp_loglik(L_: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    _req(L_, 'L'); _req(alpha, 'alpha')
    n, p = int(alpha.shape[0]), int(alpha.shape[1])
    out = torch.empty((1,), dtype=torch.float64, device=L_.device)
    st = _lib.lib().gabo_gp_loglik_f64(n, p, L_.data_ptr(), _ld(L_) if n > 1 else 1, alpha.data_ptr(),
                                       max(int(alpha.stride(0)), p) if n > 1 else p, out.data_ptr(), _stream())
    _lib.check('gabo_gp_loglik_f64', st)
    return out


def gp_append_row(l: torch.Tensor, kss: float, V: torch.Tensor, ynew: torch.Tensor, mean_const: float, Lrow: torch.Tensor,
                  vnew: torch.Tensor) -> torch.Tensor:
    """Finish a rank-1 growth of the factor on the device (gabo_gp_append_row_f64): l = L^-1 k [n, 1], V [n, p], ynew [p];
    writes Lrow [n + 1] and vnew [p]; returns the device info word (1: extended matrix not positive definite)."""
    _req(l, 'l'); _req(V, 'V'); _req(ynew, 'ynew'); _req(Lrow, 'Lrow'); _req(vnew, 'vnew')
    n, p = int(V.shape[0]), int(V.shape[1])
    info = torch.zeros((1,), dtype=torch.int32, device=V.device)
    st = _lib.lib().gabo_gp_append_row_f64(n, p, l.data_ptr(), max(int(l.stride(0)), 1), float(kss), V.data_ptr(),
                                           max(int(V.stride(0)), p), ynew.data_ptr(), float(mean_const), Lrow.data_ptr(),
                                           vnew.data_ptr(), info.data_ptr(), _stream())
    _lib.check('gabo_gp_append_row_f64', st)
    return info


def gp_grad_beta(K: torch.Tensor, neg_kinv: torch.Tensor, alpha: torch.Tensor, variance: float, beta: float) -> torch.Tensor:
    """d loglik / d beta for a kernel variance * exp(beta g): 1/(2 beta) sum_ij (alpha alpha^T - p Kinv)_ij K_ij log(K_ij/variance),
    from the lower triangles of the noise-free Gram K and of neg_kinv = -(K + sigma^2 I)^-1 (gabo_gp_grad_beta_f64)."""
    _req(K, 'K'); _req(neg_kinv, 'neg_kinv'); _req(alpha, 'alpha')
    n, p = int(alpha.shape[0]), int(alpha.shape[1])
    out = torch.empty((1,), dtype=torch.float64, device=K.device)
    Lh = _lib.lib()
    ws = torch.empty(int(Lh.gabo_gp_grad_beta_workspace_bytes()), dtype=torch.uint8, device=K.device)
    ld = lambda t: max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] > 1 else int(t.shape[1])
    st = Lh.gabo_gp_grad_beta_f64(n, p, K.data_ptr(), ld(K), neg_kinv.data_ptr(), ld(neg_kinv), alpha.data_ptr(), ld(alpha),
                                  float(variance), float(beta), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check('gabo_gp_grad_beta_f64', st)
    return out


GP_SMALL_MAX_N, GP_SMALL_MAX_P = 160, 8     # GABO_GP_SMALL_MAX_N / _P of include/gabo_b200.h


def gp_small_batch(G: torch.Tensor, Y: torch.Tensor, theta: torch.Tensor, want_grad: bool = True,
                   Lout: Optional[torch.Tensor] = None, Vout: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                   info: Optional[torch.Tensor] = None):
    """B fused small-N GP evaluations in one launch (gabo_gp_small_batch_f64): Ky_b = v_b exp(r_b G) + s2_b I.
    G [n, n] log-Gram (log_gram_), Y [n, p], theta [B, 4] = (r, v, s2, c).  Returns (out [B, 5] = loglik, d/dr, d/dv, d/ds2,
    d/dc; info [B]).  Lout [B, n, n] / Vout [B, n, p]: optional factor / whitened residual outputs."""
    G = _req(G, 'G')
    Y = _rowmajor2d(_req(Y, 'Y'))
    theta = _req(theta, 'theta').contiguous()
    if G.dim() != 2 or G.shape[0] != G.shape[1] or (G.shape[1] > 1 and G.stride(1) != 1):
        raise ValueError('G must be a square row-major matrix')
    n, p = int(G.shape[0]), int(Y.shape[1])
    if Y.shape[0] != n or theta.dim() != 2 or theta.shape[1] != 4:
        raise ValueError('Y must be [n, p] and theta [B, 4]')
    B = int(theta.shape[0])
    if out is None:
        out = torch.empty((B, 5), dtype=torch.float64, device=G.device)
    if info is None:
        info = torch.empty((B,), dtype=torch.int32, device=G.device)
    if Lout is not None and (Lout.shape != (B, n, n) or not Lout.is_contiguous()):
        raise ValueError('Lout must be a contiguous [B, n, n] tensor')
    if Vout is not None and (Vout.shape != (B, n, p) or not Vout.is_contiguous()):
        raise ValueError('Vout must be a contiguous [B, n, p] tensor')
    st = _lib.lib().gabo_gp_small_batch_f64(n, p, G.data_ptr(), _ld(G), Y.data_ptr(), _ld(Y), B, theta.data_ptr(), int(bool(want_grad)),
                                            out.data_ptr(), info.data_ptr(), Lout.data_ptr() if Lout is not None else None, n, n * n,
                                            Vout.data_ptr() if Vout is not None else None, _stream())
    _lib.check('gabo_gp_small_batch_f64', st)
    return out, info


def gp_predict(A: torch.Tensor, V: torch.Tensor, kdiag: torch.Tensor, mean_const: float = 0.0):
    """mean[M,p] = A^T V + mean_const, var[M] = kdiag - colsum(A*A) from A = L^-1 K(X,X*) and V = L^-1 (y-m)."""
    _req(A, 'A'); _req(V, 'V'); _req(kdiag, 'kdiag')
    n, M = int(A.shape[0]), int(A.shape[1])
    p = int(V.shape[1])
    mean = torch.empty((M, p), dtype=torch.float64, device=A.device)
    var = torch.empty((M,), dtype=torch.float64, device=A.device)
    st = _lib.lib().gabo_gp_predict_f64(n, M, p, A.data_ptr(), max(int(A.stride(0)), M) if n > 1 else M,
                                        V.data_ptr(), max(int(V.stride(0)), p) if n > 1 else p,
                                        kdiag.data_ptr(), float(mean_const), mean.data_ptr(), var.data_ptr(), _stream())
    _lib.check('gabo_gp_predict_f64', st)
    return mean, var


def gemm_nt_sub_(Cm: torch.Tensor, A: torch.Tensor, B: torch.Tensor, lower: bool = False) -> torch.Tensor:
    """C <- C - A B^T (lower triangle only if lower), all row-major FP64."""
    _req(Cm, 'C'); _req(A, 'A'); _req(B, 'B')
    m, n, k = int(Cm.shape[0]), int(Cm.shape[1]), int(A.shape[1])
    if A.shape[0] != m or B.shape[0] != n or B.shape[1] != k:
        raise ValueError('shape mismatch in gemm_nt_sub_')
    ld = lambda t: max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] > 1 else int(t.shape[1])
    st = _lib.lib().gabo_gemm_nt_sub_f64(m, n, k, A.data_ptr(), ld(A), B.data_ptr(), ld(B), Cm.data_ptr(), ld(Cm),
                                         1 if lower else 0, _stream())
    _lib.check('gabo_gemm_nt_sub_f64', st)
    return Cm


# ---- local building blocks of the multi-GPU factorisation (gaboflow_b200/dist.py) ------------------------------------
def _ldm(t: torch.Tensor) -> int:
    return max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] > 1 else max(int(t.shape[1]), int(t.stride(0)))


def potrf_diag_(A: torch.Tensor, sigma2: float, winv_all: torch.Tensor) -> torch.Tensor:
    """Factor one diagonal block (nb <= 512) in place and write the inverses of its 128 x 128 diagonal blocks into
    winv_all [ceil(nb/128), 128, 128].  Returns the device info word."""
    _req(A, 'A'); _req(winv_all, 'winv_all')
    nb = int(A.shape[0])
    Lh = _lib.lib()
    nbytes = int(Lh.gabo_potrf_workspace_bytes(nb))
    key = (A.device.index, 'potrf_diag', torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=A.device)
        _ws_cache[key] = ws
    info = torch.zeros((1,), dtype=torch.int32, device=A.device)
    st = Lh.gabo_potrf_diag_f64(nb, A.data_ptr(), _ldm(A), float(sigma2), info.data_ptr(), winv_all.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream())
    _lib.check('gabo_potrf_diag_f64', st)
    return info


def trsm_right_(A: torch.Tensor, L_: torch.Tensor, winv_all: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A[m, nb] <- A L^-T in place (L lower nb x nb); winv_all = inverses of L's 128 x 128 diagonal blocks if known."""
    _req(A, 'A'); _req(L_, 'L')
    m, nb = int(A.shape[0]), int(A.shape[1])
    if m == 0 or nb == 0:
        return A
    Lh = _lib.lib()
    nbytes = int(Lh.gabo_trsm_right_workspace_bytes(m, nb))
    key = (A.device.index, 'trsm_right', torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=A.device)
        _ws_cache[key] = ws
    if winv_all is not None:
        st = Lh.gabo_trsm_right_winv_f64(m, nb, L_.data_ptr(), _ldm(L_), winv_all.data_ptr(), A.data_ptr(), _ldm(A),
                                         ws.data_ptr(), ws.numel(), _stream())
        _lib.check('gabo_trsm_right_winv_f64', st)
    else:
        st = Lh.gabo_trsm_right_f64(m, nb, L_.data_ptr(), _ldm(L_), A.data_ptr(), _ldm(A), ws.data_ptr(), ws.numel(), _stream())
        _lib.check('gabo_trsm_right_f64', st)
    return A


def trsm_right_scatter_(A: torch.Tensor, L_: torch.Tensor, winv_all: torch.Tensor, dst_ptrs, ldp: int, p_row_off: int,
                        cyc_nb: int, cyc_world: int) -> torch.Tensor:
    """trsm_right_ that also stores the solved block column, in global row order, into the block-column buffers at the raw
    device pointers ``dst_ptrs`` (own + IPC-mapped peers); see gabo_trsm_right_scatter_f64."""
    import ctypes as C
    _req(A, 'A'); _req(L_, 'L'); _req(winv_all, 'winv_all')
    m, nb = int(A.shape[0]), int(A.shape[1])
    if m == 0 or nb == 0:
        return A
    Lh = _lib.lib()
    nbytes = int(Lh.gabo_trsm_right_workspace_bytes(m, nb))
    key = (A.device.index, 'trsm_right', torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=A.device)
        _ws_cache[key] = ws
    arr = (C.c_void_p * len(dst_ptrs))(*[int(q) for q in dst_ptrs])
    st = Lh.gabo_trsm_right_scatter_f64(m, nb, L_.data_ptr(), _ldm(L_), winv_all.data_ptr(), A.data_ptr(), _ldm(A),
                                        C.cast(arr, C.c_void_p), len(dst_ptrs), int(ldp), int(p_row_off), int(cyc_nb),
                                        int(cyc_world), ws.data_ptr(), ws.numel(), _stream())
    _lib.check('gabo_trsm_right_scatter_f64', st)
    return A


def flag_signal(flag_ptrs, value: int) -> None:
    """Stream-ordered store of ``value`` into the 32-bit flags at the raw device pointers ``flag_ptrs`` (peers' memory)."""
    import ctypes as C
    if not flag_ptrs:
        return
    arr = (C.c_void_p * len(flag_ptrs))(*[int(q) for q in flag_ptrs])
    _lib.check('gabo_flag_signal', _lib.lib().gabo_flag_signal(C.cast(arr, C.c_void_p), len(flag_ptrs), int(value), _stream()))


def flag_wait_geq(flag_ptr: int, value: int) -> None:
    """The current stream stalls (no SM occupied) until the 32-bit flag at ``flag_ptr`` (local memory) is >= value."""
    _lib.check('gabo_flag_wait_geq', _lib.lib().gabo_flag_wait_geq(int(flag_ptr), int(value), _stream()))


def solve_cyclic_p2p_(A_loc: torch.Tensor, B_loc: torch.Tensor, n: int, nb: int, world: int, rank: int, winv_ptrs, xarea_ptrs,
                      flag_ptrs, base: int) -> torch.Tensor:
    """Distributed forward substitution in one host call (gabo_solve_cyclic_p2p_f64).  winv_ptrs: per panel, the device
    pointer of the block inverses (0 for panels this rank does not own); xarea_ptrs / flag_ptrs: per rank."""
    import ctypes as C
    _req(A_loc, 'A_loc'); _req(B_loc, 'B_loc')
    key = (A_loc.device.index, 'solve_cyclic', torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty(256, dtype=torch.uint8, device=A_loc.device)
        _ws_cache[key] = ws
    wv = (C.c_void_p * len(winv_ptrs))(*[C.c_void_p(int(q)) if q else None for q in winv_ptrs])
    xa = (C.c_void_p * world)(*[int(q) for q in xarea_ptrs])
    fl = (C.c_void_p * world)(*[int(q) for q in flag_ptrs])
    st = _lib.lib().gabo_solve_cyclic_p2p_f64(int(n), int(nb), int(world), int(rank), A_loc.data_ptr(), _ldm(A_loc), B_loc.data_ptr(),
                                              max(int(B_loc.stride(0)), int(B_loc.shape[1])), int(B_loc.shape[1]),
                                              C.cast(wv, C.c_void_p), C.cast(xa, C.c_void_p), C.cast(fl, C.c_void_p), int(base),
                                              ws.data_ptr(), ws.numel(), _stream())
    _lib.check('gabo_solve_cyclic_p2p_f64', st)
    return B_loc


def push_signal(src: torch.Tensor, dst_ptrs, flag_ptrs, value: int) -> None:
    """Copies the contiguous FP64 tensor ``src`` into each raw device pointer of ``dst_ptrs`` (peers' memory) and then raises
    the 32-bit flags at ``flag_ptrs`` to ``value`` -- one launch (gabo_push_signal_f64)."""
    import ctypes as C
    if not dst_ptrs:
        return
    _req(src, 'src')
    if not src.is_contiguous():
        raise ValueError('push_signal needs a contiguous source')
    d = (C.c_void_p * len(dst_ptrs))(*[int(q) for q in dst_ptrs])
    f = (C.c_void_p * len(flag_ptrs))(*[int(q) for q in flag_ptrs])
    _lib.check('gabo_push_signal_f64', _lib.lib().gabo_push_signal_f64(src.data_ptr(), int(src.numel()), C.cast(d, C.c_void_p),
                                                                     C.cast(f, C.c_void_p), len(dst_ptrs), int(value), _stream()))


def copy_async(dst_ptr: int, src_ptr: int, nbytes: int) -> None:
    """cudaMemcpyAsync on the current stream between raw device pointers (local or IPC-mapped peer memory)."""
    _lib.check('gabo_copy_async', _lib.lib().gabo_copy_async(int(dst_ptr), int(src_ptr), int(nbytes), _stream()))


def syrk_cyclic_(Cm: torch.Tensor, A: torch.Tensor, B: torch.Tensor, row_base: int, col_base: int, cyc_nb: int = 0,
                 cyc_world: int = 1, cyc_rank: int = 0) -> torch.Tensor:
    """C[m,n] -= A[m,k] B[n,k]^T where global col <= global row (see gabo_syrk_cyclic_f64)."""
    _req(Cm, 'C'); _req(A, 'A'); _req(B, 'B')
    m, n, k = int(Cm.shape[0]), int(Cm.shape[1]), int(A.shape[1])
    if m == 0 or n == 0:
        return Cm
    st = _lib.lib().gabo_syrk_cyclic_f64(m, n, k, A.data_ptr(), _ldm(A), B.data_ptr(), _ldm(B), Cm.data_ptr(), _ldm(Cm),
                                         int(row_base), int(col_base), int(cyc_nb), int(cyc_world), int(cyc_rank), _stream())
    _lib.check('gabo_syrk_cyclic_f64', st)
    return Cm


def solve_update_(Y: torch.Tensor, Lp: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """Y[m,nrhs] -= Lp[m,k] X[k,nrhs]."""
    _req(Y, 'Y'); _req(Lp, 'Lp'); _req(X, 'X')
    m, k, nrhs = int(Lp.shape[0]), int(Lp.shape[1]), int(X.shape[1])
    if m == 0 or k == 0 or nrhs == 0:
        return Y
    st = _lib.lib().gabo_solve_update_f64(m, k, nrhs, Lp.data_ptr(), _ldm(Lp), X.data_ptr(), _ldm(X), Y.data_ptr(), _ldm(Y), _stream())
    _lib.check('gabo_solve_update_f64', st)
    return Y


def points_gram_cyclic(kind: int, X: torch.Tensor, bases, ldk: int, world: int, rank: int, beta: float = 1.0,
                       variance: float = 1.0, workspace: Optional[torch.Tensor] = None, recompute_frac256: int = 0) -> None:
    """Block-cyclic symmetric Gram with peer-memory mirror stores (gabo_points_gram_cyclic_hybrid_f64).  `bases`: list of
    `world` device pointers (ints), entry r = base of rank r's local buffer as mapped in this process.  recompute_frac256 = f:
    f/256 of the panel pairs owned by different ranks are recomputed by the row owner instead of mirrored over NVLink."""
    import ctypes as C
    L = _lib.lib()
    X = _rowmajor2d(_req(X, 'X'))
    n, dim = int(X.shape[0]), int(X.shape[1])
    nbytes = int(L.gabo_points_gram_cyclic_workspace_bytes(n, dim))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=X.device)
    arr = (C.c_void_p * world)(*[C.c_void_p(int(b)) for b in bases])
    st = L.gabo_points_gram_cyclic_hybrid_f64(int(kind), X.data_ptr(), n, _ld(X), dim, float(beta), float(variance),
                                              C.cast(arr, C.c_void_p), int(ldk), int(world), int(rank), int(recompute_frac256),
                                              workspace.data_ptr(), workspace.numel(), _stream())
    _lib.check('gabo_points_gram_cyclic_hybrid_f64', st)


def points_gram_cyclic_rows_i8(kind: int, X: torch.Tensor, K_local: torch.Tensor, world: int, rank: int, beta: float = 1.0,
                               variance: float = 1.0, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Multi-GPU Gram with no exchange at all: this rank computes the FULL rows of its 512-row panels (block-cyclic) into its
    local buffer with the int8-tensor-core engine (gabo_points_gram_i8_f64, cyc_world > 0)."""
    L = _lib.lib()
    X = _rowmajor2d(_req(X, 'X'))
    _req(K_local, 'K_local')
    n, dim = int(X.shape[0]), int(X.shape[1])
    nbytes = int(L.gabo_points_gram_i8_workspace_bytes(n, n))
    if workspace is None or workspace.numel() < nbytes or workspace.data_ptr() % 128:
        workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=X.device)
    st = L.gabo_points_gram_i8_f64(int(kind), X.data_ptr(), n, _ld(X), None, n, 0, dim, float(beta), float(variance),
                                   K_local.data_ptr(), int(K_local.stride(0)), 0, n, FULL, int(world), int(rank),
                                   workspace.data_ptr(), workspace.numel(), _stream())
    _lib.check('gabo_points_gram_i8_f64', st)
    return K_local


class StagedGramState:
    """Device scratch of the staged block-cyclic Gram (gabo_points_gram_cyclic_staged_f64), allocated once per (n, world, rank,
    share, split): the local staging buffer for the mirror blocks bound for other ranks, the per-chunk progress words and the
    epoch counter."""

    def __init__(self, n: int, world: int, rank: int, recompute_frac256: int, split: int, device, direct_frac256: int = 0):
        L = _lib.lib()
        self.key = (int(n), int(world), int(rank), int(recompute_frac256), int(direct_frac256), int(split))
        nbytes = int(L.gabo_points_gram_cyclic_stage_bytes(*self.key))
        self.stage = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        npanels = n // 512
        nslots = (npanels - 1 - rank) // world + 1 if npanels > rank else 0
        self.progress = torch.zeros(2 * (nslots + 3 * split) + 8, dtype=torch.int32, device=device)
        self.epoch = 0


def points_gram_cyclic_staged(kind: int, X: torch.Tensor, bases, ldk: int, world: int, rank: int, state: StagedGramState,
                              beta: float = 1.0, variance: float = 1.0, workspace: Optional[torch.Tensor] = None) -> None:
    """Block-cyclic symmetric Gram whose cross-GPU mirror blocks travel by copy engine from a local staging buffer
    (gabo_points_gram_cyclic_staged_f64); same result, bit for bit, as points_gram_cyclic."""
    import ctypes as C
    L = _lib.lib()
    X = _rowmajor2d(_req(X, 'X'))
    n, dim = int(X.shape[0]), int(X.shape[1])
    if state.key[:3] != (n, int(world), int(rank)):
        raise ValueError('StagedGramState was built for another (n, world, rank)')
    nbytes = int(L.gabo_points_gram_cyclic_workspace_bytes(n, dim))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=X.device)
    arr = (C.c_void_p * world)(*[C.c_void_p(int(b)) for b in bases])
    state.epoch += 1
    st = L.gabo_points_gram_cyclic_staged_f64(int(kind), X.data_ptr(), n, _ld(X), dim, float(beta), float(variance),
                                              C.cast(arr, C.c_void_p), int(ldk), int(world), int(rank), state.key[3], state.key[4],
                                              state.key[5], state.stage.data_ptr(), state.stage.numel(), state.progress.data_ptr(),
                                              state.progress.numel(), state.epoch, workspace.data_ptr(), workspace.numel(), _stream())
    _lib.check('gabo_points_gram_cyclic_staged_f64', st)


def sphere_posterior_grad(kind: int, X: torch.Tensor, Kx: torch.Tensor, alpha: torch.Tensor, Bm: torch.Tensor, beta: float,
                          variance: float, mean: Optional[torch.Tensor] = None, var: Optional[torch.Tensor] = None,
                          fmin: Optional[float] = None):
    """Euclidean gradients w.r.t. M candidate points on the sphere (gabo_sphere_posterior_grad_f64): returns
    (d mean/dx* [M,D], d var/dx* [M,D], d EI/dx* [M,D] or None).  Kx = K(X, X*) [N,M], alpha = Ky^-1 (y-m) [N],
    Bm = Ky^-1 Kx [N,M]; the EI gradient needs the posterior mean / var [M] and fmin."""
    X = _rowmajor2d(_req(X, 'X'))
    _req(Kx, 'Kx'); _req(alpha, 'alpha'); _req(Bm, 'Bm')
    n, dim = int(X.shape[0]), int(X.shape[1])
    M = int(Kx.shape[1])
    al = alpha.reshape(-1).contiguous()
    if al.numel() != n or Kx.shape[0] != n or tuple(Bm.shape) != (n, M):
        raise ValueError('shape mismatch in sphere_posterior_grad')
    gmean = torch.empty((M, dim), dtype=torch.float64, device=X.device)
    gvar = torch.empty((M, dim), dtype=torch.float64, device=X.device)
    dei = None
    mptr = vptr = dptr = None
    if fmin is not None:
        mean_c = _req(mean, 'mean').reshape(-1).contiguous()
        var_c = _req(var, 'var').reshape(-1).contiguous()
        dei = torch.empty((M, dim), dtype=torch.float64, device=X.device)
        mptr, vptr, dptr = mean_c.data_ptr(), var_c.data_ptr(), dei.data_ptr()
    ld = lambda t: max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] > 1 else int(t.shape[1])
    st = _lib.lib().gabo_sphere_posterior_grad_f64(int(kind), n, M, dim, X.data_ptr(), _ld(X), Kx.data_ptr(), ld(Kx),
                                                   al.data_ptr(), Bm.data_ptr(), ld(Bm), float(beta), float(variance),
                                                   mptr, vptr, float(fmin) if fmin is not None else 0.0,
                                                   gmean.data_ptr(), gvar.data_ptr(), dptr, _stream())
    _lib.check('gabo_sphere_posterior_grad_f64', st)
    return gmean, gvar, dei


def spd_ai_posterior_grad(kind: int, p_train: SpdPrep, p_cand: SpdPrep, alpha: torch.Tensor, Bm: torch.Tensor, beta: float,
                          variance: float, mean: Optional[torch.Tensor] = None, var: Optional[torch.Tensor] = None,
                          fmin: Optional[float] = None):
    """Euclidean gradients, in Mandel coordinates, w.r.t. M candidate SPD matrices under the affine-invariant kernels
    (gabo_spd_ai_posterior_grad_f64): (d mean/dx* [M,m], d var/dx* [M,m], d EI/dx* [M,m] or None).  alpha = Ky^-1 (y-m) [N],
    Bm = Ky^-1 K(X, X*) [N,M]."""
    _req(alpha, 'alpha'); _req(Bm, 'Bm')
    n, M, d = p_train.n, p_cand.n, p_train.d
    mm = d * (d + 1) // 2
    al = alpha.reshape(-1).contiguous()
    if al.numel() != n or tuple(Bm.shape) != (n, M) or p_cand.d != d or p_train.W is None or p_cand.S is None:
        raise ValueError('shape mismatch in spd_ai_posterior_grad')
    dev = p_train.info.device
    gmean = torch.empty((M, mm), dtype=torch.float64, device=dev)
    gvar = torch.empty((M, mm), dtype=torch.float64, device=dev)
    dei = None
    mptr = vptr = dptr = None
    if fmin is not None:
        mean_c = _req(mean, 'mean').reshape(-1).contiguous()
        var_c = _req(var, 'var').reshape(-1).contiguous()
        dei = torch.empty((M, mm), dtype=torch.float64, device=dev)
        mptr, vptr, dptr = mean_c.data_ptr(), var_c.data_ptr(), dei.data_ptr()
    st = _lib.lib().gabo_spd_ai_posterior_grad_f64(int(kind), d, n, M, p_train.W.data_ptr(), p_cand.S.data_ptr(), al.data_ptr(),
                                                   Bm.data_ptr(), max(int(Bm.stride(0)), M) if n > 1 else M, float(beta),
                                                   float(variance), mptr, vptr, float(fmin) if fmin is not None else 0.0,
                                                   gmean.data_ptr(), gvar.data_ptr(), dptr, _stream())
    _lib.check('gabo_spd_ai_posterior_grad_f64', st)
    return gmean, gvar, dei


def sphere_acq_fused(kind: int, X: torch.Tensor, Xc: torch.Tensor, Linv: torch.Tensor, V: torch.Tensor, alpha: Optional[torch.Tensor],
                     beta: float, variance: float, mean_const: float, fmin: float, with_grad: bool = False, out=None):
    """One launch per candidate batch (gabo_sphere_acq_fused_f64): K(X,x*) -> L^-1 -> mean / var -> EI (-> gradients).
    Returns (mean [M], var [M], ei [M], dmean, dvar, dei) with the last three None unless with_grad.  ``out``: optional dict of
    preallocated outputs (keys mean, var, ei, dmean, dvar, dei) so that repeated evaluations allocate nothing."""
    X = _rowmajor2d(_req(X, 'X')); Xc = _rowmajor2d(_req(Xc, 'Xc'))
    _req(Linv, 'Linv'); _req(V, 'V')
    n, dim, M = int(X.shape[0]), int(X.shape[1]), int(Xc.shape[0])
    dev = X.device
    o = out if out is not None else {}

    def buf(name, shape):
        t = o.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float64, device=dev)
            o[name] = t
        return t

    mean, var, ei = buf('mean', (M,)), buf('var', (M,)), buf('ei', (M,))
    gm = gv = ge = None
    if with_grad:
        gm, gv, ge = buf('dmean', (M, dim)), buf('dvar', (M, dim)), buf('dei', (M, dim))
        _req(alpha, 'alpha')
    ptr = lambda t: t.data_ptr() if t is not None else None
    st = _lib.lib().gabo_sphere_acq_fused_f64(int(kind), n, M, dim, X.data_ptr(), _ld(X), Xc.data_ptr(), _ld(Xc), Linv.data_ptr(),
                                              max(int(Linv.stride(0)), n) if n > 1 else 1, V.data_ptr(), ptr(alpha), float(beta),
                                              float(variance), float(mean_const), float(fmin), mean.data_ptr(), var.data_ptr(),
                                              ei.data_ptr(), ptr(gm), ptr(gv), ptr(ge), _stream())
    _lib.check('gabo_sphere_acq_fused_f64', st)
    return mean, var, ei, gm, gv, ge


def expected_improvement(mean: torch.Tensor, var: torch.Tensor, fmin: float) -> torch.Tensor:
    """EI over the candidates (gabo_expected_improvement_f64); mean / var are [M] or [M,1] device tensors."""
    _req(mean, 'mean'); _req(var, 'var')
    m = mean.reshape(-1).contiguous()
    v = var.reshape(-1).contiguous()
    out = torch.empty_like(m)
    st = _lib.lib().gabo_expected_improvement_f64(int(m.numel()), m.data_ptr(), v.data_ptr(), float(fmin), out.data_ptr(), _stream())
    _lib.check('gabo_expected_improvement_f64', st)
    return out.reshape(mean.shape)


def math_probe(which: int, x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Diagnostic (gabo_math_probe_f64): the lean device acos (0) / scale*exp for x <= 0 (1) / log of positive normals (2)
    that the Gram epilogues inline, evaluated element-wise on a 1-D FP64 tensor."""
    _req(x, 'x')
    xc = x.reshape(-1).contiguous()
    out = torch.empty_like(xc)
    st = _lib.lib().gabo_math_probe_f64(int(which), xc.data_ptr(), int(xc.numel()), float(scale), out.data_ptr(), _stream())
    _lib.check('gabo_math_probe_f64', st)
    return out.reshape(x.shape)
