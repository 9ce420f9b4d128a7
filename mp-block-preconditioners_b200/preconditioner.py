"""Host-side mirror of the reference's `preconditioner.py`: same class name, constructor and factory
methods, but the matrices are matrix-free operator objects backed by the CUDA plan (libmpbp.so).

Reference: /root/reference/preconditioner.py (`MultiphaseBlockPreconditioner`, :17-349) and the setup
part of `solve_with_approx_schur_pc` (/root/reference/solve.py:243-254, :280-281).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from ._cabi import SUB_JACOBI, SUB_MG, Config, check

PI = np.pi


def thn(y, x):
    """Network volume fraction (reference preconditioner.py:9-11)."""
    return 0.25 * np.sin(2 * PI * x) * np.sin(2 * PI * y) + 0.5


def ths(y, x):
    """Solvent volume fraction (reference preconditioner.py:13-15)."""
    return 1.0 - thn(y, x)


@dataclass
class SubSolver:
    """What fills the two `ilupp.ILUTPreconditioner` slots of solve.py:251/:254.

    kind 'mg': `cycles` V(nu1,nu2) cycles with damped-Jacobi smoothing ("Multigrid PC with Jacobi
    smoother", solve.py:266/:274), optionally Chebyshev-accelerated; kind 'jacobi': `sweeps` damped
    sweeps of solve.Jacobi (solve.py:149-159) from a zero guess.
    """
    kind: str = "mg"
    F_cycles: int = 4
    P_cycles: int = 2
    F_sweeps: int = 20
    P_sweeps: int = 20
    omega: float = 0.8
    nu1: int = 2
    nu2: int = 2
    n_coarse: int = 4
    cheb: bool = True
    lmin: float = 0.75
    lmax: float = 1.2
    project: bool = True
    dist_min_n: int = 0  # multi-GPU: levels with n < dist_min_n are replicated (0 = library default 512)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Plan:
    """Owns one mpbp_plan handle and the torch tensor that backs its device workspace."""

    def __init__(self, n, xi, eta_n, eta_s, c, d_u, d_p=1.0, d_div=-1.0, sub: SubSolver | None = None, theta=None,
                 device=None, rank=0, nranks=1, nccl_id: bytes | None = None, operators_only=False):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("mp-block-preconditioners_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        sub = sub or SubSolver()
        self.sub = sub
        cfg = Config()
        check(self.lib.mpbp_config_default(C.byref(cfg)))
        cfg.n, cfg.xi, cfg.eta_n, cfg.eta_s = int(n), float(xi), float(eta_n), float(eta_s)
        cfg.c, cfg.d_u, cfg.d_p, cfg.d_div = float(c), float(d_u), float(d_p), float(d_div)
        cfg.rank, cfg.nranks = int(rank), int(nranks)
        kind = {"mg": SUB_MG, "jacobi": SUB_JACOBI}[sub.kind]
        cfg.F_kind = cfg.P_kind = kind
        cfg.F_sweeps, cfg.P_sweeps = sub.F_sweeps, sub.P_sweeps
        cfg.F_cycles, cfg.P_cycles = sub.F_cycles, sub.P_cycles
        cfg.omega, cfg.nu1, cfg.nu2, cfg.n_coarse = sub.omega, sub.nu1, sub.nu2, sub.n_coarse
        cfg.cheb, cfg.lmin, cfg.lmax, cfg.project = int(sub.cheb), sub.lmin, sub.lmax, int(sub.project)
        cfg.operators_only = int(operators_only)
        cfg.dist_min_n = int(sub.dist_min_n)
        self._keep = []
        if theta is not None:
            th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(n, n))
            self._keep.append(th)
            cfg.theta_host = th.ctypes.data_as(C.c_void_p)
        if nranks > 1:
            idbuf = C.create_string_buffer(nccl_id, 128)
            self._keep.append(idbuf)
            cfg.nccl_unique_id = C.cast(idbuf, C.c_void_p)
        need = C.c_size_t()
        check(self.lib.mpbp_plan_workspace_bytes(C.byref(cfg), C.byref(need)))
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(need.value + 512, dtype=torch.uint8, device=self.device)
            cfg.workspace = C.c_void_p(self.workspace.data_ptr())
            cfg.workspace_bytes = self.workspace.numel()
            h = C.c_void_p()
            check(self.lib.mpbp_plan_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self.cfg = cfg
        self.n = int(n)
        self.rank, self.nranks = rank, nranks
        self.rows = self.lib.mpbp_plan_rows_local(h)
        self.N = self.rows * self.n  # local cells
        self.kry_ws = None

    def close(self):
        """Destroy the native plan now (idempotent)."""
        h = getattr(self, "h", None)
        if h is not None and h.value:
            self.h = None
            self.lib.mpbp_plan_destroy(h)

    def __del__(self):
        try:
            import sys
            if sys.is_finalizing():
                return  # never tear down CUDA/NCCL state from interpreter shutdown
            self.close()
        except BaseException:
            pass

    # ---- helpers -------------------------------------------------------------------------
    def stream(self):
        return _stream_ptr(self.device)

    def _to_dev(self, x, length):
        """numpy / torch in -> (contiguous CUDA float64 tensor, was_numpy)."""
        if isinstance(x, torch.Tensor):
            if x.device.type != "cuda" or x.dtype != torch.float64:
                x = x.to(device=self.device, dtype=torch.float64)
            t = x.contiguous().reshape(-1)
            was_np = False
        else:
            a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
            t = torch.from_numpy(a).to(self.device)
            was_np = True
        if t.numel() != length:
            raise ValueError(f"dimension mismatch: expected a vector of length {length}, got {t.numel()}")
        return t, was_np

    def _out(self, t, was_np):
        return t.cpu().numpy() if was_np else t

    def call(self, name, x, in_len, out_len):
        t, was_np = self._to_dev(x, in_len)
        y = torch.empty(out_len, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(getattr(self.lib, name)(self.h, t.data_ptr(), y.data_ptr(), self.stream()))
        return self._out(y, was_np)

    @property
    def launches(self):
        return self.lib.mpbp_plan_launches(self.h)

    def precond_bytes(self):
        b = C.c_double()
        check(self.lib.mpbp_precond_bytes(self.h, C.byref(b)))
        return b.value


class _Operator:
    """Matrix-free stand-in for one of the reference's dense ndarrays: has .shape, .dtype, `@`,
    .matvec and .dot over numpy arrays (numpy out) or torch CUDA tensors (torch out)."""

    dtype = np.dtype(np.float64)
    _fn = ""

    def __init__(self, plan: Plan, rows_f: int, cols_f: int):
        self.plan = plan
        self._rf, self._cf = rows_f, cols_f
        self.shape = (rows_f * plan.n * plan.n, cols_f * plan.n * plan.n)

    def matvec(self, x):
        p = self.plan
        return p.call(self._fn, x, self._cf * p.N, self._rf * p.N)

    __matmul__ = matvec
    dot = matvec
    __call__ = matvec

    def toarray(self):
        """Dense matrix by applying the operator to the identity (small n only; for tests)."""
        m = self.shape[1]
        if m > 8192:
            raise MemoryError("toarray() is for small test problems")
        eye = torch.eye(m, dtype=torch.float64, device=self.plan.device)
        cols = [self.matvec(eye[j]) for j in range(m)]
        return torch.stack(cols, dim=1).cpu().numpy()


class SystemOperator(_Operator):
    """A = [[F, G], [d_div*D, 0]] (preconditioner.py:339-341); `A @ x` of solve.py:166, apply.py:72."""
    _fn = "mpbp_apply_A"

    def __init__(self, plan):
        super().__init__(plan, 5, 5)


class VelocityOperator(_Operator):
    """F = XI_block + d_u*L (preconditioner.py:337)."""
    _fn = "mpbp_apply_F"

    def __init__(self, plan):
        super().__init__(plan, 4, 4)


class GradientOperator(_Operator):
    """G = d_p*[G_n; G_s] (preconditioner.py:313); np.matmul(G, x_p) of solve.py:273."""
    _fn = "mpbp_apply_G"

    def __init__(self, plan):
        super().__init__(plan, 4, 1)


class DivergenceOperator(_Operator):
    """D = [D_n, D_s], un-negated as returned at preconditioner.py:349; np.matmul(D, .) of solve.py:259."""

    def __init__(self, plan):
        super().__init__(plan, 1, 4)

    def matvec(self, x):
        p = self.plan
        t, was_np = p._to_dev(x, 4 * p.N)
        y = torch.empty(p.N, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            check(p.lib.mpbp_apply_D(p.h, t.data_ptr(), None, y.data_ptr(), p.stream()))
        return p._out(y, was_np)

    __matmul__ = matvec
    dot = matvec
    __call__ = matvec


class GtGOperator(_Operator):
    """Gt_G = (-D) G (solve.py:246-247): 5-point variable-coefficient periodic Laplacian."""
    _fn = "mpbp_apply_GtG"

    def __init__(self, plan):
        super().__init__(plan, 1, 1)


class GtFGOperator(_Operator):
    """Gt_F_G = (-D) F G (solve.py:248-249)."""
    _fn = "mpbp_apply_GtFG"

    def __init__(self, plan):
        super().__init__(plan, 1, 1)


class ExactSchurOperator:
    """The dense exact Schur complement S = -D F^-1 G of preconditioner.py:343-346, for SMALL grids only: it is the
    reference's O(n^6) research cross-check (solve_with_exact_schur_pc, solve.py:210-238), kept as a verification path.
    F, D, G are obtained by applying the GPU operators to the identity, the dense algebra (scipy.linalg.inv, as the
    reference) runs on the host, lazily.  Above n = MAX_N only .shape exists."""

    MAX_N = 24

    def __init__(self, plan):
        self.plan = plan
        self.shape = (plan.n * plan.n, plan.n * plan.n)
        self.dtype = np.dtype(np.float64)
        self._S = None

    def toarray(self):
        if self._S is None:
            p = self.plan
            if p.n > self.MAX_N or p.nranks > 1:
                raise NotImplementedError(f"the dense exact Schur complement is a small-n verification path (n <= {self.MAX_N}, "
                                          "one GPU); it is O(n^6) and out of scope of the hot path (SURVEY.md 2, 8b)")
            import scipy.linalg
            F = VelocityOperator(p).toarray()
            D = DivergenceOperator(p).toarray()
            G = GradientOperator(p).toarray()
            self._S = -1.0 * np.matmul(np.matmul(D, scipy.linalg.inv(F)), G)  # preconditioner.py:344-346
        return self._S

    def matvec(self, x):
        return self.toarray() @ np.asarray(x, dtype=np.float64)

    __matmul__ = matvec
    dot = matvec


ExactSchurUnsupported = ExactSchurOperator  # former name (round 1 placeholder)


class ApproxSolve(_Operator):
    """F~^-1 or (GtG)~^-1 as configured: the object solve.py:251/:254 gets from ilupp, used with `@`."""

    def __init__(self, plan, which):
        super().__init__(plan, 4 if which == "F" else 1, 4 if which == "F" else 1)
        self._fn = "mpbp_solve_F" if which == "F" else "mpbp_solve_P"


class ApproxSchurOperator(_Operator):
    """`approx_schur = LinearOperator(shape=(m,m), matvec=approx_schur_op)` (solve.py:257-281)."""
    _fn = "mpbp_precond_apply"

    def __init__(self, plan):
        super().__init__(plan, 5, 5)

    def matvec_host(self, v: np.ndarray) -> np.ndarray:
        """Host buffer in, host buffer out, copies inside the C call (the e2e path of bench.py)."""
        p = self.plan
        v = np.ascontiguousarray(v, dtype=np.float64)
        z = np.empty_like(v)
        with torch.cuda.device(p.device):
            check(p.lib.mpbp_precond_apply_host(p.h, v.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                                                p.stream()))
        return z


class MultiphaseBlockPreconditioner:
    """Drop-in for reference preconditioner.py:17-349 on a B200.

    Same constructor and factory names; the returned "matrices" are operator objects.  Extra keyword
    arguments choose the device, the approximate sub-solvers and (multi-GPU) the process group.
    """

    def __init__(self, n, xi, eta_n, eta_s, *, sub_solver: SubSolver | None = None, theta=None, device=None,
                 distributed: bool = False):
        self.n = n
        self.dx = 1 / n
        self.dy = 1 / n
        self.xi = xi
        self.eta_n = eta_n
        self.eta_s = eta_s
        self.sub_solver = sub_solver or SubSolver()
        self.theta = theta
        self.device = device
        self.distributed = distributed
        self._plans = {}

    # -- plan cache: the reference rebuilds its matrices on every call (solve.py:46-47, :243-244) --
    def plan(self, c, d_u, d_p=1.0, d_div=-1.0, *, xi=None, eta_n=None, eta_s=None, operators_only=False) -> Plan:
        key = (float(c), float(d_u), float(d_p), float(d_div), xi, eta_n, eta_s, operators_only)
        if key not in self._plans:
            rank, nranks, nid = 0, 1, None
            if self.distributed:
                import torch.distributed as dist
                rank, nranks = dist.get_rank(), dist.get_world_size()
                if nranks > 1:
                    buf = torch.zeros(128, dtype=torch.uint8)
                    if rank == 0:
                        raw = C.create_string_buffer(128)
                        check(_cabi.load().mpbp_nccl_unique_id(raw))
                        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
                    if dist.get_backend() == "nccl":
                        dbuf = buf.cuda()
                        dist.broadcast(dbuf, src=0)
                        buf = dbuf.cpu()
                    else:
                        dist.broadcast(buf, src=0)
                    nid = bytes(buf.numpy().tobytes())
            self._plans[key] = Plan(self.n, self.xi if xi is None else xi, self.eta_n if eta_n is None else eta_n,
                                    self.eta_s if eta_s is None else eta_s, c, d_u, d_p, d_div, sub=self.sub_solver,
                                    theta=self.theta, device=self.device, rank=rank, nranks=nranks, nccl_id=nid,
                                    operators_only=operators_only)
        return self._plans[key]

    def close(self):
        """Destroy all native plans of this object now."""
        for pl in self._plans.values():
            pl.close()
        self._plans.clear()

    def get_big_A_matrix(self, c, d_u, d_p: float = 1.0, d_div: float = -1.0):
        """(A, S, F, D, G) as at preconditioner.py:299-349; S is the lazily built dense exact Schur complement
        (small n only, see ExactSchurOperator)."""
        p = self.plan(c, d_u, d_p, d_div)
        return SystemOperator(p), ExactSchurOperator(p), VelocityOperator(p), DivergenceOperator(p), GradientOperator(p)

    def get_thn_vals(self, n, row_on_grid, col_on_grid, is_ths):
        """preconditioner.py:26-84: the six cell-centred volume fractions around the u-face of cell
        (row_on_grid, col_on_grid), periodic: (i,j), (i+1,j), (i,j+1), (i+1,j+1), (i,j-1), (i+1,j-1) in the reference's
        naming, i.e. columns col-1 / col and rows row, row-1, row+1.  (The CUDA kernels recompute these averages in
        registers from the cell field; this method exists for callers of the reference's API.)"""
        dx, dy = self.dx, self.dy
        cw, ce = (col_on_grid - 1) % n, col_on_grid % n
        rn, rs = (row_on_grid - 1) % n, (row_on_grid + 1) % n
        cell = lambda r, c: thn(-(r + 0.5) * dy, (c + 0.5) * dx)
        vals = (cell(row_on_grid, cw), cell(row_on_grid, ce), cell(rn, cw), cell(rn, ce), cell(rs, cw), cell(rs, ce))
        return tuple(1.0 - v for v in vals) if is_ths else vals

    def get_block_matrices(self, is_ths):
        """(L, D, XI, G) of one phase (preconditioner.py:86-297) as operator objects on 2N / N vectors."""
        return _phase_blocks(self, bool(is_ths))

    def approx_schur_operator(self, c, d_u, d_p: float = 1.0, d_div: float = -1.0) -> ApproxSchurOperator:
        """The `approx_schur` LinearOperator of solve.py:280-281 for these coefficients."""
        return ApproxSchurOperator(self.plan(c, d_u, d_p, d_div))

    def derived_operators(self, c, d_u, d_p: float = 1.0, d_div: float = -1.0):
        """(Gt_G, Gt_F_G, F_inv, Gt_G_factorization) of solve.py:246-254."""
        p = self.plan(c, d_u, d_p, d_div)
        return GtGOperator(p), GtFGOperator(p), ApproxSolve(p, "F"), ApproxSolve(p, "P")


class _PhaseOp:
    """One single-phase block (L, D, XI or G of preconditioner.py:86-297) evaluated through the
    two-phase kernels by zero-padding the other phase."""

    dtype = np.dtype(np.float64)

    def __init__(self, plan, kind, is_ths):
        self.plan, self.kind, self.is_ths = plan, kind, is_ths
        N = plan.n * plan.n
        self.shape = {"L": (2 * N, 2 * N), "XI": (2 * N, 2 * N), "D": (N, 2 * N), "G": (2 * N, N)}[kind]

    def matvec(self, x):
        p, N = self.plan, self.plan.N
        lo = 2 * N if self.is_ths else 0
        was_np = not isinstance(x, torch.Tensor)
        if self.kind == "G":
            y = p.call("mpbp_apply_G", x, N, 4 * N)
            return y[lo:lo + 2 * N]
        xt, _ = p._to_dev(x, 2 * N)
        full = torch.zeros(4 * N, dtype=torch.float64, device=p.device)
        full[lo:lo + 2 * N] = xt
        if self.kind == "D":
            y = DivergenceOperator(p).matvec(full)
        else:
            y = p.call("mpbp_apply_F", full, 4 * N, 4 * N)[lo:lo + 2 * N]
        return y.cpu().numpy() if was_np else y

    __matmul__ = matvec
    dot = matvec


def _phase_blocks(bp: MultiphaseBlockPreconditioner, is_ths: bool):
    # L: F with c=0, xi=0, d_u=1, eta=1; XI: F with c=0, eta=0, d_u=-1 (F = XI_block + d_u*L, :331-337)
    pL = bp.plan(0.0, 1.0, 1.0, -1.0, xi=0.0, eta_n=1.0, eta_s=1.0, operators_only=True)
    pX = bp.plan(0.0, -1.0, 1.0, -1.0, eta_n=0.0, eta_s=0.0, operators_only=True)
    return _PhaseOp(pL, "L", is_ths), _PhaseOp(pL, "D", is_ths), _PhaseOp(pX, "XI", is_ths), _PhaseOp(pL, "G", is_ths)
