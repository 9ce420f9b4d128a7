"""CPU oracle for the block-preconditioned Krylov solve of the two-phase MAC Stokes system.

TEST INFRASTRUCTURE ONLY.  This module is a numpy/scipy restatement of the reference's
algorithm for the hot path (operator definition, block preconditioner, relaxation, Krylov).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it; the
product package never does.

Parity pinning: the operator / RHS / preconditioner-structure restatements are pinned against the
reference's own dense code (imported from /root/reference in the build container by
`tests/golden/make_golden.py`; the resulting fixtures are committed under `tests/golden/`).
The two third-party pieces the reference calls and that are absent everywhere (ilupp 1.0.2 ILUT,
pyamg fgmres) are *not* pinned by any reference test ("parity unpinned" for those): the sub-solves
are defined here (damped Jacobi as `solve.Jacobi`, and the multigrid the reference's comments name)
and the Krylov method is scipy's own `gmres` (the "scipy path" of BASELINE.json).

Reference citations are to /root/reference/<file>:<line>.

Index conventions (preconditioner.py:100-106, utils.py:178-208): fields are n x n, row-major
k = r*n + c, unknown ordering [u_n | v_n | u_s | v_s | p].  Row index r grows downward
(y = -(r+1/2)h at cell centres), u[r,c] sits on the left face of cell (r,c), v[r,c] on its top face.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

PI = np.pi


# ----------------------------------------------------------------------------------------------
# coefficient field (preconditioner.py:9-15)
# ----------------------------------------------------------------------------------------------
def thn(y, x):
    """Network volume fraction, preconditioner.py:9-11."""
    return 0.25 * np.sin(2 * PI * x) * np.sin(2 * PI * y) + 0.5


def ths(y, x):
    """Solvent volume fraction, preconditioner.py:13-15."""
    return 1.0 - thn(y, x)


def _rc(n):
    r = np.arange(n, dtype=np.float64)[:, None]
    c = np.arange(n, dtype=np.float64)[None, :]
    return r, c


def cell_theta(n):
    """theta_n at cell centres x=(c+1/2)h, y=-(r+1/2)h (preconditioner.py:26-72 sample points)."""
    h = 1 / n
    r, c = _rc(n)
    return thn(-(r + 0.5) * h, (c + 0.5) * h)


def face_theta(n):
    """theta_n sampled analytically AT the faces (mass term, preconditioner.py:325-326)."""
    h = 1 / n
    r, c = _rc(n)
    tu = thn(-(r + 0.5) * h, c * h + 0 * r)
    tv = thn(-r * h, (c + 0.5) * h)
    return tu, tv


# shifts: W(a)[r,c]=a[r,c-1], E -> c+1, N -> r-1, S -> r+1 (periodic)
def W(a):
    return np.roll(a, 1, axis=1)


def E(a):
    return np.roll(a, -1, axis=1)


def Nn(a):
    return np.roll(a, 1, axis=0)


def S(a):
    return np.roll(a, -1, axis=0)


def restrict_cell(a):
    """4-cell average of a cell-centred field (coarse cell (R,C) <- fine (2R..2R+1, 2C..2C+1))."""
    return 0.25 * (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2])


# ----------------------------------------------------------------------------------------------
# stencil term lists -> sparse matrices
# ----------------------------------------------------------------------------------------------
class _Terms:
    """Collects (out_field, in_field, dr, dc, coef[n,n]) terms and assembles a CSR matrix."""

    def __init__(self, n, n_out, n_in):
        self.n, self.n_out, self.n_in = n, n_out, n_in
        self.rows, self.cols, self.vals = [], [], []
        r, c = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        self._r, self._c = r, c

    def add(self, fo, fi, dr, dc, coef):
        n = self.n
        N = n * n
        rr = (self._r + dr) % n
        cc = (self._c + dc) % n
        self.rows.append((fo * N + self._r * n + self._c).ravel())
        self.cols.append((fi * N + rr * n + cc).ravel())
        self.vals.append(np.broadcast_to(coef, (n, n)).astype(np.float64).ravel())

    def csr(self):
        N = self.n * self.n
        m = sp.coo_matrix(
            (np.concatenate(self.vals), (np.concatenate(self.rows), np.concatenate(self.cols))),
            shape=(self.n_out * N, self.n_in * N),
        ).tocsr()
        m.sum_duplicates()
        return m


def phase_blocks(t, xi):
    """One phase's L (2N x 2N), D (N x 2N), XI (2N diag), G (2N x N) as sparse matrices.

    Restates preconditioner.py:86-297 (`get_block_matrices`) for the cell-centred fraction `t`
    (theta_n, or 1-theta_n when is_ths, :74-81).
    """
    n = t.shape[0]
    h = 1.0 / n
    ih2 = 1.0 / (h * h)
    ih = 1.0 / h
    tE, tW = t, W(t)
    node = 0.25 * (t + W(t) + Nn(t) + Nn(W(t)))  # top-left corner of cell (r,c); :112, :195
    nN, nS = node, S(node)
    nL, nR = node, E(node)
    tC, tN = t, Nn(t)
    fu = 0.5 * (t + W(t))  # :114 (thn_iph_j of the u-row), :200 (thn_imh_j of the v-row)
    fv = 0.5 * (t + Nn(t))  # :120, :198

    L = _Terms(n, 2, 2)
    # u rows, preconditioner.py:127-179
    L.add(0, 0, 0, 0, -(tE + tW) * ih2 - (nN + nS) * ih2)
    L.add(0, 0, 0, -1, tW * ih2)
    L.add(0, 0, 0, 1, tE * ih2)
    L.add(0, 0, -1, 0, nN * ih2)
    L.add(0, 0, 1, 0, nS * ih2)
    L.add(0, 1, 0, 0, (nN - tE) * ih2)
    L.add(0, 1, 0, -1, (tW - nN) * ih2)
    L.add(0, 1, 1, -1, (nS - tW) * ih2)
    L.add(0, 1, 1, 0, (tE - nS) * ih2)
    # v rows, preconditioner.py:242-295
    L.add(1, 1, 0, 0, -(tN + tC) * ih2 - (nL + nR) * ih2)
    L.add(1, 1, 0, -1, nL * ih2)
    L.add(1, 1, 0, 1, nR * ih2)
    L.add(1, 1, -1, 0, tN * ih2)
    L.add(1, 1, 1, 0, tC * ih2)
    L.add(1, 0, 0, 0, (nL - tC) * ih2)
    L.add(1, 0, 0, 1, (tC - nR) * ih2)
    L.add(1, 0, -1, 0, (tN - nL) * ih2)
    L.add(1, 0, -1, 1, (nR - tN) * ih2)

    G = _Terms(n, 2, 1)  # :203-219
    G.add(0, 0, 0, 0, fu * ih)
    G.add(0, 0, 0, -1, -fu * ih)
    G.add(1, 0, 0, 0, -fv * ih)
    G.add(1, 0, -1, 0, fv * ih)

    D = _Terms(n, 1, 2)  # :221-238
    D.add(0, 0, 0, 1, E(fu) * ih)
    D.add(0, 0, 0, 0, -fu * ih)
    D.add(0, 1, 0, 0, fv * ih)
    D.add(0, 1, 1, 0, -S(fv) * ih)

    XI = np.concatenate([(xi * fu * (1.0 - fu)).ravel(), (xi * fv * (1.0 - fv)).ravel()])  # :124-125
    return L.csr(), D.csr(), XI, G.csr()


class Operators:
    """Sparse restatement of `get_big_A_matrix` (preconditioner.py:299-349) and of the derived
    matrices of `solve_with_approx_schur_pc` (solve.py:246-249).

    theta: optional cell-centred theta_n array (default: the analytic field).
    mass: 'analytic' -> c*theta at the face evaluated analytically (the reference, :325-329);
          'average'  -> c * two-cell face average (used on rediscretised coarse levels).
    """

    def __init__(self, n, xi, eta_n, eta_s, c, d_u, d_p=1.0, d_div=-1.0, theta=None, mass="analytic"):
        self.n, self.xi, self.eta_n, self.eta_s = n, xi, eta_n, eta_s
        self.c, self.d_u, self.d_p, self.d_div = c, d_u, d_p, d_div
        N = n * n
        self.N = N
        tn = cell_theta(n) if theta is None else np.asarray(theta, dtype=np.float64)
        ts = 1.0 - tn
        self.theta = tn
        L_n, D_n, XI_n, G_n = phase_blocks(tn, xi)
        L_s, D_s, XI_s, G_s = phase_blocks(ts, xi)
        if mass == "analytic":
            mu, mv = face_theta(n)
        else:
            mu, mv = 0.5 * (tn + W(tn)), 0.5 * (tn + Nn(tn))
        w_n = c * np.concatenate([mu.ravel(), mv.ravel()])  # :325-326
        w_s = c * np.concatenate([(1.0 - mu).ravel(), (1.0 - mv).ravel()])  # :328-329
        L = sp.block_diag([eta_n * L_n, eta_s * L_s], format="csr")  # :310
        XIb = sp.bmat(
            [[sp.diags(w_n - d_u * XI_n), sp.diags(d_u * XI_n)], [sp.diags(d_u * XI_s), sp.diags(w_s - d_u * XI_s)]],
            format="csr",
        )  # :331-336
        self.F = (XIb + d_u * L).tocsr()  # :337
        self.D = sp.hstack([D_n, D_s], format="csr")  # :311 (un-negated, as returned :349)
        self.G = (d_p * sp.vstack([G_n, G_s], format="csr")).tocsr()  # :313
        self.A = sp.bmat([[self.F, self.G], [d_div * self.D, None]], format="csr")  # :339-341
        mD = -1.0 * self.D  # solve.py:246
        self.GtG = (mD @ self.G).tocsr()  # solve.py:247
        self.GtFG = (mD @ self.F @ self.G).tocsr()  # solve.py:248-249
        self.GtG.sum_duplicates()
        self.GtFG.sum_duplicates()


# ----------------------------------------------------------------------------------------------
# manufactured solution / RHS (solve.py:52-78 through utils.py:159-210), vectorised
# ----------------------------------------------------------------------------------------------
def manufactured(n, c, d, xi, eta_n, eta_s, nu=1.0, b_p_sign=-1.0):
    """(u_vec, b_vec) of `solve.main` (variable-theta branch, solve.py:70-81).

    b_p_sign=-1 is solve.py:78; +1 reproduces apply.py:66.
    """
    h = 1 / n
    r, cc = _rc(n)
    etan, etas = eta_n, eta_s
    yu, xu = -(r + 0.5) * h + 0 * cc, cc * h + 0 * r  # utils.py:187
    yv, xv = -r * h + 0 * cc, (cc + 0.5) * h + 0 * r  # utils.py:188
    yp, xp = -(r + 0.5) * h + 0 * cc, (cc + 0.5) * h + 0 * r  # utils.py:193

    u_n_x = lambda y, x: np.sin(2 * PI * x) * np.cos(2 * PI * y)
    u_n_y = lambda y, x: np.cos(2 * PI * x) * np.sin(2 * PI * y)
    u_s_x = lambda y, x: -np.sin(2 * PI * x) * np.cos(2 * PI * y)
    u_s_y = lambda y, x: -np.cos(2 * PI * x) * np.sin(2 * PI * y)
    b_n_x = lambda y, x: (np.cos(2*PI*y)*np.sin(2*PI*x)*(4*c*nu-4*d*(8*etan*nu*PI*PI+xi)+2*nu*(c-16*d*etan*PI*PI)*np.sin(2*PI*x)*np.sin(2*PI*y)+d*xi*np.sin(2*PI*x)*np.sin(2*PI*x)*np.sin(2*PI*y)*np.sin(2*PI*y)))/(8*nu)
    b_n_y = lambda y, x: (np.cos(2*PI*x)*np.sin(2*PI*y)*(4*c*nu-4*d*(8*etan*nu*PI*PI+xi)+2*nu*(c-16*d*etan*PI*PI)*np.sin(2*PI*x)*np.sin(2*PI*y)+d*xi*np.sin(2*PI*x)*np.sin(2*PI*x)*np.sin(2*PI*y)*np.sin(2*PI*y)))/(8*nu)
    b_s_x = lambda y, x: (np.cos(2*PI*y)*np.sin(2*PI*x)*(-4*c*nu+4*d*(8*etas*nu*PI*PI+xi)+2*nu*(c-16*d*etas*PI*PI)*np.sin(2*PI*x)*np.sin(2*PI*y)-d*xi*np.sin(2*PI*x)*np.sin(2*PI*x)*np.sin(2*PI*y)*np.sin(2*PI*y)))/(8*nu)
    b_s_y = lambda y, x: (np.cos(2*PI*x)*np.sin(2*PI*y)*(-4*c*nu+4*d*(8*etas*nu*PI*PI+xi)+2*nu*(c-16*d*etas*PI*PI)*np.sin(2*PI*x)*np.sin(2*PI*y)-d*xi*np.sin(2*PI*x)*np.sin(2*PI*x)*np.sin(2*PI*y)*np.sin(2*PI*y)))/(8*nu)
    b_p = lambda y, x: b_p_sign * PI * np.sin(4 * PI * x) * np.sin(4 * PI * y)

    u_vec = np.concatenate([u_n_x(yu, xu).ravel(), u_n_y(yv, xv).ravel(), u_s_x(yu, xu).ravel(),
                            u_s_y(yv, xv).ravel(), np.zeros(n * n)])
    b_vec = np.concatenate([b_n_x(yu, xu).ravel(), b_n_y(yv, xv).ravel(), b_s_x(yu, xu).ravel(),
                            b_s_y(yv, xv).ravel(), b_p(yp, xp).ravel()])
    return u_vec, b_vec


# error norms, utils.py:7-17
def weighted_L2(a, b, w):
    q = a - b
    return np.sqrt((w * q * q).sum())


def weighted_L1(a, b, w):
    return (w * np.abs(a - b)).sum()


def max_norm(a, b):
    return np.abs(a - b).max()


# ----------------------------------------------------------------------------------------------
# relaxation (solve.py:149-159) and the multigrid sub-solvers
# ----------------------------------------------------------------------------------------------
def jacobi(A, b, N, x, omega=1.0):
    """`solve.Jacobi` (solve.py:149-159) with optional damping: x <- x + omega*(b - A x)/diag(A).

    omega=1 is algebraically the reference's x <- (b - R x)/D.
    """
    if not sp.issparse(A):
        A = sp.csr_matrix(A)
    dg = A.diagonal()
    x = np.array(x, dtype=np.float64, copy=True)
    for _ in range(N):
        x = x + omega * (b - mv(A, x)) / dg
    return x


def restrict_u(f):
    """Full weighting of a left-face field: (1/4,1/2,1/4) across columns 2C-1,2C,2C+1; (1/2,1/2) over rows."""
    a = 0.5 * (f[0::2, :] + f[1::2, :])
    return 0.25 * W(a)[:, 0::2] + 0.5 * a[:, 0::2] + 0.25 * E(a)[:, 0::2]


def restrict_v(f):
    """Full weighting of a top-face field: (1/4,1/2,1/4) across rows 2R-1,2R,2R+1; (1/2,1/2) over columns."""
    a = 0.5 * (f[:, 0::2] + f[:, 1::2])
    return 0.25 * Nn(a)[0::2, :] + 0.5 * a[0::2, :] + 0.25 * S(a)[0::2, :]


def prolong_u(fc):
    """4 x transpose of restrict_u: linear in x, piecewise constant in y."""
    nc = fc.shape[0]
    a = np.empty((nc, 2 * nc))
    a[:, 0::2] = fc
    a[:, 1::2] = 0.5 * (fc + E(fc))
    out = np.empty((2 * nc, 2 * nc))
    out[0::2, :] = a
    out[1::2, :] = a
    return out


def prolong_v(fc):
    nc = fc.shape[0]
    a = np.empty((2 * nc, nc))
    a[0::2, :] = fc
    a[1::2, :] = 0.5 * (fc + S(fc))
    out = np.empty((2 * nc, 2 * nc))
    out[:, 0::2] = a
    out[:, 1::2] = a
    return out


def prolong_cell(fc):
    return np.kron(fc, np.ones((2, 2)))


class SubSolverConfig:
    """Definition of the two approximate solves (the slots ilupp.ILUT fills in solve.py:251/254).

    kind: 'jacobi'  -> `sweeps` damped-Jacobi sweeps from a zero initial guess
          'mg'      -> `cycles` V(nu1,nu2) cycles (rediscretised hierarchy down to n_coarse, dense
                       (pseudo-)inverse there), optionally Chebyshev-accelerated (`cheb=True`) over the
                       interval [lmin, lmax] containing the spectrum of the V-cycle-preconditioned
                       operator (measured: [0.785, 1.16] for F, [0.866, 1.16] for GtG).
    """

    def __init__(self, kind="mg", sweeps=20, omega=0.8, cycles=2, nu1=2, nu2=2, n_coarse=4,
                 cheb=False, lmin=0.75, lmax=1.2, project=True):
        self.kind, self.sweeps, self.omega = kind, sweeps, omega
        self.cycles, self.nu1, self.nu2, self.n_coarse = cycles, nu1, nu2, n_coarse
        self.cheb, self.lmin, self.lmax, self.project = cheb, lmin, lmax, project


class _Level:
    pass


# ----------------------------------------------------------------------------------------------
# rounding model for the conditioning envelopes of the residual-history tests (tests/golden/make_golden.py)
# ----------------------------------------------------------------------------------------------
# Any correct fp64 evaluation of a stencil row y_i = sum_j a_ij x_j satisfies |fl(y_i) - y_i| <= k u sum_j |a_ij||x_j|
# (the standard backward-error bound); HOW the error is distributed depends on the evaluation order -- this oracle
# sums coefficient x value terms (large cancellation on smooth fields), the CUDA kernels difference first (flux
# form).  With the model switched on, every sparse mat-vec of the sub-solvers returns
# y_i + amp * sqrt(k_i) * (|A| |x|)_i * N(0,1), k_i = number of terms of row i (the root-sum-square of k_i independent
# unit-roundoff errors, the usual probabilistic refinement of the k u bound): the set of results a correct
# implementation with another evaluation order may return.
_ROUND = {"amp": 0.0, "rng": None, "abs": {}}


def set_rounding_model(amp=0.0, seed=0):
    _ROUND["amp"] = float(amp)
    _ROUND["rng"] = np.random.default_rng(seed) if amp > 0 else None


def mv(A, x):
    """A @ x, plus the rounding model's perturbation when it is switched on (sparse A only)."""
    y = A @ x
    if _ROUND["amp"] > 0.0 and sp.issparse(A):
        key = id(A)
        ent = _ROUND["abs"].get(key)
        if ent is None or ent[0] is not A:
            Ac = A.tocsr()
            ent = (A, abs(Ac), np.sqrt(np.maximum(np.diff(Ac.indptr), 1).astype(np.float64)))
            _ROUND["abs"][key] = ent
        y = y + _ROUND["amp"] * ent[2] * (ent[1] @ np.abs(x)) * _ROUND["rng"].standard_normal(len(y))
    return y


class Multigrid:
    """Rediscretised geometric multigrid for F (4N unknowns) and GtG (N unknowns)."""

    def __init__(self, ops: Operators, cfg: SubSolverConfig):
        self.cfg = cfg
        self.levels = []
        n, theta = ops.n, ops.theta
        cur = ops
        while True:
            lv = _Level()
            lv.n, lv.F, lv.P = cur.n, cur.F, cur.GtG
            lv.dF, lv.dP = cur.F.diagonal(), cur.GtG.diagonal()
            self.levels.append(lv)
            if n <= cfg.n_coarse or n % 2:
                break
            n //= 2
            theta = restrict_cell(theta)
            cur = Operators(n, ops.xi, ops.eta_n, ops.eta_s, ops.c, ops.d_u, ops.d_p, ops.d_div,
                            theta=theta, mass="average")
        last = self.levels[-1]
        last.Finv = np.linalg.inv(last.F.toarray())
        Nc = last.n * last.n
        e = np.full((Nc, 1), 1.0 / np.sqrt(Nc))
        # pseudo-inverse of the singular SPSD 5-point operator: (P + e e^T)^-1 - e e^T
        last.Pinv = np.linalg.inv(last.P.toarray() + e @ e.T) - e @ e.T

    # transfers on stacked vectors
    @staticmethod
    def _rF(r, n):
        f = r.reshape(4, n, n)
        return np.concatenate([restrict_u(f[0]).ravel(), restrict_v(f[1]).ravel(),
                               restrict_u(f[2]).ravel(), restrict_v(f[3]).ravel()])

    @staticmethod
    def _pF(e, nc):
        f = e.reshape(4, nc, nc)
        return np.concatenate([prolong_u(f[0]).ravel(), prolong_v(f[1]).ravel(),
                               prolong_u(f[2]).ravel(), prolong_v(f[3]).ravel()])

    def _vcycle(self, which, l, b):
        lv, cfg = self.levels[l], self.cfg
        A, dg = (lv.F, lv.dF) if which == "F" else (lv.P, lv.dP)
        if l == len(self.levels) - 1:
            return (lv.Finv if which == "F" else lv.Pinv) @ b
        x = cfg.omega * b / dg
        for _ in range(cfg.nu1 - 1):
            x = x + cfg.omega * (b - mv(A, x)) / dg
        r = b - mv(A, x)
        n = lv.n
        if which == "F":
            rc = self._rF(r, n)
        else:
            rc = restrict_cell(r.reshape(n, n)).ravel()
        ec = self._vcycle(which, l + 1, rc)
        if which == "F":
            x = x + self._pF(ec, n // 2)
        else:
            x = x + prolong_cell(ec.reshape(n // 2, n // 2)).ravel()
        for _ in range(cfg.nu2):
            x = x + cfg.omega * (b - mv(A, x)) / dg
        return x

    def solve(self, which, b):
        """Fixed linear operator b -> x~ = B b (zero initial guess)."""
        cfg = self.cfg
        A = self.levels[0].F if which == "F" else self.levels[0].P
        if not cfg.cheb:
            x = self._vcycle(which, 0, b)
            for _ in range(cfg.cycles - 1):
                x = x + self._vcycle(which, 0, b - mv(A, x))
        else:
            # Chebyshev iteration on B A with spectrum in [lmin, lmax] (B = one V-cycle), k = cycles steps
            lmin, lmax = cfg.lmin, cfg.lmax
            th, de = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
            sig = th / de
            rho_k = 1.0 / sig
            z = self._vcycle(which, 0, b)
            dvec = z / th
            x = dvec.copy()
            for _ in range(cfg.cycles - 1):
                z = self._vcycle(which, 0, b - mv(A, x))
                rho_n = 1.0 / (2.0 * sig - rho_k)
                dvec = rho_n * rho_k * dvec + (2.0 * rho_n / de) * z
                x = x + dvec
                rho_k = rho_n
        if which == "P" and cfg.project:
            x = x - x.mean()
        return x


class SubSolver:
    """Object with `@` filling the `ilupp.ILUTPreconditioner(...)` slot (solve.py:251, :254)."""

    def __init__(self, ops: Operators, which: str, cfg: SubSolverConfig, mg: Multigrid | None = None):
        self.ops, self.which, self.cfg, self.mg = ops, which, cfg, mg
        self.A = ops.F if which == "F" else ops.GtG

    def __matmul__(self, b):
        cfg = self.cfg
        if cfg.kind == "jacobi":
            x = jacobi(self.A, b, cfg.sweeps, np.zeros_like(b), cfg.omega)
            if self.which == "P" and cfg.project:
                x = x - x.mean()
            return x
        if cfg.kind == "exact":
            return self._exact(b)
        return self.mg.solve(self.which, b)

    def _exact(self, b):
        if not hasattr(self, "_lu"):
            if self.which == "F":
                self._lu = spla.splu(sp.csc_matrix(self.A))
            else:
                Nn_ = self.A.shape[0]
                e = np.full((Nn_, 1), 1.0 / np.sqrt(Nn_))
                self._pinv = np.linalg.inv(self.A.toarray() + e @ e.T) - e @ e.T
        return self._lu.solve(b) if self.which == "F" else self._pinv @ b


class ApproxSchur:
    """`approx_schur_op` (solve.py:257-277), steps 1-8 of SURVEY 3.2, sign quirks preserved."""

    def __init__(self, ops: Operators, cfg: SubSolverConfig):
        self.ops, self.cfg = ops, cfg
        mg = Multigrid(ops, cfg) if cfg.kind == "mg" else None
        self.F_inv = SubSolver(ops, "F", cfg, mg)
        self.P_inv = SubSolver(ops, "P", cfg, mg)
        self.shape = ops.A.shape
        self.dtype = np.float64

    def matvec(self, v):
        o = self.ops
        nF = o.F.shape[1]
        Finv_v = self.F_inv @ v[:nF]  # :258
        rhs = mv(o.D, Finv_v) + v[nF:]  # :259
        x_a = self.P_inv @ rhs  # :265
        x_b = mv(o.GtFG, x_a)  # :267
        x_p = self.P_inv @ x_b  # :271
        G_xp = mv(o.G, x_p)  # :273
        Finv_G = self.F_inv @ G_xp  # :274
        return np.concatenate([Finv_v - Finv_G, x_p])  # :275-276

    def linear_operator(self):
        return spla.LinearOperator(shape=self.shape, matvec=self.matvec, dtype=np.float64)  # :280-281


# ----------------------------------------------------------------------------------------------
# Krylov
# ----------------------------------------------------------------------------------------------
def gmres_scipy(A, b, M=None, rtol=1e-8, restart=20, maxiter=None, x0=None):
    """The reference's "scipy path": scipy.sparse.linalg.gmres (left preconditioning, MGS, Givens).

    Returns (x, info, history) with history = callback 'pr_norm' values (presid/||b||), one per
    inner iteration.
    """
    hist = []
    x, info = spla.gmres(A, b, x0=x0, M=M, rtol=rtol, atol=0.0, restart=restart, maxiter=maxiter,
                         callback=hist.append, callback_type="pr_norm")
    return x, info, np.array(hist)


def fgmres(A, b, M=None, x0=None, tol=1e-8, maxiter=150, callback=None, restart=None):
    """Textbook right-preconditioned flexible GMRES with the call signature solve.py:285 uses
    (`pyamg.krylov.fgmres`; MGS instead of pyamg's Householder -- parity unpinned, pyamg is absent).

    One un-restarted cycle of at most `maxiter` inner iterations (restart=None), stop on
    ||r|| < tol*||b|| (recurrence residual); callback(x_k) every inner iteration.
    Returns (x, info); the residual history is left in fgmres.last_history.
    """
    matvec = (lambda z: A @ z)
    psolve = (lambda z: z) if M is None else (lambda z: M @ z if not hasattr(M, "matvec") else M.matvec(z))
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    bn = np.linalg.norm(b)
    if bn == 0:
        bn = 1.0
    m = maxiter if restart is None else restart
    hist = []
    it = 0
    info = maxiter
    while it < maxiter:
        r = b - matvec(x)
        beta = np.linalg.norm(r)
        if beta < tol * bn:
            info = 0
            break
        V = np.zeros((m + 1, n))
        Z = np.zeros((m, n))
        Hm = np.zeros((m + 1, m))
        Hraw = np.zeros((m + 1, m))  # the Hessenberg matrix before the Givens rotations (Ritz values)
        cs, sn = np.zeros(m), np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        V[0] = r / beta
        j_done = 0
        for j in range(m):
            Z[j] = psolve(V[j])
            w = matvec(Z[j])
            for i in range(j + 1):
                Hm[i, j] = np.dot(V[i], w)
                w = w - Hm[i, j] * V[i]
            Hm[j + 1, j] = np.linalg.norm(w)
            if Hm[j + 1, j] != 0:
                V[j + 1] = w / Hm[j + 1, j]
            Hraw[:, j] = Hm[:, j]
            fgmres.last_hessenberg = Hraw[: j + 2, : j + 1]
            for i in range(j):
                t = cs[i] * Hm[i, j] + sn[i] * Hm[i + 1, j]
                Hm[i + 1, j] = -sn[i] * Hm[i, j] + cs[i] * Hm[i + 1, j]
                Hm[i, j] = t
            den = np.hypot(Hm[j, j], Hm[j + 1, j])
            cs[j], sn[j] = Hm[j, j] / den, Hm[j + 1, j] / den
            Hm[j, j], Hm[j + 1, j] = den, 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            it += 1
            j_done = j + 1
            res = abs(g[j + 1])
            hist.append(res / bn)
            if callback is not None:
                y = np.linalg.solve(np.triu(Hm[:j_done, :j_done]), g[:j_done])
                callback(x + y @ Z[:j_done])
            if res < tol * bn or it >= maxiter:
                break
        y = np.linalg.solve(np.triu(Hm[:j_done, :j_done]), g[:j_done])
        x = x + y @ Z[:j_done]
        if res < tol * bn:
            info = 0
            break
    fgmres.last_history = np.array(hist)
    return x, info
