"""Mirror of the hot-path part of the reference's `solve.py`: the Krylov drivers the reference calls
(`pyamg.krylov.fgmres` at solve.py:207/:237/:285, `scipy.sparse.linalg.gmres` at :12/:221), `Jacobi`
(:149-159), `print_true_res_norm` (:161-170), `main` (:17-84), `solve_without_pc` (:202-208) and
`solve_with_approx_schur_pc` (:240-286).  All vector work runs in libmpbp.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._cabi import ITER_CB, SIDE_LEFT, SIDE_RIGHT, GmresOpts, check
from .preconditioner import (ApproxSchurOperator, GtGOperator, MultiphaseBlockPreconditioner, SubSolver,
                             SystemOperator, VelocityOperator)
from .utils import PI, fill_sol_and_RHS_vecs, manufactured_device, print_norms


def _krylov(A, b, M, x0, rtol, restart, maxiter, side, force_iters=0, host=False, callback=None):
    """One call of mpbp_gmres / mpbp_gmres_host. Returns (x, info, history ndarray).
    callback(x_k) (right side only) is called after every inner iteration with the current iterate, formed on the
    device as x0 + Z y_k (pyamg's callback semantics, solve.py:285); numpy b -> numpy x_k, torch b -> torch x_k."""
    if not isinstance(A, SystemOperator):
        raise TypeError("A must be the SystemOperator returned by get_big_A_matrix")
    if M is not None and not (isinstance(M, ApproxSchurOperator) and M.plan is A.plan):
        raise TypeError("M must be None or the approx-Schur operator of the same plan as A")
    p = A.plan
    lib = p.lib
    o = GmresOpts()
    check(lib.mpbp_gmres_opts_default(C.byref(o)))
    o.rtol, o.restart, o.maxiter, o.side = float(rtol), int(restart), int(maxiter), int(side)
    o.use_precond = int(M is not None)
    x0_nz = x0 is not None and bool((x0 != 0).any())  # a zero initial guess costs nothing extra (solve.py:205-207)
    o.x0_nonzero = int(x0_nz)
    o.force_iters = int(force_iters)
    keep = []
    if callback is not None:
        if side != SIDE_RIGHT or host:
            raise NotImplementedError("the per-iteration iterate callback exists for the right-preconditioned device path")
        xk_buf = torch.empty(5 * p.N, dtype=torch.float64, device=p.device)
        as_np = not isinstance(b, torch.Tensor)

        def _cb(_user, it, relres):
            callback(xk_buf.cpu().numpy() if as_np else xk_buf.clone())
            return 0
        cfn = ITER_CB(_cb)
        keep += [xk_buf, cfn]
        o.xk_buf = C.c_void_p(xk_buf.data_ptr())
        o.iter_cb = C.cast(cfn, C.c_void_p)
    need = C.c_size_t()
    check(lib.mpbp_gmres_workspace_bytes(p.h, C.byref(o), C.byref(need)))
    if p.kry_ws is None or p.kry_ws.numel() < need.value:
        p.kry_ws = None
        p.kry_ws = torch.empty(need.value, dtype=torch.uint8, device=p.device)
    o.workspace = C.c_void_p(p.kry_ws.data_ptr())
    o.workspace_bytes = p.kry_ws.numel()
    cap = max(1, force_iters, restart * maxiter if side == SIDE_LEFT else maxiter)
    cap = min(cap, 1 << 20)
    hist = (C.c_double * cap)()
    nit, info = C.c_int(0), C.c_int(0)
    length = 5 * p.N
    with torch.cuda.device(p.device):
        if host:
            bh = np.ascontiguousarray(b, dtype=np.float64)
            xh = np.zeros(length) if not x0_nz else np.array(x0, dtype=np.float64, copy=True)
            check(lib.mpbp_gmres_host(p.h, bh.ctypes.data_as(C.c_void_p), xh.ctypes.data_as(C.c_void_p), C.byref(o),
                                      hist, cap, C.byref(nit), C.byref(info), p.stream()))
            x = xh
        else:
            bt, was_np = p._to_dev(b, length)
            if not x0_nz:
                xt = torch.empty(length, dtype=torch.float64, device=p.device)
            else:
                xt, _ = p._to_dev(x0, length)
                xt = xt.clone()
            check(lib.mpbp_gmres(p.h, bt.data_ptr(), xt.data_ptr(), C.byref(o), hist, cap, C.byref(nit),
                                 C.byref(info), p.stream()))
            x = p._out(xt, was_np)
    return x, info.value, np.array(hist[: min(nit.value, cap)])


def fgmres(A, b, x0=None, tol=1e-5, restart=None, maxiter=None, M=None, callback=None, residuals=None, host=False):
    """Right-preconditioned flexible GMRES with the call shape of `pyamg.krylov.fgmres` as the
    reference uses it (solve.py:207, :237, :285).  Returns (x, info); info 0 = converged.

    restart=None: one cycle of at most `maxiter` inner iterations.  `residuals` (a list) receives the
    relative recurrence-residual history.  `callback(x_k)` is called once per inner iteration with the current
    iterate x_k = x0 + Z y_k (pyamg's semantics; one multi-axpy per iteration, no extra preconditioner apply).
    """
    if maxiter is None:
        maxiter = min(A.shape[0], 40)
    m = maxiter if restart is None else restart
    total = maxiter if restart is None else restart * maxiter
    x, info, hist = _krylov(A, b, M, x0, tol, m, total, SIDE_RIGHT, host=host, callback=callback)
    if residuals is not None:
        residuals[:] = list(hist)
    fgmres.last_history = hist
    return x, info


def last_hessenberg(A):
    """(k+1) x k upper Hessenberg matrix of the last Arnoldi cycle of the last fgmres / gmres call on A's plan."""
    p = A.plan
    k = C.c_int(0)
    check(p.lib.mpbp_gmres_last_hessenberg(p.h, None, 0, C.byref(k)))
    H = np.zeros((k.value + 1, max(k.value, 1)))
    if k.value:
        check(p.lib.mpbp_gmres_last_hessenberg(p.h, H.ctypes.data_as(C.POINTER(C.c_double)), H.shape[1], C.byref(k)))
    return H[:, :k.value]


def spectral_diagnostics(A, nev=10):
    """Replacement of the reference's dense spectrum analysis (compute_preconditioned_A + get_eigenvals / SLEPc,
    solve.py:103-200, :304-309) as a by-product of the solve: Ritz values of the preconditioned operator A M^-1
    (fgmres; M A for gmres) from the Hessenberg matrix of the last Arnoldi cycle.  Returns a dict with all Ritz
    values, the `nev` largest in magnitude (what SLEPc's default EPS returns, solve.py:121), the harmonic Ritz
    values (better for the eigenvalues closest to 0) and the spread max|theta| / min|theta|."""
    H = last_hessenberg(A)
    k = H.shape[1]
    if k == 0:
        return {"k": 0, "ritz": np.zeros(0), "largest": np.zeros(0), "harmonic": np.zeros(0), "spread": float("nan")}
    Hk = H[:k, :k]
    ritz = np.linalg.eigvals(Hk)
    ek = np.zeros(k)
    ek[-1] = 1.0
    try:  # harmonic Ritz values: eig(H_k + h_{k+1,k}^2 H_k^{-H} e_k e_k^T)
        f = np.linalg.solve(Hk.conj().T, ek)
        harm = np.linalg.eigvals(Hk + (H[k, k - 1] ** 2) * np.outer(f, ek))
    except np.linalg.LinAlgError:
        harm = ritz
    order = np.argsort(-np.abs(ritz))
    return {"k": k, "ritz": ritz[order], "largest": ritz[order][:nev], "harmonic": harm[np.argsort(-np.abs(harm))],
            "spread": float(np.abs(ritz).max() / max(np.abs(ritz).min(), 1e-300))}


def gmres(A, b, x0=None, *, rtol=1e-5, atol=0.0, restart=None, maxiter=None, M=None, callback=None,
          callback_type=None, host=False):
    """`scipy.sparse.linalg.gmres` semantics on the GPU (left preconditioning, MGS, Givens, restart,
    presid/ptol tolerance control; scipy/sparse/linalg/_isolve/iterative.py).  callback_type
    'pr_norm' (default here) receives presid/||b|| once per inner iteration.  Returns (x, info)."""
    if atol != 0.0:
        raise NotImplementedError("atol is not supported (the reference never passes it)")
    if restart is None:
        restart = 20
    restart = min(restart, A.shape[0])
    if maxiter is None:
        maxiter = 10 * A.shape[0]
    x, info, hist = _krylov(A, b, M, x0, rtol, restart, maxiter, SIDE_LEFT, host=host)
    gmres.last_history = hist
    if callback is not None:
        if callback_type not in (None, "pr_norm", "legacy"):
            raise NotImplementedError("only callback_type='pr_norm' is supported")
        for v in hist:
            callback(v)
    return x, info


def Jacobi(A, b, N, x, omega=1.0):
    """solve.Jacobi (solve.py:149-159): N sweeps x <- (b - R x)/diag(A), on F or Gt_G.
    omega < 1 gives the damped variant x <- x + omega (b - A x)/diag(A)."""
    p = A.plan
    if isinstance(A, VelocityOperator):
        fn, length = p.lib.mpbp_jacobi_F, 4 * p.N
    elif isinstance(A, GtGOperator):
        fn, length = p.lib.mpbp_jacobi_P, p.N
    else:
        raise TypeError("Jacobi is defined on the F and Gt_G operators")
    bt, was_np = p._to_dev(b, length)
    xt, _ = p._to_dev(x, length)
    xt = xt.clone()
    with torch.cuda.device(p.device):
        check(fn(p.h, bt.data_ptr(), xt.data_ptr(), int(N), float(omega), p.stream()))
    return p._out(xt, was_np)


def print_true_res_norm(A, b_vec, out=None, verbose=True):
    """solve.py:161-170: callback printing ||b - A x_k|| each iteration (values appended to `out`)."""
    iteration = 0
    bn = float(np.linalg.norm(b_vec)) if not isinstance(b_vec, torch.Tensor) else float(torch.linalg.norm(b_vec))

    def callback(xk):
        nonlocal iteration
        iteration += 1
        residual = b_vec - (A @ xk)
        rn = float(np.linalg.norm(residual)) if not isinstance(residual, torch.Tensor) else float(torch.linalg.norm(residual))
        if out is not None:
            out.append(rn / bn)
        if verbose:
            print(f"GMRES Iteration {iteration}: True residual norm = {rn}, Rel residual norm: {rn / bn}")

    return callback


def main(n: int = 4, c: int = 1, d: int = -1, xi: float = 1.0, eta_n: float = 1.0, eta_s: float = 1.0,
         sub_solver: SubSolver | None = None, device_vectors: bool = False):
    """solve.main (solve.py:17-84): builds A and the manufactured (b_vec, u_vec) (variable-theta branch).
    device_vectors=True assembles the two vectors on the GPU (torch tensors) instead of on the host."""
    block_prec = MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub_solver)
    A, S, F, D, G = block_prec.get_big_A_matrix(c=c, d_u=d)
    if device_vectors:
        u_vec, b_vec = manufactured_device(A.plan)
        return A, b_vec, u_vec
    nu, etan, etas = 1.0, eta_n, eta_s
    u_n_x_fcn = lambda y, x: np.sin(2*PI*x)*np.cos(2*PI*y)
    u_n_y_fcn = lambda y, x: np.cos(2*PI*x)*np.sin(2*PI*y)
    u_s_x_fcn = lambda y, x: -np.sin(2*PI*x)*np.cos(2*PI*y)
    u_s_y_fcn = lambda y, x: -np.cos(2*PI*x)*np.sin(2*PI*y)
    p_fcn = lambda y, x: 0.0
    sxy = lambda y, x: np.sin(2*PI*x)*np.sin(2*PI*y)
    brk_n = lambda y, x: 4*c*nu - 4*d*(8*etan*nu*PI*PI + xi) + 2*nu*(c - 16*d*etan*PI*PI)*sxy(y, x) + d*xi*sxy(y, x)**2
    brk_s = lambda y, x: -4*c*nu + 4*d*(8*etas*nu*PI*PI + xi) + 2*nu*(c - 16*d*etas*PI*PI)*sxy(y, x) - d*xi*sxy(y, x)**2
    b_n_x_fcn = lambda y, x: np.cos(2*PI*y)*np.sin(2*PI*x)*brk_n(y, x)/(8*nu)   # solve.py:72
    b_n_y_fcn = lambda y, x: np.cos(2*PI*x)*np.sin(2*PI*y)*brk_n(y, x)/(8*nu)   # solve.py:73
    b_s_x_fcn = lambda y, x: np.cos(2*PI*y)*np.sin(2*PI*x)*brk_s(y, x)/(8*nu)   # solve.py:75
    b_s_y_fcn = lambda y, x: np.cos(2*PI*x)*np.sin(2*PI*y)*brk_s(y, x)/(8*nu)   # solve.py:76
    b_p_fcn = lambda y, x: -PI*np.sin(4*PI*x)*np.sin(4*PI*y)                     # solve.py:78
    u_vec, b_vec = fill_sol_and_RHS_vecs(n, u_n_x_fcn, u_n_y_fcn, u_s_x_fcn, u_s_y_fcn, p_fcn,
                                         b_n_x_fcn, b_n_y_fcn, b_s_x_fcn, b_s_y_fcn, b_p_fcn)
    return A, b_vec, u_vec


def solve_without_pc(n, A, b_vec, u_vec, verbose=True):
    """solve.py:202-208."""
    x_initial = np.zeros(5 * n * n)
    cb = print_true_res_norm(A, b_vec, verbose=verbose) if verbose else None
    u_approx, info = fgmres(A, b_vec, M=None, x0=x_initial, tol=1e-8, maxiter=100, callback=cb)
    if verbose:
        print_norms(u_approx, u_vec, 1 / n, 1 / n, n)
    return u_approx, info


def _fgmres_host_M(A, b, M, tol, maxiter, callback=None):
    """Right-preconditioned FGMRES for an ARBITRARY host callable M (small verification problems): A.x runs on the GPU
    operator, the Arnoldi bookkeeping in numpy.  Same algorithm as mpbp_gmres RIGHT (MGS, Givens, one cycle)."""
    b = np.asarray(b, dtype=np.float64)
    m = maxiter
    x = np.zeros_like(b)
    bn = np.linalg.norm(b) or 1.0
    V = np.zeros((m + 1, len(b)))
    Z = np.zeros((m, len(b)))
    H = np.zeros((m + 1, m))
    cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
    g[0] = np.linalg.norm(b)
    V[0] = b / g[0]
    hist, info, jd = [], maxiter, 0
    for j in range(m):
        Z[j] = M(V[j])
        w = A @ Z[j]
        for i in range(j + 1):
            H[i, j] = np.dot(V[i], w)
            w = w - H[i, j] * V[i]
        H[j + 1, j] = np.linalg.norm(w)
        if H[j + 1, j] != 0:
            V[j + 1] = w / H[j + 1, j]
        for i in range(j):
            t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
            H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
            H[i, j] = t
        den = np.hypot(H[j, j], H[j + 1, j])
        cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
        H[j, j], H[j + 1, j] = den, 0.0
        g[j + 1] = -sn[j] * g[j]
        g[j] = cs[j] * g[j]
        jd = j + 1
        hist.append(abs(g[j + 1]) / bn)
        if callback is not None:
            callback(x + np.linalg.solve(np.triu(H[:jd, :jd]), g[:jd]) @ Z[:jd])
        if hist[-1] < tol:
            info = 0
            break
    y = np.linalg.solve(np.triu(H[:jd, :jd]), g[:jd])
    return x + y @ Z[:jd], info, np.array(hist)


def solve_with_exact_schur_pc(n, xi, etan, etas, c, d, b_vec, u_vec, verbose=True):
    """solve.py:210-238, the reference's small-n cross-check of the block preconditioner: the exact block-LU inverse
    with dense pinv(F) / lstsq and an inner scipy gmres on the dense exact Schur complement S = -D F^-1 G
    (preconditioner.py:343-346), first applied once to b (a direct solve), then used as M in fGMRES(maxiter=40).
    Dense O(n^6) host algebra on matrices pulled from the GPU operators: n <= ExactSchurOperator.MAX_N.
    Returns (u_direct, u_gmres, info, history)."""
    from scipy.linalg import lstsq, pinv
    from scipy.sparse.linalg import gmres as scipy_gmres
    dx = dy = 1 / n
    block_prec = MultiphaseBlockPreconditioner(n, xi, etan, etas)
    A, S, F, D, G = block_prec.get_big_A_matrix(c=c, d_u=d)
    Sd, Fd, Dd, Gd = S.toarray(), F.toarray(), D.toarray(), G.toarray()
    A_inv = pinv(Fd)                                             # :217
    nF = Fd.shape[1]

    def exact_schur_op(v):                                       # :216-227
        Ainv_v = lstsq(Fd, v[:nF])[0]
        rhs_interim = np.matmul(Dd, Ainv_v) + v[nF:]
        x_p = -1.0 * scipy_gmres(Sd, rhs_interim)[0]
        G_xp = np.matmul(Gd, x_p)
        Ainv_G_xp = np.matmul(A_inv, G_xp)
        return np.concatenate((Ainv_v - Ainv_G_xp, x_p), axis=0)
    b_vec = np.asarray(b_vec, dtype=np.float64)
    u_direct = exact_schur_op(b_vec)                             # :229
    if verbose:
        print("\nPrinting error norms for solving Ax=b using schur complement:")
        print_norms(u_direct, u_vec, dx, dy, n)
        print("\nPrinting error norms for solving Ax=b using GMRES with exact Schur complement as preconditioner:")
    cb = print_true_res_norm(A, b_vec, verbose=verbose) if verbose else None
    u_gmres, info, hist = _fgmres_host_M(A, b_vec, exact_schur_op, 1e-8, 40, callback=cb)   # :237
    if verbose:
        print_norms(u_gmres, u_vec, dx, dy, n)
    return u_direct, u_gmres, info, hist


def solve_with_approx_schur_pc(n, xi, etan, etas, c, d, b_vec, u_vec, sub_solver: SubSolver | None = None,
                               tol=1e-8, maxiter=150, restart=None, side="right", verbose=True):
    """solve.py:240-286: fGMRES on A with the approximate-Schur block preconditioner.
    Returns (u_approx, info, relative residual history)."""
    block_prec = MultiphaseBlockPreconditioner(n, xi, etan, etas, sub_solver=sub_solver)
    A, S, F, D, G = block_prec.get_big_A_matrix(c=c, d_u=d)          # :243-244
    approx_schur = block_prec.approx_schur_operator(c=c, d_u=d)       # :246-281
    if verbose:
        print("\nPrinting error norms for solving Ax=b using fGMRES with approx schur complement as preconditioner:")
    if side == "right":
        u_approx, info = fgmres(A, b_vec, M=approx_schur, tol=tol, maxiter=maxiter, restart=restart)   # :285
        hist = fgmres.last_history
    else:
        u_approx, info = gmres(A, b_vec, M=approx_schur, rtol=tol, restart=restart, maxiter=maxiter)
        hist = gmres.last_history
    if verbose:
        for k, r in enumerate(hist):
            print(f"GMRES Iteration {k + 1}: Rel residual norm: {r}")
        print_norms(u_approx, u_vec, 1 / n, 1 / n, n)                 # :286
    return u_approx, info, hist


if __name__ == "__main__":
    # defaults of solve.py:291-297
    n, c, d, xi, eta_n, eta_s = 16, 1, -1, 1.0, 100.0, 1.0
    A, b_vec, u_vec = main(n=n, c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
    solve_with_approx_schur_pc(n, xi, eta_n, eta_s, c, d, b_vec, u_vec)
