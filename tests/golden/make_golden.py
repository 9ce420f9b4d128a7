"""Generates tests/golden/*.npz by running the REFERENCE's own code (imported from /root/reference).

Run once in the build container (the reference checkout does not exist on the GPU box):
    python tests/golden/make_golden.py

The reference's `solve.py` imports matplotlib / petsc4py / slepc4py / ilupp / pyamg, none of which is
installed; stub modules are registered first.  The two stubs that carry
arithmetic are
  * ilupp.ILUTPreconditioner(A, fill_in, threshold)  -> the oracle's sub-solver object (`@`),
  * pyamg.krylov.fgmres(A, b, M=..., tol=, maxiter=, callback=) -> the oracle's fgmres restatement,
so `solve.solve_with_approx_schur_pc` (hence the verbatim `approx_schur_op` closure, solve.py:257-277),
`solve.main`, `solve.Jacobi` and `preconditioner.get_big_A_matrix` run unmodified.
"""
import contextlib
import io
import os
import re
import sys
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402

import mpbp_oracle as O  # noqa: E402

# ---- stubs ------------------------------------------------------------------------------------
_state = {}


class _ILUTStub:
    def __init__(self, A_csc, fill_in=None, threshold=None):
        assert fill_in == 100 and threshold == 0.001  # solve.py:251, :254
        ops, cfg = _state["ops"], _state["cfg"]
        which = "F" if A_csc.shape[0] == 4 * ops.N else "P"
        if "mg" not in _state and cfg.kind == "mg":
            _state["mg"] = O.Multigrid(ops, cfg)
        self.sub = O.SubSolver(ops, which, cfg, _state.get("mg"))

    def __matmul__(self, v):
        return self.sub @ v


def _fgmres_stub(A, b, M=None, x0=None, tol=1e-5, maxiter=None, callback=None, **kw):
    _state["M"] = M
    _state["A"] = A
    x, info = O.fgmres(A, b, M=M, x0=x0, tol=tol, maxiter=maxiter, callback=callback)
    _state["hist"] = O.fgmres.last_history.copy()
    _state["x"] = x.copy()
    return x, info


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "petsc4py", "slepc4py", "ilupp", "pyamg", "pyamg.krylov"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["petsc4py"].PETSc = object()
    sys.modules["slepc4py"].SLEPc = object()
    sys.modules["ilupp"].ILUTPreconditioner = _ILUTStub
    sys.modules["pyamg"].krylov = sys.modules["pyamg.krylov"]
    sys.modules["pyamg.krylov"].fgmres = _fgmres_stub


def _csr_parts(M, prefix):
    m = sp.csr_matrix(M)
    return {f"{prefix}_data": m.data, f"{prefix}_indices": m.indices, f"{prefix}_indptr": m.indptr,
            f"{prefix}_shape": np.array(m.shape)}


def main():
    _install_stubs()
    import preconditioner as refpc
    import solve as refsolve
    import utils as refutils

    rng = np.random.default_rng(20250101)
    # ---- operators, vectors, applies: (n, xi, eta_n, eta_s, c, d) ---------------------------------
    cases = [(4, 0.7, 1.0, 2.0, 1.3, -1.0), (8, 1.0, 100.0, 1.0, 1.0, -1.0), (16, 1.0, 100.0, 1.0, 1.0, -1.0),
             (12, 0.5, 3.0, 2.0, 0.7, -0.9), (16, 1.0, 1.0e4, 1.0, 1.0, -1.0)]
    for (n, xi, eta_n, eta_s, c, d) in cases:
        N = n * n
        bp = refpc.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s)
        A, S, F, D, G = bp.get_big_A_matrix(c=c, d_u=d)
        mD = -1.0 * D
        GtG = np.matmul(mD, G)
        GtFG = np.matmul(np.matmul(mD, F), G)
        x = rng.standard_normal(5 * N)
        out = {}
        out.update(_csr_parts(A, "A"))
        out.update(x=x, Ax=A @ x, Fx=F @ x[:4 * N], Dx=D @ x[:4 * N], Gp=G @ x[4 * N:], GtGp=GtG @ x[4 * N:],
                   GtFGp=GtFG @ x[4 * N:], params=np.array([n, xi, eta_n, eta_s, c, d]),
                   normA=np.linalg.norm(A), sumabsA=np.abs(A).sum(), traceF=np.trace(F), normF=np.linalg.norm(F),
                   normG=np.linalg.norm(G), normD=np.linalg.norm(D), normS=np.linalg.norm(S))
        # solve.main: manufactured vectors (through the reference's own lambdas and fill loop)
        A2, b_vec, u_vec = refsolve.main(n=n, c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
        assert np.array_equal(A, A2)
        out.update(b_vec=b_vec, u_vec=u_vec, Au=A @ u_vec, Ab=A @ b_vec)
        # solve.Jacobi verbatim (undamped, as written): 3 sweeps on F and on Gt_G from x=0
        bF, bP = x[:4 * N], x[4 * N:] - x[4 * N:].mean()
        out.update(jacF3=refsolve.Jacobi(F, bF, 3, 0 * bF), jacP3=refsolve.Jacobi(GtG, bP, 3, 0 * bP), bP=bP)
        np.savez_compressed(os.path.join(HERE, f"ops_n{n}_eta{int(eta_n)}.npz"), **out)
        print("ops", n, eta_n, "nnz", (A != 0).sum())

    # ---- preconditioner apply + Krylov through the reference's solve_with_approx_schur_pc ----------
    solves = [
        ("mgcheb", 16, 1.0, 100.0, 1.0, 1, -1, dict(kind="mg", cycles=4, cheb=True)),
        ("mgplain", 16, 1.0, 100.0, 1.0, 1, -1, dict(kind="mg", cycles=2, cheb=False)),
        ("jacobi", 16, 1.0, 100.0, 1.0, 1, -1, dict(kind="jacobi", sweeps=20, omega=0.8)),
        ("mgcheb", 32, 1.0, 1.0, 1.0, 1, -1, dict(kind="mg", cycles=4, cheb=True)),
    ]
    for (tag, n, xi, eta_n, eta_s, c, d, kw) in solves:
        N = n * n
        cfgF = O.SubSolverConfig(**kw)
        ops = O.Operators(n, xi, eta_n, eta_s, c, d)
        _state.clear()
        _state.update(ops=ops, cfg=cfgF)
        A, b_vec, u_vec = refsolve.main(n=n, c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
        true_res = []

        orig = refsolve.print_true_res_norm

        def cb_factory(Amat, bvec):
            inner = orig(Amat, bvec)

            def cb(xk):
                true_res.append(np.linalg.norm(bvec - Amat @ xk) / np.linalg.norm(bvec))
                inner(xk)
            return cb

        refsolve.print_true_res_norm = cb_factory
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            refsolve.solve_with_approx_schur_pc(n, xi, eta_n, eta_s, c, d, b_vec, u_vec)
        refsolve.print_true_res_norm = orig
        text = buf.getvalue()
        norms = [float(v) for v in re.findall(r"_norm for n = \d+ is ([0-9.eE+-]+)", text)]
        M = _state["M"]  # the reference's LinearOperator around its verbatim approx_schur_op closure
        v = rng.standard_normal(5 * N)
        v[4 * N:] -= v[4 * N:].mean()
        out = dict(params=np.array([n, xi, eta_n, eta_s, c, d]), b_vec=b_vec, u_vec=u_vec, v=v, Mv=M.matvec(v),
                   Mb=M.matvec(b_vec), hist=_state["hist"], true_res=np.array(true_res), x=_state["x"],
                   err_norms=np.array(norms))
        # conditioning of the history itself: what another correct fp64 implementation of the same algorithm may
        # legitimately return.  The same Krylov driver is re-run 17 times:
        #   * once with the operators of the independent C restatement of the oracle (oracle/mpbp_oracle_c.c),
        #   * 16 times with the oracle's operators under the standard rounding model (oracle/mpbp_oracle.py:mv):
        #     every sparse mat-vec y = B x of the system operator, of the sub-solvers and of the preconditioner's
        #     D / Gt_F_G / G stages returns y_i + u sqrt(k_i) (|B||x|)_i N(0,1) (u = 1.1e-16, k_i terms in row i), b a 1-ulp perturbation.
        #     That is the backward-error bound every correct evaluation satisfies; which part of it an implementation
        #     realises depends on its evaluation order (this oracle and the reference sum coefficient x value terms,
        #     the CUDA kernels difference first and apply Gt_F_G as the chain -D(F(G x))).
        # For FGMRES the C oracle's own Krylov driver adds three more runs.  The envelope is the largest relative
        # change of the history per iteration over all of them.
        import scipy.sparse.linalg as spla
        import c_oracle
        ckw = dict(kind=cfgF.kind, F_cycles=cfgF.cycles, P_cycles=cfgF.cycles, F_sweeps=cfgF.sweeps, P_sweeps=cfgF.sweeps,
                   omega=cfgF.omega, cheb=cfgF.cheb)
        co = c_oracle.COracle(n, xi, eta_n, eta_s, c, d, **ckw)
        m5 = len(b_vec)
        M_model = O.ApproxSchur(ops, cfgF)  # == the reference's closure up to rounding (tests/test_oracle_golden.py)
        assert np.abs(M_model.matvec(v) - M.matvec(v)).max() <= 1e-9 * np.abs(M.matvec(v)).max()

        def lin(f):
            return spla.LinearOperator((m5, m5), dtype=np.float64, matvec=f)

        def envelope(run, h0, extra=()):
            env = np.zeros(len(h0))
            prng = np.random.default_rng(99)

            def fold(h):
                k = min(len(h), len(h0))
                env[:k] = np.maximum(env[:k], np.abs(h[:k] - h0[:k]) / h0[:k])
                if len(h) != len(h0):
                    env[k:] = np.inf
            fold(run(b_vec, lin(lambda z: co.apply_A(np.ascontiguousarray(z))), lin(lambda z: co.precond(np.ascontiguousarray(z)))))
            for i in range(16):  # 16 runs under the rounding model
                O.set_rounding_model(1.1e-16, seed=1000 + i)
                bp_ = b_vec * (1.0 + 1.2e-16 * prng.standard_normal(b_vec.shape))
                try:
                    fold(run(bp_, lin(lambda z: O.mv(ops.A, z)), lin(M_model.matvec)))
                finally:
                    O.set_rounding_model(0.0)
            for h in extra:
                fold(h)
            return env

        def run_fg(bb, Aop, Mop):
            O.fgmres(Aop, bb, M=Mop, tol=1e-8, maxiter=150)
            return O.fgmres.last_history.copy()

        c_hists = []
        for threads in (1, 3, 8):
            c_oracle.set_threads(threads)
            c_hists.append(c_oracle.COracle(n, xi, eta_n, eta_s, c, d, **ckw).fgmres(b_vec, tol=1e-8, restart=150,
                                                                                     maxiter=150)[2])
        out["hist_sens"] = envelope(run_fg, _state["hist"], extra=c_hists)
        # left-preconditioned scipy gmres on the reference's dense A with the same closure
        for restart in (20, 150):
            xs, info, hs = O.gmres_scipy(A, b_vec, M=M, rtol=1e-8, restart=restart, maxiter=40)
            out[f"scipy_hist_r{restart}"] = hs
            out[f"scipy_x_r{restart}"] = xs
            out[f"scipy_info_r{restart}"] = np.array(info)
            out[f"scipy_sens_r{restart}"] = envelope(
                lambda bb, Aop, Mop: O.gmres_scipy(Aop, bb, M=Mop, rtol=1e-8, restart=restart, maxiter=40)[2], hs)
        print("  sens fgmres max", out["hist_sens"].max(), "scipy r20", out["scipy_sens_r20"].max(), "r150",
              out["scipy_sens_r150"].max())
        np.savez_compressed(os.path.join(HERE, f"solve_{tag}_n{n}_eta{int(eta_n)}.npz"), **out)
        print("solve", tag, n, eta_n, "fgmres its", len(_state["hist"]), "scipy its", len(out["scipy_hist_r20"]),
              len(out["scipy_hist_r150"]), "err norms", norms)

    # ---- known-answer tables: per-operator truncation errors (utils.py:42-157) and apply.py logic ----
    kat = {}
    for n in (8, 16, 32):
        bp = refpc.MultiphaseBlockPreconditioner(n, 1.0, 1.0, 1.0)
        L_n, D_n, XI_n, G_n = bp.get_block_matrices(is_ths=False)
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            refutils.check_individual_operators(n, 1.0, L_n, D_n, XI_n, G_n, True, True, True, True)
        vals = [float(v) for v in re.findall(r"_norm for n = \d+ is ([0-9.eE+-]+)", buf.getvalue())]
        # print order: D (L1,L2), G, XI, L
        kat[f"opcheck_n{n}"] = np.array(vals)
        for eta_n in (1.0, 100.0):
            A, S, F, D, G = refpc.MultiphaseBlockPreconditioner(n, 1.0, eta_n, 1.0).get_big_A_matrix(c=1.0, d_u=-1.0)
            _, b_vec, u_vec = refsolve.main(n=n, c=1, d=-1, xi=1.0, eta_n=eta_n, eta_s=1.0)
            bp_ = A @ u_vec
            w = (1 / n) * (1 / n)
            kat[f"apply_n{n}_eta{int(eta_n)}"] = np.array([refutils.weighted_L1(b_vec, bp_, w),
                                                         refutils.weighted_L2(b_vec, bp_, w),
                                                         refutils.max_norm(b_vec, bp_)])
        print("kat", n, kat[f"opcheck_n{n}"])
    # solve_with_exact_schur_pc (solve.py:210-238) run verbatim at n=8 (dense exact S; pyamg.fgmres -> oracle FGMRES)
    n8, c8, d8, xi8, en8, es8 = 8, 1, -1, 1.0, 100.0, 1.0
    A8, b8, u8 = refsolve.main(n=n8, c=c8, d=d8, xi=xi8, eta_n=en8, eta_s=es8)
    _state.clear()
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        refsolve.solve_with_exact_schur_pc(n8, xi8, en8, es8, c8, d8, b8, u8)
    vals = [float(v) for v in re.findall(r"_norm for n = \d+ is ([0-9.eE+-]+)", buf.getvalue())]
    kat["exact_schur_n8_norms"] = np.array(vals)          # direct solve (L1, L2, max), then fGMRES (L1, L2, max)
    kat["exact_schur_n8_hist"] = _state["hist"]
    kat["exact_schur_n8_S_fro"] = np.array(np.linalg.norm(refpc.MultiphaseBlockPreconditioner(n8, xi8, en8, es8).get_big_A_matrix(c=c8, d_u=d8)[1]))
    print("exact schur n=8:", vals, "its", len(_state["hist"]))
    # get_thn_vals (preconditioner.py:26-84) for every u-face of an 8 x 8 grid, both phases
    bp8 = refpc.MultiphaseBlockPreconditioner(8, 1.0, 1.0, 1.0)
    kat["thn_vals_n8"] = np.array([[[bp8.get_thn_vals(8, r, cc, bool(ph)) for cc in range(8)] for r in range(8)]
                                   for ph in (0, 1)])
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **kat)


if __name__ == "__main__":
    main()
