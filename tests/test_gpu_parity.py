"""Parity of the CUDA path (through the Python host -> C ABI -> kernels) against the CPU oracle and
the golden fixtures generated from the reference's own code.  fp64; tolerances are stated per test."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import mpbp_oracle as O
from conftest import golden, hist_check, relerr

pytestmark = pytest.mark.gpu

OPS_FIXTURES = ["ops_n4_eta1.npz", "ops_n8_eta100.npz", "ops_n16_eta100.npz", "ops_n12_eta3.npz", "ops_n16_eta10000.npz"]


def _bp(mp, params, **kw):
    n, xi, eta_n, eta_s, c, d = params
    return mp.MultiphaseBlockPreconditioner(int(n), xi, eta_n, eta_s, **kw), int(n), c, d


@pytest.mark.parametrize("fx", OPS_FIXTURES)
def test_operators_vs_reference_golden(mp, fx):
    """A, F, D, G, Gt_G, Gt_F_G applied to a seeded vector vs the reference's dense matmuls. tol 1e-13 rel."""
    g = golden(fx)
    bp, n, c, d = _bp(mp, g["params"])
    N = n * n
    A, S, F, D, G = bp.get_big_A_matrix(c=c, d_u=d)
    GtG, GtFG, _, _ = bp.derived_operators(c=c, d_u=d)
    x = g["x"]
    assert A.shape == (5 * N, 5 * N) and F.shape == (4 * N, 4 * N) and D.shape == (N, 4 * N) and G.shape == (4 * N, N)
    assert relerr(A @ x, g["Ax"]) < 1e-13
    assert relerr(F @ x[:4 * N], g["Fx"]) < 1e-13
    assert relerr(D @ x[:4 * N], g["Dx"]) < 1e-13
    assert relerr(G @ x[4 * N:], g["Gp"]) < 1e-13
    assert relerr(GtG @ x[4 * N:], g["GtGp"]) < 1e-13
    assert relerr(GtFG @ x[4 * N:], g["GtFGp"]) < 1e-12
    assert relerr(A @ g["u_vec"], g["Au"]) < 1e-13
    assert relerr(A @ g["b_vec"], g["Ab"]) < 1e-13


@pytest.mark.parametrize("fx", OPS_FIXTURES[:3])
def test_dense_matrix_entries_vs_reference(mp, fx):
    """Every entry of A (operator applied to the identity) vs the reference's dense A."""
    g = golden(fx)
    bp, n, c, d = _bp(mp, g["params"])
    A = bp.get_big_A_matrix(c=c, d_u=d)[0]
    ref = sp.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=tuple(g["A_shape"])).toarray()
    got = A.toarray()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-14
    assert abs(np.linalg.norm(got) - float(g["normA"])) / float(g["normA"]) < 1e-13


@pytest.mark.parametrize("fx", OPS_FIXTURES)
def test_manufactured_vectors(mp, fx):
    """solve.main's (u_vec, b_vec): host mirror and device kernel vs the reference's fill loop."""
    g = golden(fx)
    n, xi, eta_n, eta_s, c, d = g["params"]
    A, b_vec, u_vec = mp.main(n=int(n), c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
    assert relerr(b_vec, g["b_vec"]) < 1e-13 and relerr(u_vec, g["u_vec"]) < 1e-14
    A, b_dev, u_dev = mp.main(n=int(n), c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s, device_vectors=True)
    assert relerr(b_dev.cpu().numpy(), g["b_vec"]) < 1e-13
    assert relerr(u_dev.cpu().numpy(), g["u_vec"]) < 1e-13


@pytest.mark.parametrize("fx", OPS_FIXTURES)
def test_jacobi_vs_reference(mp, fx):
    """solve.Jacobi run verbatim by the reference (3 undamped sweeps from 0) on F and Gt_G. tol 1e-12."""
    g = golden(fx)
    bp, n, c, d = _bp(mp, g["params"])
    N = n * n
    F = bp.get_big_A_matrix(c=c, d_u=d)[2]
    GtG = bp.derived_operators(c=c, d_u=d)[0]
    bF = g["x"][:4 * N]
    assert relerr(mp.Jacobi(F, bF, 3, 0 * bF), g["jacF3"]) < 1e-12
    assert relerr(mp.Jacobi(GtG, g["bP"], 3, 0 * g["bP"]), g["jacP3"]) < 1e-12


@pytest.mark.parametrize("n", [4, 6, 8, 30, 33, 64, 100, 121, 256])
def test_operators_vs_oracle_sizes(mp, n):
    """Ragged sizes (not multiples of the 30-column warp tile, odd n, n smaller than a warp)."""
    rng = np.random.default_rng(n)
    xi, eta_n, eta_s, c, d, dp, dd = 0.8, 50.0, 2.0, 1.1, -0.9, 1.2, -1.1
    ops = O.Operators(n, xi, eta_n, eta_s, c, d, dp, dd)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s)
    p = bp.plan(c, d, dp, dd, operators_only=True)
    from mp_block_preconditioners_b200.preconditioner import (DivergenceOperator, GradientOperator, GtFGOperator,
                                                              GtGOperator, SystemOperator, VelocityOperator)
    N = n * n
    x = rng.standard_normal(5 * N)
    assert relerr(SystemOperator(p) @ x, ops.A @ x) < 1e-13
    assert relerr(VelocityOperator(p) @ x[:4 * N], ops.F @ x[:4 * N]) < 1e-13
    assert relerr(DivergenceOperator(p) @ x[:4 * N], ops.D @ x[:4 * N]) < 1e-13
    assert relerr(GradientOperator(p) @ x[4 * N:], ops.G @ x[4 * N:]) < 1e-13
    assert relerr(GtGOperator(p) @ x[4 * N:], ops.GtG @ x[4 * N:]) < 1e-13
    assert relerr(GtFGOperator(p) @ x[4 * N:], ops.GtFG @ x[4 * N:]) < 1e-12


@pytest.mark.parametrize("n,eta_n", [(16, 100.0), (32, 1.0), (64, 1e3), (48, 10.0)])
@pytest.mark.parametrize("cheb", [False, True])
def test_subsolvers_vs_oracle(mp, n, eta_n, cheb):
    """One V-cycle, the configured F~^-1 / (GtG)~^-1 and damped Jacobi vs the oracle. tol 1e-10 rel."""
    rng = np.random.default_rng(7)
    xi, eta_s, c, d = 1.0, 1.0, 1.0, -1.0
    sub = mp.SubSolver(kind="mg", F_cycles=3, P_cycles=2, cheb=cheb)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub)
    p = bp.plan(c, d)
    GtG, GtFG, Finv, Pinv = bp.derived_operators(c, d)
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    N = n * n
    bF = rng.standard_normal(4 * N)
    bP = rng.standard_normal(N)
    bP -= bP.mean()
    for kF, kP in ((1, 1), (3, 2)):
        cfgF = O.SubSolverConfig(kind="mg", cycles=kF, cheb=cheb)
        cfgP = O.SubSolverConfig(kind="mg", cycles=kP, cheb=cheb)
        mgF, mgP = O.Multigrid(ops, cfgF), O.Multigrid(ops, cfgP)
        if kF == 1:
            import torch
            yF = p.call("mpbp_vcycle_F", bF, 4 * N, 4 * N)
            yP = p.call("mpbp_vcycle_P", bP, N, N)
            refF, refP = mgF._vcycle("F", 0, bF), mgP._vcycle("P", 0, bP)
        else:
            yF, yP = Finv @ bF, Pinv @ bP
            refF, refP = mgF.solve("F", bF), mgP.solve("P", bP)
        assert relerr(yF, refF) < 1e-10, (kF, relerr(yF, refF))
        assert relerr(yP, refP) < 1e-10, (kP, relerr(yP, refP))
    F = bp.get_big_A_matrix(c, d)[2]
    assert relerr(mp.Jacobi(F, bF, 5, 0 * bF, omega=0.8), O.jacobi(ops.F, bF, 5, 0 * bF, 0.8)) < 1e-12
    assert relerr(mp.Jacobi(GtG, bP, 5, 0 * bP, omega=0.8), O.jacobi(ops.GtG, bP, 5, 0 * bP, 0.8)) < 1e-12


SOLVES = [("solve_mgcheb_n16_eta100.npz", dict(kind="mg", F_cycles=4, P_cycles=4, cheb=True)),
          ("solve_mgplain_n16_eta100.npz", dict(kind="mg", F_cycles=2, P_cycles=2, cheb=False)),
          ("solve_jacobi_n16_eta100.npz", dict(kind="jacobi", F_sweeps=20, P_sweeps=20, omega=0.8)),
          ("solve_mgcheb_n32_eta1.npz", dict(kind="mg", F_cycles=4, P_cycles=4, cheb=True))]


@pytest.mark.parametrize("fx,subkw", SOLVES)
def test_precond_apply_vs_reference_closure(mp, fx, subkw):
    """z = M v vs the reference's verbatim approx_schur_op closure (solve.py:257-277). tol 1e-9 rel."""
    g = golden(fx)
    bp, n, c, d = _bp(mp, g["params"], sub_solver=mp.SubSolver(**subkw))
    M = bp.approx_schur_operator(c=c, d_u=d)
    v = g["v"].copy()
    z = M @ v
    assert np.array_equal(v, g["v"])  # input untouched
    assert z.shape == v.shape and z is not v
    assert relerr(z, g["Mv"]) < 1e-9, relerr(z, g["Mv"])
    assert relerr(M @ g["b_vec"], g["Mb"]) < 1e-9
    assert relerr(M.matvec_host(v), g["Mv"]) < 1e-9


@pytest.mark.parametrize("fx,subkw", SOLVES)
def test_fgmres_history_vs_reference_run(mp, fx, subkw):
    """Right-preconditioned FGMRES (solve.py:285) residual history, iterate and error norms vs the golden
    run of the reference's solve_with_approx_schur_pc.  History within 1e-10 relative, count +-1."""
    g = golden(fx)
    n, xi, eta_n, eta_s, c, d = g["params"]
    n = int(n)
    u, info, hist = mp.solve_with_approx_schur_pc(n, xi, eta_n, eta_s, c, d, g["b_vec"], g["u_vec"],
                                                  sub_solver=mp.SubSolver(**subkw), verbose=False)
    assert info == 0
    hist_check(hist, g["hist"], g["hist_sens"], label=f"fgmres {fx}")
    # both runs stop at ||r|| < 1e-8 ||b||; the iterates agree to that accuracy times the conditioning
    assert relerr(u, g["x"]) < 1e-5
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    assert np.linalg.norm(g["b_vec"] - ops.A @ u) < 1.01e-8 * np.linalg.norm(g["b_vec"])
    w = (1 / n) * (1 / n)
    got = [mp.weighted_L1(u, g["u_vec"], w), mp.weighted_L2(u, g["u_vec"], w), mp.max_norm(u, g["u_vec"])]
    assert np.allclose(got, g["err_norms"], rtol=1e-6)


@pytest.mark.parametrize("fx,subkw", SOLVES)
@pytest.mark.parametrize("restart", [20, 150])
def test_gmres_left_history_vs_scipy(mp, fx, subkw, restart):
    """scipy.sparse.linalg.gmres semantics (left preconditioning): pr_norm history vs scipy run on the
    reference's dense A with its verbatim preconditioner closure. 1e-10 relative, count +-1."""
    g = golden(fx)
    bp, n, c, d = _bp(mp, g["params"], sub_solver=mp.SubSolver(**subkw))
    A = bp.get_big_A_matrix(c=c, d_u=d)[0]
    M = bp.approx_schur_operator(c=c, d_u=d)
    hist = []
    x, info = mp.gmres(A, g["b_vec"], M=M, rtol=1e-8, restart=restart, maxiter=40, callback=hist.append,
                       callback_type="pr_norm")
    ref = g[f"scipy_hist_r{restart}"]
    assert info == int(g[f"scipy_info_r{restart}"])
    hist_check(hist, ref, g[f"scipy_sens_r{restart}"], label=f"gmres_left r{restart} {fx}")
    assert relerr(x, g[f"scipy_x_r{restart}"]) < 1e-5


def test_true_residual_callback(mp):
    """print_true_res_norm (solve.py:161-170) through the fgmres callback verification mode."""
    g = golden("solve_mgcheb_n16_eta100.npz")
    n, xi, eta_n, eta_s, c, d = g["params"]
    sub = mp.SubSolver(kind="mg", F_cycles=4, P_cycles=4, cheb=True)
    bp = mp.MultiphaseBlockPreconditioner(int(n), xi, eta_n, eta_s, sub_solver=sub)
    A = bp.get_big_A_matrix(c=c, d_u=d)[0]
    M = bp.approx_schur_operator(c=c, d_u=d)
    out = []
    x, info = mp.fgmres(A, g["b_vec"], M=M, tol=1e-8, maxiter=150, callback=mp.print_true_res_norm(A, g["b_vec"], out, verbose=False))
    assert len(out) == len(g["true_res"])
    out = np.array(out)
    # the true residual ||b - A x_k|| carries an O(eps * cond) floor the recurrence residual does not: 1e-6 strict level
    hist_check(out, g["true_res"], g["hist_sens"][:len(out)], label="true residual", strict_rel=1e-6)
    # the recurrence residual the solver reports tracks the true residual (right preconditioning)
    assert np.allclose(out, mp.fgmres.last_history, rtol=1e-3)


def test_unpreconditioned_matches_oracle(mp):
    """solve_without_pc (solve.py:202-208): M=None, 100 iterations, does not converge; history parity."""
    n, xi, eta_n, eta_s, c, d = 16, 1.0, 100.0, 1.0, 1, -1
    A, b_vec, u_vec = mp.main(n=n, c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
    x, info = mp.solve_without_pc(n, A, b_vec, u_vec, verbose=False)
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    xo, info_o = O.fgmres(ops.A, b_vec, M=None, x0=np.zeros(5 * n * n), tol=1e-8, maxiter=100)
    assert info != 0 and info_o != 0
    h, ho = mp.fgmres.last_history, O.fgmres.last_history.copy()
    assert len(h) == len(ho) == 100
    # conditioning of the oracle's own history: 16 re-runs under the rounding model (every A.x returns
    # y_i + u sqrt(k_i) (|A||x|)_i N(0,1), oracle/mpbp_oracle.py:mv) with 1-ulp perturbations of b
    import scipy.sparse.linalg as spla
    sens = np.zeros(100)
    prng = np.random.default_rng(5)
    Aop = spla.LinearOperator(ops.A.shape, dtype=np.float64, matvec=lambda z: O.mv(ops.A, z))
    for i in range(16):
        O.set_rounding_model(1.1e-16, seed=500 + i)
        try:
            O.fgmres(Aop, b_vec * (1 + 1.2e-16 * prng.standard_normal(b_vec.shape)), M=None, tol=1e-8, maxiter=100)
        finally:
            O.set_rounding_model(0.0)
        sens = np.maximum(sens, np.abs(O.fgmres.last_history - ho) / ho)
    hist_check(h, ho, sens, label="unpreconditioned")
    assert np.allclose(h[:20], ho[:20], rtol=1e-9)


def test_known_answer_operator_checks(mp):
    """utils.check_individual_operators (utils.py:42-157) on the GPU blocks of get_block_matrices: L1/L2 truncation
    errors of D, G, XI, L for n = 8, 16, 32 vs the values printed by the reference (second order)."""
    kat = golden("known_answers.npz")
    for n in (8, 16, 32):
        bp = mp.MultiphaseBlockPreconditioner(n, 1.0, 1.0, 1.0)
        L, D, XI, G = bp.get_block_matrices(is_ths=False)
        res = mp.check_individual_operators(n, 1.0, L, D, XI, G, True, True, True, True, verbose=False)
        got = [v for k in ("D", "G", "XI", "L") for v in res[k]]
        assert np.allclose(got, kat[f"opcheck_n{n}"], rtol=1e-7), (n, got, kat[f"opcheck_n{n}"])


def test_exact_schur_small_n_path(mp):
    """solve_with_exact_schur_pc (solve.py:210-238) with the dense S = -D F^-1 G built from the GPU
    operators (preconditioner.py:343-346), vs the reference's own run at n=8: ||S||_F, the error norms of the direct
    block-LU solve and of the fGMRES(maxiter=40) solve, and the residual history."""
    kat = golden("known_answers.npz")
    if "exact_schur_n8_norms" not in kat.files:
        pytest.skip("fixture predates the exact-Schur path")
    n, c, d, xi, eta_n, eta_s = 8, 1, -1, 1.0, 100.0, 1.0
    A, b_vec, u_vec = mp.main(n=n, c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
    S = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s).get_big_A_matrix(c=c, d_u=d)[1]
    assert S.shape == (n * n, n * n)
    assert abs(np.linalg.norm(S.toarray()) - float(kat["exact_schur_n8_S_fro"])) < 1e-9 * float(kat["exact_schur_n8_S_fro"])
    u_direct, u_gmres, info, hist = mp.solve_with_exact_schur_pc(n, xi, eta_n, eta_s, c, d, b_vec, u_vec, verbose=False)
    w = (1 / n) ** 2
    got = [mp.weighted_L1(u_direct, u_vec, w), mp.weighted_L2(u_direct, u_vec, w), mp.max_norm(u_direct, u_vec),
           mp.weighted_L1(u_gmres, u_vec, w), mp.weighted_L2(u_gmres, u_vec, w), mp.max_norm(u_gmres, u_vec)]
    assert np.allclose(got, kat["exact_schur_n8_norms"], rtol=1e-4), (got, kat["exact_schur_n8_norms"])
    ref = kat["exact_schur_n8_hist"]
    assert abs(len(hist) - len(ref)) <= 1
    k = min(len(hist), len(ref), 4)
    assert np.allclose(hist[:k], ref[:k], rtol=1e-3)
    with pytest.raises(NotImplementedError):
        mp.MultiphaseBlockPreconditioner(64, 1.0, 1.0, 1.0).get_big_A_matrix(1.0, -1.0)[1].toarray()


def test_apply_check_known_answers(mp):
    """apply.py logic (A u_exact vs b_exact; apply.py:71-81): second-order errors vs the reference's values."""
    from mp_block_preconditioners_b200.apply import apply_check
    kat = golden("known_answers.npz")
    for n in (8, 16, 32):
        for eta_n in (1.0, 100.0):
            got = apply_check(n=n, xi=1.0, eta_n=eta_n, eta_s=1.0, c=1.0, d=-1.0, verbose=False)
            assert np.allclose(got, kat[f"apply_n{n}_eta{int(eta_n)}"], rtol=1e-9), (n, eta_n, got)


def test_blas1_kernels(mp):
    import ctypes as C
    import torch
    bp = mp.MultiphaseBlockPreconditioner(64, 1.0, 1.0, 1.0)
    p = bp.plan(1.0, -1.0, operators_only=True)
    lib = p.lib
    from mp_block_preconditioners_b200._cabi import check
    for length in (1, 31, 1000, 5 * 64 * 64, 1_000_003):
        g = torch.Generator(device="cuda").manual_seed(length)
        x = torch.randn(length, dtype=torch.float64, device="cuda", generator=g)
        y = torch.randn(length, dtype=torch.float64, device="cuda", generator=g)
        V = torch.randn(11, length, dtype=torch.float64, device="cuda", generator=g)
        r = C.c_double()
        check(lib.mpbp_dot(p.h, x.data_ptr(), y.data_ptr(), length, C.byref(r), p.stream()))
        ref = float((x.cpu().numpy() * y.cpu().numpy()).sum())
        assert abs(r.value - ref) <= 1e-12 * float((x.abs() * y.abs()).sum())
        check(lib.mpbp_nrm2(p.h, x.data_ptr(), length, C.byref(r), p.stream()))
        assert abs(r.value - float(torch.linalg.norm(x))) <= 1e-13 * r.value
        out = (C.c_double * 11)()
        check(lib.mpbp_multi_dot(p.h, V.data_ptr(), length, 11, x.data_ptr(), length, out, p.stream()))
        refm = (V @ x).cpu().numpy()
        assert np.allclose(np.array(out[:]), refm, rtol=0, atol=1e-12 * float((V.abs() @ x.abs()).max()))
        al = np.linspace(-1, 1, 11)
        y2 = y.clone()
        check(lib.mpbp_multi_axpy(p.h, V.data_ptr(), length, 11, al.ctypes.data_as(C.POINTER(C.c_double)),
                                  y2.data_ptr(), length, p.stream()))
        refy = y + torch.from_numpy(al).cuda() @ V
        assert float((y2 - refy).abs().max()) <= 1e-13 * float(refy.abs().max() + 1)
        y3 = y.clone()
        check(lib.mpbp_axpy(p.h, 0.37, x.data_ptr(), y3.data_ptr(), length, p.stream()))
        assert float((y3 - (y + 0.37 * x)).abs().max()) <= 1e-15 * float(y.abs().max() + 1)
        out3 = (C.c_double * 3)()
        check(lib.mpbp_wnorms(p.h, x.data_ptr(), y.data_ptr(), length, 0.25, out3, p.stream()))
        q = (x - y).abs()
        assert np.allclose(out3[:], [0.25 * float(q.sum()), float(torch.sqrt(0.25 * (q * q).sum())), float(q.max())], rtol=1e-12)
    # determinism: identical bits run to run
    x = torch.randn(5 * 64 * 64, dtype=torch.float64, device="cuda")
    vals = set()
    for _ in range(5):
        r = C.c_double()
        check(lib.mpbp_dot(p.h, x.data_ptr(), x.data_ptr(), x.numel(), C.byref(r), p.stream()))
        vals.add(r.value)
    assert len(vals) == 1


def test_user_supplied_theta_field(mp):
    """General (non-analytic) cell-centred theta_n input: the mass term then uses two-cell face
    averages.  Operators, sub-solves and the preconditioner vs the oracle built from the same field."""
    n, xi, eta_n, eta_s, c, d = 32, 0.9, 20.0, 1.5, 1.2, -1.0
    r = (np.arange(n) + 0.5)[:, None] / n
    cc = (np.arange(n) + 0.5)[None, :] / n
    theta = 0.5 + 0.3 * np.sin(2 * np.pi * (cc + 2 * r)) * np.cos(4 * np.pi * cc) + 0.05 * np.cos(6 * np.pi * r)
    ops = O.Operators(n, xi, eta_n, eta_s, c, d, theta=theta, mass="average")
    sub = mp.SubSolver(kind="mg", F_cycles=3, P_cycles=2, cheb=True)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub, theta=theta)
    A, S, F, D, G = bp.get_big_A_matrix(c, d)
    GtG, GtFG, Finv, Pinv = bp.derived_operators(c, d)
    M = bp.approx_schur_operator(c, d)
    rng = np.random.default_rng(11)
    N = n * n
    x = rng.standard_normal(5 * N)
    x[4 * N:] -= x[4 * N:].mean()
    assert relerr(A @ x, ops.A @ x) < 1e-13
    assert relerr(GtFG @ x[4 * N:], ops.GtFG @ x[4 * N:]) < 1e-12
    Mo = O.ApproxSchur(ops, O.SubSolverConfig(kind="mg", cycles=3, cheb=True))
    cfgP = O.SubSolverConfig(kind="mg", cycles=2, cheb=True)
    Mo.P_inv = O.SubSolver(ops, "P", cfgP, O.Multigrid(ops, cfgP))
    assert relerr(Finv @ x[:4 * N], Mo.F_inv @ x[:4 * N]) < 1e-10
    assert relerr(Pinv @ x[4 * N:], Mo.P_inv @ x[4 * N:]) < 1e-10
    assert relerr(M @ x, Mo.matvec(x)) < 1e-9


def test_fused_mgs_and_graph_replay_are_bitwise_neutral(mp, monkeypatch):
    """The fused Gram-Schmidt kernel and the CUDA-graph replay of V-cycles are pure scheduling changes:
    residual histories and iterates must be bit-identical to the unfused / eager execution."""
    n, xi, eta_n, eta_s, c, d = 64, 1.0, 100.0, 1.0, 1, -1
    res = {}
    for tag, env in (("default", {}), ("plain", {"MPBP_FUSED_MGS": "0", "MPBP_GRAPH": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        A, b_vec, u_vec = mp.main(n=n, c=c, d=d, xi=xi, eta_n=eta_n, eta_s=eta_s)
        bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s)
        A = bp.get_big_A_matrix(c=c, d_u=d)[0]
        M = bp.approx_schur_operator(c=c, d_u=d)
        xr, ir = mp.fgmres(A, b_vec, M=M, tol=1e-8, maxiter=60)
        hr = mp.fgmres.last_history.copy()
        xl, il = mp.gmres(A, b_vec, M=M, rtol=1e-8, restart=20, maxiter=10)
        hl = mp.gmres.last_history.copy()
        res[tag] = (xr, hr, xl, hl, ir, il)
        for k in env:
            monkeypatch.delenv(k)
    a, b = res["default"], res["plain"]
    assert a[4] == b[4] == 0
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("fuse", ["1", "2", "3", "4", "8", "0"])
@pytest.mark.parametrize("n,eta_n", [(64, 1e3), (48, 10.0), (16, 100.0)])
def test_fused_smoothing_kernels_vs_oracle(mp, monkeypatch, fuse, n, eta_n):
    """MPBP_FUSE bit 0: the two pre-smoothing sweeps in one kernel; bit 1: prolongation + first post-sweep in
    one kernel; bit 2: residual + restriction; bit 3: the same three fusions on the pressure-Poisson cycle
    (csrc/poisson.cuh); default 15.  Same V-cycles, so the oracle tolerances of the unfused path (0) apply."""
    monkeypatch.setenv("MPBP_FUSE", fuse)
    xi, eta_s, c, d = 1.0, 1.0, 1.0, -1.0
    sub = mp.SubSolver(kind="mg", F_cycles=3, P_cycles=2, cheb=True)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub)
    p = bp.plan(c, d)
    monkeypatch.delenv("MPBP_FUSE")
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    rng = np.random.default_rng(3)
    N = n * n
    v = rng.standard_normal(5 * N)
    v[4 * N:] -= v[4 * N:].mean()
    mg1 = O.Multigrid(ops, O.SubSolverConfig(kind="mg", cycles=1))
    assert relerr(p.call("mpbp_vcycle_F", v[:4 * N], 4 * N, 4 * N), mg1._vcycle("F", 0, v[:4 * N])) < 1e-10
    assert relerr(p.call("mpbp_vcycle_P", v[4 * N:], N, N), mg1._vcycle("P", 0, v[4 * N:])) < 1e-10
    Mo = O.ApproxSchur(ops, O.SubSolverConfig(kind="mg", cycles=3, cheb=True))
    cfgP = O.SubSolverConfig(kind="mg", cycles=2, cheb=True)
    Mo.P_inv = O.SubSolver(ops, "P", cfgP, O.Multigrid(ops, cfgP))
    M = bp.approx_schur_operator(c, d)
    assert relerr(M @ v, Mo.matvec(v)) < 1e-9


def test_fused_pressure_cycle_matches_the_unfused_one_to_rounding(mp, monkeypatch):
    """MPBP_FUSE bit 3 (csrc/poisson.cuh): the fused pressure-Poisson kernels perform the arithmetic of the passes they
    replace in the same order (1/diag is stored exactly as the sweeps compute it; bit-identical on the CPU shim).  On the
    device nvcc contracts multiply-adds per instantiation, so the V-cycle and the configured (GtG)~^-1 agree to
    rounding (1e-13 of the result's scale), not bit for bit."""
    n = 96
    rng = np.random.default_rng(9)
    v = rng.standard_normal(n * n)
    v -= v.mean()
    out = {}
    for fuse in ("7", "15"):
        monkeypatch.setenv("MPBP_FUSE", fuse)
        bp = mp.MultiphaseBlockPreconditioner(n, 1.0, 1e3, 1.0, sub_solver=mp.SubSolver(kind="mg", F_cycles=2, P_cycles=3, cheb=True))
        p = bp.plan(1.0, -1.0)
        monkeypatch.delenv("MPBP_FUSE")
        GtG, GtFG, Finv, Pinv = bp.derived_operators(1.0, -1.0)
        out[fuse] = (p.call("mpbp_vcycle_P", v, n * n, n * n), Pinv @ v)
    assert relerr(out["15"][0], out["7"][0]) < 1e-13 and relerr(out["15"][1], out["7"][1]) < 1e-13


@pytest.mark.parametrize("cell", ["0", "64", "512"])
@pytest.mark.parametrize("n,eta_n,n_coarse", [(128, 1e3, 4), (96, 10.0, 4), (256, 1e4, 16)])
def test_cell_parallel_small_level_kernels_vs_oracle(mp, monkeypatch, cell, n, eta_n, n_coarse):
    """MPBP_CELL: whole-grid levels below level 0 with n <= MPBP_CELL run the cell-parallel kernels of csrc/cell.cuh
    (one thread per cell) instead of the marching kernels; 0 switches them off.  Same V-cycle: the oracle tolerances
    apply for every setting, and the settings agree with each other far below them.  n_coarse = 16 is the benchmarked
    hierarchy (dense solve on the 16 x 16 grid, warp-per-row kernel)."""
    monkeypatch.setenv("MPBP_CELL", cell)
    xi, eta_s, c, d = 1.0, 1.0, 1.0, -1.0
    sub = mp.SubSolver(kind="mg", F_cycles=3, P_cycles=2, cheb=True, n_coarse=n_coarse)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub)
    p = bp.plan(c, d)
    monkeypatch.delenv("MPBP_CELL")
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    rng = np.random.default_rng(5)
    N = n * n
    v = rng.standard_normal(5 * N)
    v[4 * N:] -= v[4 * N:].mean()
    mg1 = O.Multigrid(ops, O.SubSolverConfig(kind="mg", cycles=1, n_coarse=n_coarse))
    # two correct dense inverses of the coarsest velocity block (Gauss-Jordan here, LAPACK in the oracle) differ by
    # eps * cond, cond ~ 260 * eta_n/eta_s * (n_coarse/4)^2 (cf. tests/test_c_oracle.py): measured 2.2e-10 at 1e4, 16
    tolF = max(1e-10, 2e-13 * eta_n * (n_coarse / 4) ** 2)
    assert relerr(p.call("mpbp_vcycle_F", v[:4 * N], 4 * N, 4 * N), mg1._vcycle("F", 0, v[:4 * N])) < tolF
    Mo = O.ApproxSchur(ops, O.SubSolverConfig(kind="mg", cycles=3, cheb=True, n_coarse=n_coarse))
    cfgP = O.SubSolverConfig(kind="mg", cycles=2, cheb=True, n_coarse=n_coarse)
    Mo.P_inv = O.SubSolver(ops, "P", cfgP, O.Multigrid(ops, cfgP))
    M = bp.approx_schur_operator(c, d)
    assert relerr(M @ v, Mo.matvec(v)) < 1e-9


@pytest.mark.skipif(os.environ.get("MPBP_TEST_EXPERIMENTAL") != "1",
                    reason="experimental persistent coarse V-cycle kernel (off by default); set MPBP_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("n,eta_n", [(64, 1e3), (128, 10.0)])
def test_experimental_persistent_coarse_vcycle(mp, monkeypatch, n, eta_n):
    """MPBP_COARSE=64: every level with n <= 64 runs inside one single-CTA kernel (csrc/coarse.cuh).
    Logic-checked on the CPU shim (tests/test_emu_kernels.py); this is its GPU parity test."""
    monkeypatch.setenv("MPBP_COARSE", "64")
    xi, eta_s, c, d = 1.0, 1.0, 1.0, -1.0
    sub = mp.SubSolver(kind="mg", F_cycles=3, P_cycles=2, cheb=True)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub)
    p = bp.plan(c, d)
    monkeypatch.delenv("MPBP_COARSE")
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    rng = np.random.default_rng(5)
    N = n * n
    v = rng.standard_normal(5 * N)
    v[4 * N:] -= v[4 * N:].mean()
    mg1 = O.Multigrid(ops, O.SubSolverConfig(kind="mg", cycles=1))
    assert relerr(p.call("mpbp_vcycle_F", v[:4 * N], 4 * N, 4 * N), mg1._vcycle("F", 0, v[:4 * N])) < 1e-10
    assert relerr(p.call("mpbp_vcycle_P", v[4 * N:], N, N), mg1._vcycle("P", 0, v[4 * N:])) < 1e-10
    Mo = O.ApproxSchur(ops, O.SubSolverConfig(kind="mg", cycles=3, cheb=True))
    cfgP = O.SubSolverConfig(kind="mg", cycles=2, cheb=True)
    Mo.P_inv = O.SubSolver(ops, "P", cfgP, O.Multigrid(ops, cfgP))
    M = bp.approx_schur_operator(c, d)
    assert relerr(M @ v, Mo.matvec(v)) < 1e-9


def test_error_behaviour(mp):
    from mp_block_preconditioners_b200._cabi import MpbpError
    bp = mp.MultiphaseBlockPreconditioner(16, 1.0, 1.0, 1.0)
    A = bp.get_big_A_matrix(1.0, -1.0)[0]
    with pytest.raises(ValueError):
        A @ np.zeros(7)  # dimension mismatch, as np.matmul would raise
    with pytest.raises(MpbpError):
        mp.MultiphaseBlockPreconditioner(1, 1.0, 1.0, 1.0).get_big_A_matrix(1.0, -1.0)
    with pytest.raises(MpbpError):
        mp.MultiphaseBlockPreconditioner(4094, 1.0, 1.0, 1.0).get_big_A_matrix(1.0, -1.0)  # cannot coarsen
    p = bp.plan(1.0, -1.0, operators_only=True)
    with pytest.raises(MpbpError):
        p.call("mpbp_solve_F", np.zeros(4 * 256), 4 * 256, 4 * 256)
    with pytest.raises(TypeError):
        mp.fgmres(np.eye(3), np.zeros(3))


@pytest.mark.parametrize("n", [2048, 4096])
def test_full_size_properties(mp, n):
    """Size-independent properties at BASELINE.json's grid sizes (the oracle cannot run there):
    symmetry of A, G^T = -D, linearity, F positive, GtG 1 = 0, sum(D w) = 0, M linear."""
    import torch
    eta_n = 1e3 if n == 2048 else 1e4
    bp = mp.MultiphaseBlockPreconditioner(n, 1.0, eta_n, 1.0, sub_solver=mp.SubSolver(F_cycles=1, P_cycles=1, cheb=False))
    A, S, F, D, G = bp.get_big_A_matrix(1.0, -1.0)
    GtG = bp.derived_operators(1.0, -1.0)[0]
    N = n * n
    gen = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(5 * N, dtype=torch.float64, device="cuda", generator=gen)
    y = torch.randn(5 * N, dtype=torch.float64, device="cuda", generator=gen)
    Ax, Ay = A @ x, A @ y
    s1, s2 = float(torch.dot(y, Ax)), float(torch.dot(x, Ay))
    scale = float(torch.linalg.norm(y) * torch.linalg.norm(Ax))
    assert abs(s1 - s2) < 1e-12 * scale                      # A symmetric (Gt = -D, F symmetric)
    assert float(torch.linalg.norm(A @ (2.0 * x - 3.0 * y) - (2.0 * Ax - 3.0 * Ay))) < 1e-12 * float(torch.linalg.norm(Ax))
    w, pr = x[:4 * N], y[4 * N:]
    Gp, Dw = G @ pr, D @ w
    assert abs(float(torch.dot(Gp, w)) + float(torch.dot(pr, Dw))) < 1e-12 * float(torch.linalg.norm(Gp) * torch.linalg.norm(w))
    assert abs(float(Dw.sum())) < 1e-9 * float(Dw.abs().sum())   # discrete divergence theorem
    assert float(torch.dot(w, F @ w)) > 0                        # F SPD
    ones = torch.ones(N, dtype=torch.float64, device="cuda")
    assert float((GtG @ ones).abs().max()) < 1e-6 * n * n        # constants are the null space
    # block structure: A [w;0] = [F w; -D w]
    z = torch.cat([w, torch.zeros(N, dtype=torch.float64, device="cuda")])
    Az = A @ z
    assert float((Az[:4 * N] - F @ w).abs().max()) <= 1e-13 * float((F @ w).abs().max())
    assert float((Az[4 * N:] + Dw).abs().max()) <= 1e-13 * float(Dw.abs().max())
    if n == 2048:
        M = bp.approx_schur_operator(1.0, -1.0)
        x2 = x.clone(); y2 = y.clone()
        x2[4 * N:] -= x2[4 * N:].mean(); y2[4 * N:] -= y2[4 * N:].mean()
        Mx, My = M @ x2, M @ y2
        lin = M @ (x2 + 0.5 * y2)
        assert float(torch.linalg.norm(lin - (Mx + 0.5 * My))) < 1e-9 * float(torch.linalg.norm(Mx))


def test_spectral_diagnostics_from_the_hessenberg(mp):
    """Ritz values of A M^-1 from the GMRES Hessenberg replace the reference's dense
    compute_preconditioned_A + SLEPc analysis (solve.py:103-200, :304-309).  (i) the GPU's Hessenberg matrix equals the
    oracle's FGMRES Hessenberg on a well-conditioned configuration (Jacobi sub-solves), 1e-8 relative; (ii) at n=8 the
    Ritz values lie in the convex hull of the dense spectrum of A M^-1 and the largest one has converged to its
    largest eigenvalue (what SLEPc's default largest-magnitude EPS would return, solve.py:121-123)."""
    g = golden("solve_jacobi_n16_eta100.npz")
    n, xi, eta_n, eta_s, c, d = g["params"]
    n = int(n)
    subkw = dict(kind="jacobi", F_sweeps=20, P_sweeps=20, omega=0.8)
    bp = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=mp.SubSolver(**subkw))
    A = bp.get_big_A_matrix(c=c, d_u=d)[0]
    M = bp.approx_schur_operator(c=c, d_u=d)
    mp.fgmres(A, g["b_vec"], M=M, tol=1e-8, maxiter=150)
    H = mp.last_hessenberg(A)
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    Mo = O.ApproxSchur(ops, O.SubSolverConfig(kind="jacobi", sweeps=20, omega=0.8))
    O.fgmres(ops.A, g["b_vec"], M=Mo.linear_operator(), tol=1e-8, maxiter=150)
    Ho = O.fgmres.last_hessenberg
    assert H.shape == Ho.shape == (len(g["hist"]) + 1, len(g["hist"]))
    assert np.abs(H - Ho).max() < 1e-8 * np.abs(Ho).max()
    sd = mp.spectral_diagnostics(A)
    ritz_o = np.linalg.eigvals(Ho[:-1])
    assert np.abs(np.sort_complex(sd["ritz"]) - np.sort_complex(ritz_o)).max() < 1e-6 * np.abs(ritz_o).max()
    # small dense check of what the Ritz values mean
    n = 8
    bp = mp.MultiphaseBlockPreconditioner(n, 1.0, 100.0, 1.0, sub_solver=mp.SubSolver(kind="mg", F_cycles=4, P_cycles=4, cheb=True))
    A = bp.get_big_A_matrix(1.0, -1.0)[0]
    M = bp.approx_schur_operator(1.0, -1.0)
    Ad = A.toarray()
    Md = np.column_stack([M @ e for e in np.eye(5 * n * n)])
    ev = np.linalg.eigvals(Ad @ Md)
    rng = np.random.default_rng(3)
    mp.fgmres(A, rng.standard_normal(5 * n * n), M=M, tol=1e-14, maxiter=60)
    sd = mp.spectral_diagnostics(A)
    assert sd["k"] >= 20
    assert sd["ritz"].real.max() <= ev.real.max() * (1 + 1e-6) and sd["ritz"].real.min() >= ev.real.min() - 1e-6 * abs(ev).max()
    assert abs(abs(sd["largest"][0]) - abs(ev).max()) < 1e-3 * abs(ev).max()


def test_iterate_callback_is_one_pass_per_iteration(mp):
    """fgmres(callback=) hands out x_k = x0 + Z y_k each inner iteration (pyamg semantics, solve.py:285, :163-169): the
    last iterate is the returned solution, every iterate's true residual tracks the recurrence residual, and the
    whole callback run launches only a few kernels more than the plain solve (no re-solves)."""
    g = golden("solve_mgcheb_n32_eta1.npz")
    n, xi, eta_n, eta_s, c, d = g["params"]
    bp = mp.MultiphaseBlockPreconditioner(int(n), xi, eta_n, eta_s, sub_solver=mp.SubSolver(kind="mg", F_cycles=4, P_cycles=4, cheb=True))
    A = bp.get_big_A_matrix(c=c, d_u=d)[0]
    M = bp.approx_schur_operator(c=c, d_u=d)
    l0 = A.plan.launches
    x_plain, _ = mp.fgmres(A, g["b_vec"], M=M, tol=1e-8, maxiter=150)
    l1 = A.plan.launches
    its = []
    x_cb, _ = mp.fgmres(A, g["b_vec"], M=M, tol=1e-8, maxiter=150, callback=lambda xk: its.append(xk.copy()))
    l2 = A.plan.launches
    assert len(its) == len(mp.fgmres.last_history)
    assert np.array_equal(x_cb, x_plain) and relerr(its[-1], x_plain) < 1e-12
    bn = np.linalg.norm(g["b_vec"])
    true = np.array([np.linalg.norm(g["b_vec"] - A @ xk) / bn for xk in its])
    assert np.allclose(true, mp.fgmres.last_history, rtol=1e-3)
    assert (l2 - l1) - (l1 - l0) <= 4 * len(its)
