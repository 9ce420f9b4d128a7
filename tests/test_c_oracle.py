"""CPU tests: the OpenMP C restatement (oracle/mpbp_oracle_c.c, written with the reference's coefficient
table) against the numpy oracle and the reference's golden vectors.  It is the second, independent
oracle and the multi-threaded CPU baseline of bench.py."""
import numpy as np
import pytest

import mpbp_oracle as O
from c_oracle import COracle
from conftest import golden, relerr


@pytest.mark.parametrize("fx", ["ops_n4_eta1.npz", "ops_n8_eta100.npz", "ops_n16_eta100.npz", "ops_n12_eta3.npz",
                                "ops_n16_eta10000.npz"])
def test_c_oracle_operators_vs_reference_golden(fx):
    g = golden(fx)
    n, xi, eta_n, eta_s, c, d = g["params"]
    n = int(n)
    N = n * n
    co = COracle(n, xi, eta_n, eta_s, c, d)
    x = g["x"]
    assert relerr(co.apply_A(x), g["Ax"]) < 1e-13
    assert relerr(co.apply_F(x[:4 * N]), g["Fx"]) < 1e-13
    assert relerr(co.apply_D(x[:4 * N]), g["Dx"]) < 1e-14
    assert relerr(co.apply_G(x[4 * N:]), g["Gp"]) < 1e-14
    assert relerr(co.apply_GtG(x[4 * N:]), g["GtGp"]) < 1e-13
    assert relerr(co.apply_A(g["u_vec"]), g["Au"]) < 1e-13


@pytest.mark.parametrize("n,eta_n,cheb,n_coarse", [(16, 100.0, True, 4), (32, 1.0, False, 4), (48, 1e3, True, 4),
                                                   (64, 1e4, True, 4), (64, 1e4, True, 16)])
def test_c_oracle_subsolvers_and_preconditioner_vs_numpy_oracle(n, eta_n, cheb, n_coarse):
    """n_coarse = 16 is the benchmarked hierarchy (bench.py: dense solve on the 16x16 grid)."""
    xi, eta_s, c, d = 1.0, 1.0, 1.0, -1.0
    ops = O.Operators(n, xi, eta_n, eta_s, c, d)
    cfgF = O.SubSolverConfig(kind="mg", cycles=3, cheb=cheb, n_coarse=n_coarse)
    cfgP = O.SubSolverConfig(kind="mg", cycles=2, cheb=cheb, n_coarse=n_coarse)
    Mo = O.ApproxSchur(ops, cfgF)
    Mo.P_inv = O.SubSolver(ops, "P", cfgP, O.Multigrid(ops, cfgP))
    co = COracle(n, xi, eta_n, eta_s, c, d, F_cycles=3, P_cycles=2, cheb=cheb, n_coarse=n_coarse)
    rng = np.random.default_rng(n)
    N = n * n
    v = rng.standard_normal(5 * N)
    v[4 * N:] -= v[4 * N:].mean()
    mg1 = O.Multigrid(ops, O.SubSolverConfig(kind="mg", cycles=1, n_coarse=n_coarse))
    # the coarsest 4x4 velocity block has condition number ~ 260 * eta_n/eta_s: two correct dense inverses
    # (Gauss-Jordan here, LAPACK in numpy) differ by eps * cond, which bounds the agreement of the F solves
    tolF = max(1e-10, 2e-13 * eta_n)
    assert relerr(co.vcycle("F", v[:4 * N]), mg1._vcycle("F", 0, v[:4 * N])) < tolF
    assert relerr(co.vcycle("P", v[4 * N:]), mg1._vcycle("P", 0, v[4 * N:])) < 1e-10
    assert relerr(co.solve("F", v[:4 * N]), Mo.F_inv @ v[:4 * N]) < tolF
    assert relerr(co.solve("P", v[4 * N:]), Mo.P_inv @ v[4 * N:]) < 1e-10
    assert relerr(co.precond(v), Mo.matvec(v)) < max(1e-9, 10 * tolF)


def test_c_oracle_jacobi_and_fgmres_vs_reference_run():
    """Damped-Jacobi sub-solves (well-conditioned history): FGMRES vs the golden run of the reference's
    solve_with_approx_schur_pc, 1e-8 relative per iteration."""
    g = golden("solve_jacobi_n16_eta100.npz")
    n, xi, eta_n, eta_s, c, d = g["params"]
    co = COracle(int(n), xi, eta_n, eta_s, c, d, kind="jacobi", F_sweeps=20, P_sweeps=20, omega=0.8)
    assert relerr(co.precond(g["v"]), g["Mv"]) < 1e-10
    x, info, hist = co.fgmres(g["b_vec"], tol=1e-8, restart=150, maxiter=150)
    assert info == 0 and len(hist) == len(g["hist"])
    assert np.allclose(hist, g["hist"], rtol=1e-8)
    assert relerr(x, g["x"]) < 1e-6


def test_c_oracle_custom_theta_and_threads():
    n = 24
    r = (np.arange(n) + 0.5)[:, None] / n
    cc = (np.arange(n) + 0.5)[None, :] / n
    theta = 0.5 + 0.3 * np.sin(2 * np.pi * (cc + 2 * r)) * np.cos(4 * np.pi * cc)
    ops = O.Operators(n, 0.9, 20.0, 1.5, 1.2, -1.0, theta=theta, mass="average")
    co = COracle(n, 0.9, 20.0, 1.5, 1.2, -1.0, theta=theta)
    x = np.random.default_rng(1).standard_normal(5 * n * n)
    assert relerr(co.apply_A(x), ops.A @ x) < 1e-13
    assert co.threads >= 1
