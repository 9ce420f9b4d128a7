"""CPU logic checks of the CUDA kernels compiled for the host on a SIMT-on-CPU shim (tests/emu): first the
GPU-proven marching kernels (this validates the shim itself), then the kernels that have not been on a B200
yet (the persistent coarse V-cycle of csrc/coarse.cuh).  Small grids only: every CUDA thread is an OS thread."""
import os
import sys

import numpy as np
import pytest

import mpbp_oracle as O
from conftest import relerr

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402

XI, ETA_N, ETA_S, CC, D = 0.9, 30.0, 1.5, 1.1, -1.0


def _setup(n, analytic=True):
    if analytic:
        theta = O.cell_theta(n)
        ops = O.Operators(n, XI, ETA_N, ETA_S, CC, D)
    else:
        r = (np.arange(n) + 0.5)[:, None] / n
        c = (np.arange(n) + 0.5)[None, :] / n
        theta = 0.5 + 0.3 * np.sin(2 * np.pi * (c + 2 * r)) * np.cos(4 * np.pi * c)
        ops = O.Operators(n, XI, ETA_N, ETA_S, CC, D, theta=theta, mass="average")
    return theta, ops, emu.params(XI, ETA_N, ETA_S, CC, D)


@pytest.mark.parametrize("n,analytic", [(8, True), (12, False)])
def test_shim_reproduces_gpu_proven_kernels(n, analytic):
    theta, ops, prm = _setup(n, analytic)
    mm = 1 if analytic else 0
    rng = np.random.default_rng(n)
    N = n * n
    x = rng.standard_normal(5 * N)
    b = rng.standard_normal(4 * N)
    assert relerr(emu.stokes(0, True, n, prm, mm, theta, x, rs=4), ops.A @ x) < 1e-13
    assert relerr(emu.stokes(0, False, n, prm, mm, theta, x[:4 * N], rs=5), ops.F @ x[:4 * N]) < 1e-13
    assert relerr(emu.stokes(1, False, n, prm, mm, theta, x[:4 * N], b, rs=4), b - ops.F @ x[:4 * N]) < 1e-13
    jac = x[:4 * N] + 0.8 * (b - ops.F @ x[:4 * N]) / ops.F.diagonal()
    assert relerr(emu.stokes(2, False, n, prm, mm, theta, x[:4 * N], b, rs=4), jac) < 1e-13
    assert relerr(emu.jacobi0_F(n, prm, mm, theta, b, rs=4), 0.8 * b / ops.F.diagonal()) < 1e-13
    p = x[4 * N:]
    assert relerr(emu.poisson(0, n, prm, theta, p), ops.GtG @ p) < 1e-13
    assert relerr(emu.poisson(2, n, prm, theta, p, b[:N]), p + 0.8 * (b[:N] - ops.GtG @ p) / ops.GtG.diagonal()) < 1e-13


def test_fused_presmoothing_kernel_logic():
    n = 8
    theta, ops, prm = _setup(n, True)
    rng = np.random.default_rng(1)
    b = rng.standard_normal(4 * n * n)
    dg = ops.F.diagonal()
    x1 = 0.8 * b / dg
    x2 = x1 + 0.8 * (b - ops.F @ x1) / dg
    got = emu.stokes_fused(0, n, prm, 1, theta, b, b, wd=0.8 / dg, rs=4)
    assert relerr(got, x2) < 1e-13


@pytest.mark.parametrize("n,analytic,nu", [(8, True, (2, 2)), (16, False, (2, 2)), (16, True, (1, 3)), (32, True, (2, 2))])
def test_persistent_coarse_vcycle_vs_oracle(n, analytic, nu):
    """csrc/coarse.cuh: the whole V-cycle below a grid size in one single-CTA kernel == the oracle's V-cycle."""
    theta, ops, prm = _setup(n, analytic)
    cfg = O.SubSolverConfig(kind="mg", cycles=1, nu1=nu[0], nu2=nu[1])
    mg = O.Multigrid(ops, cfg)
    last = mg.levels[-1]
    rng = np.random.default_rng(7)
    N = n * n
    bF = rng.standard_normal(4 * N)
    bP = rng.standard_normal(N)
    bP -= bP.mean()
    mm = 1 if analytic else 0
    xF = emu.coarse_vcycle(True, n, prm, mm, theta, 4, 0.8, nu[0], nu[1], last.Finv, bF)
    assert relerr(xF, mg._vcycle("F", 0, bF)) < 1e-11
    xP = emu.coarse_vcycle(False, n, prm, mm, theta, 4, 0.8, nu[0], nu[1], last.Pinv, bP)
    assert relerr(xP, mg._vcycle("P", 0, bP)) < 1e-11


@pytest.mark.parametrize("P,rounds", [(2, 1), (2, 3), (4, 2)])
def test_slab_halo_push_path_on_the_shim(P, rounds):
    """The device side of the multi-GPU path (k_halo_push into the ring neighbours' comm buffers, alternating
    slots keyed by the device-resident sequence number, in-kernel resolve of the halo pointers) with all ranks
    emulated in one process: the assembled slab results equal the global A.x."""
    n = 16
    theta, ops, prm = _setup(n, True)
    x = np.random.default_rng(P).standard_normal(5 * n * n)
    got = emu.slab_apply_A(P, rounds, n, prm, theta, x, rs=3)
    assert relerr(got, ops.A @ x) < 1e-13


def test_div_grad_transfer_and_fused_prolongation_kernels_on_the_shim():
    n = 12
    theta, ops, prm = _setup(n, False)
    rng = np.random.default_rng(2)
    N = n * n
    w, pvec, add = rng.standard_normal(4 * N), rng.standard_normal(N), rng.standard_normal(N)
    assert relerr(emu.div(n, prm, theta, w, add), ops.D @ w + add) < 1e-13
    assert relerr(emu.div(n, prm, theta, w, None, scale=-1.0), -(ops.D @ w)) < 1e-13
    assert relerr(emu.grad(n, prm, theta, pvec), ops.G @ pvec) < 1e-13
    # transfers vs the oracle's restrict/prolong helpers
    f4 = w.reshape(4, n, n)
    ref_r = np.concatenate([O.restrict_u(f4[0]).ravel(), O.restrict_v(f4[1]).ravel(), O.restrict_u(f4[2]).ravel(),
                            O.restrict_v(f4[3]).ravel()])
    assert relerr(emu.transfer(0, n, w, np.zeros(N)), ref_r) < 1e-14
    ec = rng.standard_normal(N)  # coarse 4 fields of (n/2)^2
    e4 = ec.reshape(4, n // 2, n // 2)
    ref_p = w + np.concatenate([O.prolong_u(e4[0]).ravel(), O.prolong_v(e4[1]).ravel(), O.prolong_u(e4[2]).ravel(),
                                O.prolong_v(e4[3]).ravel()])
    assert relerr(emu.transfer(1, n, ec, w.copy()), ref_p) < 1e-14
    assert relerr(emu.transfer(2, n, pvec, np.zeros(N // 4)), O.restrict_cell(pvec.reshape(n, n)).ravel()) < 1e-14
    assert relerr(emu.transfer(3, n, ec[:N // 4], pvec.copy()),
                  pvec + O.prolong_cell(ec[:N // 4].reshape(n // 2, n // 2)).ravel()) < 1e-14
    # fused prolongation + Jacobi sweep (k_stokes_fused<1>) == prolong_add followed by one damped sweep
    b = rng.standard_normal(4 * N)
    xt = ref_p
    ref = xt + 0.8 * (b - ops.F @ xt) / ops.F.diagonal()
    assert relerr(emu.stokes_fused(1, n, prm, 0, theta, w, b, ec=ec, rs=4), ref) < 1e-13


@pytest.mark.parametrize("P,rs", [(2, 3), (2, 8), (4, 2), (4, 4)])
def test_fused_halo_push_chain_on_the_shim(P, rs):
    """Experimental fused push (k_stokes_push): two Jacobi sweeps that push their own boundary rows to the ring
    neighbours, followed by a residual that consumes them -- no k_halo_push launch after the first exchange.
    rs=8 makes one strip per 8-row slab (a block that is both the first and the last strip)."""
    n = 16
    theta, ops, prm = _setup(n, True)
    rng = np.random.default_rng(P + rs)
    x0, b = rng.standard_normal(4 * n * n), rng.standard_normal(4 * n * n)
    dg = ops.F.diagonal()
    x1 = x0 + 0.8 * (b - ops.F @ x0) / dg
    x2 = x1 + 0.8 * (b - ops.F @ x1) / dg
    got = emu.slab_fused_push_chain(P, n, prm, theta, x0, b, rs=rs)
    assert relerr(got, b - ops.F @ x2) < 1e-12


@pytest.mark.parametrize("n,analytic,rs,re", [(8, True, 4, 0), (12, False, 6, 0), (32, True, 4, 0), (36, False, 8, 0),
                                              (32, True, 6, 4), (36, False, 10, 8)])
def test_unified_marching_kernel_and_its_fused_variants(n, analytic, rs, re):
    """csrc/stokes.cuh: every (IN, MODE, EP) instantiation the plan launches, against compositions of the oracle's
    operators.  n=32/36 with short strips exercise the lean interior instantiation (strips with r0 >= 2 and
    r1 + 4 <= rows) next to the edge one; n=36 is not a multiple of the warp tile; re > 0 is the single-wave
    decomposition (two short edge strips of `re` rows, interior strips of `rs` rows, the last one ragged)."""
    theta, ops, prm = _setup(n, analytic)
    mm = 1 if analytic else 0
    rng = np.random.default_rng(n + rs)
    N = n * n
    x5 = rng.standard_normal(5 * N)
    x, b = x5[:4 * N], rng.standard_normal(4 * N)
    dg = ops.F.diagonal()
    sweep = lambda v: v + 0.8 * (b - ops.F @ v) / dg
    kw = dict(rs=rs, re=re)
    assert relerr(emu.stokes_x(0, 0, 0, n, prm, mm, theta, x5, with_p=True, **kw), ops.A @ x5) < 1e-13
    assert relerr(emu.stokes_x(0, 0, 0, n, prm, mm, theta, x, **kw), ops.F @ x) < 1e-13
    assert relerr(emu.stokes_x(0, 1, 0, n, prm, mm, theta, x, b, **kw), b - ops.F @ x) < 1e-13
    assert relerr(emu.stokes_x(0, 2, 0, n, prm, mm, theta, x, b, **kw), sweep(x)) < 1e-13
    # IN 1: both pre-smoothing sweeps from a zero guess in one pass
    x1 = 0.8 * b / dg
    assert relerr(emu.stokes_x(1, 2, 0, n, prm, mm, theta, b, b, wd=0.8 / dg, **kw), sweep(x1)) < 1e-13
    # IN 2: coarse-grid correction + first post-smoothing sweep
    ec = rng.standard_normal(N)
    e4 = ec.reshape(4, n // 2, n // 2)
    xt = x + np.concatenate([O.prolong_u(e4[0]).ravel(), O.prolong_v(e4[1]).ravel(), O.prolong_u(e4[2]).ravel(),
                             O.prolong_v(e4[3]).ravel()])
    assert relerr(emu.stokes_x(2, 2, 0, n, prm, mm, theta, x, b, ec=ec, **kw), sweep(xt)) < 1e-13
    # EP 1: last sweep + Chebyshev update (z never stored): later cycle, first cycle, last cycle
    d0, xk0 = rng.standard_normal(4 * N), rng.standard_normal(4 * N)
    z = sweep(x)
    d1, xk1 = emu.stokes_x(0, 2, 1, n, prm, mm, theta, x, b, d=d0, xk=xk0, cheb=(0.3, 1.7), flags=(1, 1, 1), **kw)
    assert relerr(d1, 0.3 * d0 + 1.7 * z) < 1e-13 and relerr(xk1, xk0 + 0.3 * d0 + 1.7 * z) < 1e-13
    d1, xk1 = emu.stokes_x(0, 2, 1, n, prm, mm, theta, x, b, d=d0, xk=xk0, cheb=(0.0, 1.7), flags=(0, 0, 1), **kw)
    assert relerr(d1, 1.7 * z) < 1e-13 and relerr(xk1, 1.7 * z) < 1e-13
    d1, xk1 = emu.stokes_x(0, 2, 1, n, prm, mm, theta, x, b, d=d0, xk=xk0, cheb=(0.3, 1.7), flags=(1, 1, 0), **kw)
    assert np.array_equal(d1, d0) and relerr(xk1, xk0 + 0.3 * d0 + 1.7 * z) < 1e-13
    # IN 2 + EP 1 (V(2,1)-style cycles: the only post-sweep is also the last one)
    d1, xk1 = emu.stokes_x(2, 2, 1, n, prm, mm, theta, x, b, ec=ec, d=d0, xk=xk0, cheb=(0.3, 1.7), **kw)
    assert relerr(xk1, xk0 + 0.3 * d0 + 1.7 * sweep(xt)) < 1e-13
    # EP 2: residual + full-weighting restriction (only the coarse rhs is stored)
    r4 = (b - ops.F @ x).reshape(4, n, n)
    ref_r = np.concatenate([O.restrict_u(r4[0]).ravel(), O.restrict_v(r4[1]).ravel(), O.restrict_u(r4[2]).ravel(),
                            O.restrict_v(r4[3]).ravel()])
    assert relerr(emu.stokes_x(0, 1, 2, n, prm, mm, theta, x, b, **kw), ref_r) < 1e-13


@pytest.mark.parametrize("P,n,rs", [(2, 16, 4), (4, 16, 4), (2, 32, 4), (4, 32, 8), (3, 24, 4)])
def test_fused_vcycle_kernels_on_distributed_levels(P, n, rs):
    """Distributed-level data paths of the fused V-cycle with all ranks emulated: static halo rows of omega/diag(F),
    halo rows stashed by the residual kernel and re-read by the prolongation + sweep kernel while the live comm slots
    carry the coarse correction, fused pushes, producer credits, strip reordering."""
    theta, ops, prm = _setup(n, True)
    rng = np.random.default_rng(P * n + rs)
    N = n * n
    b, ec = rng.standard_normal(4 * N), rng.standard_normal(N)
    dg = ops.F.diagonal()
    wd = 0.8 / dg
    sweep = lambda v: v + 0.8 * (b - ops.F @ v) / dg
    x2 = sweep(wd * b)
    e4 = ec.reshape(4, n // 2, n // 2)
    xt = x2 + np.concatenate([O.prolong_u(e4[0]).ravel(), O.prolong_v(e4[1]).ravel(), O.prolong_u(e4[2]).ravel(),
                              O.prolong_v(e4[3]).ravel()])
    got_x, got_r = emu.slab_fused_vcycle_chain(P, n, prm, theta, b, wd, ec, rs=rs)
    assert relerr(got_r, b - ops.F @ x2) < 1e-12
    assert relerr(got_x, sweep(sweep(xt))) < 1e-12


@pytest.mark.timeout(180)
@pytest.mark.parametrize("P,n,rs", [(2, 16, 4), (4, 16, 4), (2, 32, 8), (4, 32, 4), (3, 24, 4), (1, 16, 4)])
def test_fused_residual_restriction_on_slabs(P, n, rs):
    """EP 2 with PUSH (distributed levels): the residual is restricted in registers; a slab's first coarse row is
    completed with the previous rank's last-row half-sums, exchanged inside the kernel; the coarse rhs's boundary rows
    reach the ring neighbours' halo areas."""
    theta, ops, prm = _setup(n, True)
    rng = np.random.default_rng(7 * P + n + rs)
    N = n * n
    x, b = rng.standard_normal(4 * N), rng.standard_normal(4 * N)
    r4 = (b - ops.F @ x).reshape(4, n, n)
    ref = np.stack([O.restrict_u(r4[0]), O.restrict_v(r4[1]), O.restrict_u(r4[2]), O.restrict_v(r4[3])])  # [4][nc][nc]
    got, rows = emu.slab_residual_restrict(P, n, prm, theta, x, b, rs=rs)
    assert relerr(got, ref.ravel()) < 1e-13
    nc, rc = n // 2, n // 2 // P
    for g in range(P):
        top = ref[:, (g * rc - 1) % nc, :]        # the previous rank's last coarse row
        bot = ref[:, ((g + 1) * rc) % nc, :]      # the next rank's first coarse row
        assert relerr(rows[g, 0], top) < 1e-13 and relerr(rows[g, 1], bot) < 1e-13


@pytest.mark.parametrize("n", [8, 12, 36, 64])
def test_cell_parallel_kernels_of_the_small_levels(n):
    """csrc/cell.cuh: one thread per cell, the same fused variants as the marching kernels (levels below 0: face-average
    mass term, user-style theta), against compositions of the oracle's operators."""
    theta, ops, prm = _setup(n, False)
    rng = np.random.default_rng(n)
    N = n * n
    x, b = rng.standard_normal(4 * N), rng.standard_normal(4 * N)
    dg = ops.F.diagonal()
    sweep = lambda v: v + 0.8 * (b - ops.F @ v) / dg
    assert relerr(emu.cell(0, n, prm, theta, x=x, b=b), sweep(x)) < 1e-13
    assert relerr(emu.cell(1, n, prm, theta, b=b, wd=0.8 / dg), sweep(0.8 * b / dg)) < 1e-13
    ec = rng.standard_normal(N)
    e4 = ec.reshape(4, n // 2, n // 2)
    xt = x + np.concatenate([O.prolong_u(e4[0]).ravel(), O.prolong_v(e4[1]).ravel(), O.prolong_u(e4[2]).ravel(),
                             O.prolong_v(e4[3]).ravel()])
    assert relerr(emu.cell(2, n, prm, theta, x=x, b=b, ec=ec), sweep(xt)) < 1e-13
    r4 = (b - ops.F @ x).reshape(4, n, n)
    ref_r = np.concatenate([O.restrict_u(r4[0]).ravel(), O.restrict_v(r4[1]).ravel(), O.restrict_u(r4[2]).ravel(),
                            O.restrict_v(r4[3]).ravel()])
    assert relerr(emu.cell(3, n, prm, theta, x=x, b=b), ref_r) < 1e-13


@pytest.mark.parametrize("n,analytic,rs", [(8, True, 4), (12, False, 2), (32, True, 4), (36, False, 6), (64, True, 8)])
def test_fused_pressure_poisson_kernels(n, analytic, rs):
    """csrc/poisson.cuh: pre-smoothing pair, residual + 4-cell-average restriction, piecewise-constant prolongation +
    sweep (with and without the Chebyshev epilogue), against compositions of the oracle's GtG operator and transfers."""
    theta, ops, prm = _setup(n, analytic)
    rng = np.random.default_rng(n + rs)
    N = n * n
    p, b = rng.standard_normal(N), rng.standard_normal(N)
    dg = ops.GtG.diagonal()
    sweep = lambda v: v + 0.8 * (b - ops.GtG @ v) / dg
    assert relerr(emu.poisson_f(1, n, prm, theta, b=b, wd=1.0 / dg, rs=rs), sweep(0.8 * b / dg)) < 1e-13
    ref_r = O.restrict_cell((b - ops.GtG @ p).reshape(n, n)).ravel()
    assert relerr(emu.poisson_f(3, n, prm, theta, x=p, b=b, rs=rs), ref_r) < 1e-13
    ec = rng.standard_normal(N // 4)
    pt = p + O.prolong_cell(ec.reshape(n // 2, n // 2)).ravel()
    assert relerr(emu.poisson_f(2, n, prm, theta, x=p, b=b, ec=ec, rs=rs), sweep(pt)) < 1e-13
    d0, xk0 = rng.standard_normal(N), rng.standard_normal(N)
    d1, xk1 = emu.poisson_f(4, n, prm, theta, x=p, b=b, ec=ec, d=d0, xk=xk0, cheb=(0.3, 1.7), rs=rs)
    assert relerr(d1, 0.3 * d0 + 1.7 * sweep(pt)) < 1e-13 and relerr(xk1, xk0 + 0.3 * d0 + 1.7 * sweep(pt)) < 1e-13
    d1, xk1 = emu.poisson_f(4, n, prm, theta, x=p, b=b, ec=ec, d=d0, xk=xk0, cheb=(0.0, 1.7), flags=(0, 0, 1), rs=rs)
    assert relerr(d1, 1.7 * sweep(pt)) < 1e-13 and relerr(xk1, 1.7 * sweep(pt)) < 1e-13


@pytest.mark.timeout(180)
@pytest.mark.parametrize("P,seed", [(3, 1), (4, 2), (4, 3), (2, 4)])
def test_ll_halo_protocol_under_rank_skew(P, seed):
    """Slot reuse of the LL halo protocol with free-running ranks (>= 3 ranks, skew between push and consumer):
    every emulated rank runs seven exchanging kernels in its own thread with random delays; only the sequence tags, the
    three slots and the producer credit order them.  A protocol error shows up as wrong rows or as a hang (timeout)."""
    n = 24 if P == 3 else 16
    theta, ops, prm = _setup(n, True)
    rng = np.random.default_rng(seed)
    x, b = rng.standard_normal(4 * n * n), rng.standard_normal(4 * n * n)
    dg = ops.F.diagonal()
    for _ in range(6):
        x_next = x + 0.8 * (b - ops.F @ x) / dg
        if _ == 0:
            x0 = x
        x = x_next
    got = emu.slab_push_chain_skewed(P, n, prm, theta, x0, b, sweeps=6, rs=4, seed=seed)
    assert relerr(got, b - ops.F @ x) < 1e-11


@pytest.mark.timeout(180)
@pytest.mark.parametrize("P,n,rs,seed", [(2, 16, 4, 1), (4, 16, 4, 2), (3, 24, 2, 3), (2, 32, 16, 4), (4, 32, 4, 5)])
def test_pressure_kernels_on_slabs_with_fused_pushes(P, n, rs, seed):
    """k_poisson on distributed levels: edge strips fetch the ring neighbours' rows (a separate instantiation of the
    march) and push their own; free-running ranks as in the skew test.  rs=16 at n=32, P=2 makes one strip per slab (a
    block that is both the first and the last strip)."""
    theta, ops, prm = _setup(n, True)
    rng = np.random.default_rng(seed)
    N = n * n
    p, b = rng.standard_normal(N), rng.standard_normal(N)
    dg = ops.GtG.diagonal()
    x = p
    for _ in range(4):
        x = x + 0.8 * (b - ops.GtG @ x) / dg
    got = emu.slab_poisson_chain(P, n, prm, theta, p, b, sweeps=4, rs=rs, seed=seed)
    assert relerr(got, b - ops.GtG @ x) < 1e-12
