"""ctypes front end of the SIMT-on-CPU build of the CUDA kernels (tests/emu/emu_kernels.cpp).
TEST INFRASTRUCTURE ONLY: logic checks of kernels on the GPU-less build container."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libmpbp_emu.so")
SRCS = [os.path.join(HERE, "emu_kernels.cpp"), os.path.join(HERE, "cuda_emu.h"),
        os.path.join(ROOT, "mp-block-preconditioners_b200", "csrc", "stencil.cuh"),
        os.path.join(ROOT, "mp-block-preconditioners_b200", "csrc", "ll.cuh"),
        os.path.join(ROOT, "mp-block-preconditioners_b200", "csrc", "cell.cuh"),
        os.path.join(ROOT, "mp-block-preconditioners_b200", "csrc", "poisson.cuh"),
        os.path.join(ROOT, "mp-block-preconditioners_b200", "csrc", "coarse.cuh"),
        os.path.join(ROOT, "mp-block-preconditioners_b200", "csrc", "stokes.cuh")]
_dp = C.POINTER(C.c_double)
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in SRCS):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            # -Bsymbolic: the kernel templates have the same mangled names as the CUDA launch stubs inside libmpbp.so
            # (which other tests load RTLD_GLOBAL); this library must bind to its own host-compiled kernels
            cmd = ["g++", "-O1", "-std=c++17", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-Wl,-Bsymbolic",
                   "-I" + HERE, "-o", OUT, SRCS[0]]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("g++ failed:\n" + res.stderr[-4000:])
        _lib = C.CDLL(OUT)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def pad_theta(theta):
    """(n+2) x n padded theta (row -1 and row n are the periodic neighbours), the layout plan.cu uploads."""
    return np.ascontiguousarray(np.concatenate([theta[-1:], theta, theta[:1]], axis=0))


def params(xi, eta_n, eta_s, c, d_u, d_p=1.0, d_div=-1.0):
    return np.array([xi, eta_n, eta_s, c, d_u, d_p, d_div], dtype=np.float64)


def stokes(mode, with_p, n, prm, mass_mode, theta, x, b=None, rs=8, pf=3, omega=0.8):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(5 * n * n if with_p else 4 * n * n)
    thp = pad_theta(theta)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    load().emu_stokes(mode, int(with_p), n, _p(prm), mass_mode, _p(thp), _p(x), _p(bb), _p(y), rs, pf, C.c_double(omega))
    return y


def stokes_fused(variant, n, prm, mass_mode, theta, x, b, wd=None, ec=None, rs=8, pf=3, omega=0.8):
    y = np.zeros(4 * n * n)
    thp = pad_theta(theta)
    x = np.ascontiguousarray(x, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    wd = None if wd is None else np.ascontiguousarray(wd, dtype=np.float64)
    ec = None if ec is None else np.ascontiguousarray(ec, dtype=np.float64)
    load().emu_stokes_fused(variant, n, _p(prm), mass_mode, _p(thp), _p(x), _p(b), _p(wd), _p(ec), _p(y), rs, pf,
                            C.c_double(omega))
    return y


def stokes_x(IN, MODE, EP, n, prm, mass_mode, theta, x, b=None, with_p=False, wd=None, ec=None, d=None, xk=None,
             cheb=None, flags=(1, 1, 1), rs=8, pf=3, omega=0.8, re=0):
    """csrc/stokes.cuh: k_stokes_x<IN, MODE, WITH_P, EP>.  Returns y (EP 0), (d, xk) (EP 1) or the coarse rhs (EP 2)."""
    c = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
    x, b, wd, ec = c(x), c(b), c(wd), c(ec)
    thp = pad_theta(theta)
    out = np.zeros(n * n if EP == 2 else (5 if with_p else 4) * n * n)
    if EP == 1:
        d, xk = np.array(d, dtype=np.float64), np.array(xk, dtype=np.float64)
    ch = None if cheb is None else np.array(cheb, dtype=np.float64)
    fl = (C.c_int * 3)(*flags)
    load().emu_stokes_x(IN, MODE, int(with_p), EP, n, _p(prm), mass_mode, _p(thp), _p(x), _p(b), _p(wd), _p(ec),
                        _p(d), _p(xk), _p(ch), fl, _p(out), rs, pf, C.c_double(omega), re)
    return (d, xk) if EP == 1 else out


def jacobi0_F(n, prm, mass_mode, theta, b, rs=8, omega=0.8):
    y = np.zeros(4 * n * n)
    thp = pad_theta(theta)
    b = np.ascontiguousarray(b, dtype=np.float64)
    load().emu_jacobi0_F(n, _p(prm), mass_mode, _p(thp), _p(b), _p(y), rs, C.c_double(omega))
    return y


def poisson(mode, n, prm, theta, p, b=None, rs=8, omega=0.8):
    y = np.zeros(n * n)
    thp = pad_theta(theta)
    p = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    load().emu_poisson(mode, n, _p(prm), _p(thp), _p(p), _p(b), _p(y), rs, C.c_double(omega))
    return y


def coarse_vcycle(isF, n, prm, mass_mode, theta, n_coarse, omega, nu1, nu2, Minv, b, threads=256):
    """Minv: row-major dense (pseudo-)inverse of the coarsest level."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    Mt = np.ascontiguousarray(Minv, dtype=np.float64)  # row-major, as plan.cu uploads it
    th = np.ascontiguousarray(theta, dtype=np.float64)
    load().emu_coarse_vcycle(int(isF), n, _p(prm), mass_mode, _p(th), n_coarse, C.c_double(omega), nu1, nu2, _p(Mt),
                             Minv.shape[0], _p(b), _p(x), threads)
    return x


def slab_apply_A(P, rounds, n, prm, theta, x, rs=4):
    """A.x with the grid split into P row slabs whose halo rows travel through the peer-memory push path
    (k_halo_push -> comm buffer slots/flags -> in-kernel resolve), all "ranks" emulated in this process."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    load().emu_slab_apply_A(P, rounds, n, _p(prm), _p(th), _p(x), _p(y), rs)
    return y


def div(n, prm, theta, w, add=None, scale=1.0, rs=4):
    out = np.zeros(n * n)
    thp = pad_theta(theta)
    w = np.ascontiguousarray(w, dtype=np.float64)
    add = None if add is None else np.ascontiguousarray(add, dtype=np.float64)
    load().emu_div_grad(0, n, _p(prm), _p(thp), _p(w), _p(add), _p(out), rs, C.c_double(scale))
    return out


def grad(n, prm, theta, p, rs=4):
    out = np.zeros(4 * n * n)
    thp = pad_theta(theta)
    p = np.ascontiguousarray(p, dtype=np.float64)
    load().emu_div_grad(1, n, _p(prm), _p(thp), _p(p), None, _p(out), rs, C.c_double(1.0))
    return out


def transfer(what, nf, x, out):
    x = np.ascontiguousarray(x, dtype=np.float64)
    load().emu_transfer(what, nf, _p(x), _p(out))
    return out


def slab_fused_push_chain(P, n, prm, theta, x0, b, rs=4, omega=0.8):
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros_like(b)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    load().emu_slab_fused_push_chain(P, n, _p(prm), _p(th), _p(x0), _p(b), _p(out), rs, C.c_double(omega))
    return out


def slab_fused_vcycle_chain(P, n, prm, theta, b, wd, ec, rs=4, omega=0.8):
    c = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    b, wd, ec, th = c(b), c(wd), c(ec), c(theta)
    out_x, out_r = np.zeros_like(b), np.zeros_like(b)
    load().emu_slab_fused_vcycle_chain(P, n, _p(prm), _p(th), _p(b), _p(wd), _p(ec), _p(out_x), _p(out_r), rs,
                                       C.c_double(omega))
    return out_x, out_r


def slab_residual_restrict(P, n, prm, theta, x, b, rs=4):
    """EP 2 + PUSH: fused residual + restriction on P emulated slabs (ranks run concurrently).  Returns the assembled
    coarse rhs and the coarse rows each rank received from its ring neighbours, [P][2 (top, bottom)][4][n/2]."""
    c = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    x, b, th = c(x), c(b), c(theta)
    nc = n // 2
    out = np.zeros(4 * nc * nc)
    rows = np.zeros((P, 2, 4, nc))
    load().emu_slab_residual_restrict(P, n, _p(prm), _p(th), _p(x), _p(b), _p(out), _p(rows), rs)
    return out, rows


def cell(IN, n, prm, theta, x=None, b=None, wd=None, ec=None, omega=0.8):
    """csrc/cell.cuh: k_cell_sweep<IN> (IN 0/1/2) or k_cell_rr (IN 3) on a whole-grid level with face-average mass."""
    c = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
    x, b, wd, ec = c(x), c(b), c(wd), c(ec)
    thp = pad_theta(theta)
    out = np.zeros(n * n if IN == 3 else 4 * n * n)
    load().emu_cell(IN, n, _p(prm), _p(thp), _p(x), _p(b), _p(wd), _p(ec), _p(out), C.c_double(omega))
    return out


def poisson_f(IN, n, prm, theta, x=None, b=None, wd=None, ec=None, d=None, xk=None, cheb=None, flags=(1, 1, 1), rs=4,
              omega=0.8):
    """csrc/poisson.cuh: k_poisson_f -- IN 1 pair, 3 residual + restriction, 2 prolongation + sweep, 4 the same with the
    Chebyshev epilogue (returns (d, xk))."""
    c = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
    x, b, wd, ec = c(x), c(b), c(wd), c(ec)
    thp = pad_theta(theta)
    out = np.zeros((n // 2) ** 2 if IN == 3 else n * n)
    if IN == 4:
        d, xk = np.array(d, dtype=np.float64), np.array(xk, dtype=np.float64)
    ch = None if cheb is None else np.array(cheb, dtype=np.float64)
    fl = (C.c_int * 3)(*flags)
    load().emu_poisson_f(IN, n, _p(prm), _p(thp), _p(x), _p(b), _p(wd), _p(ec), _p(d), _p(xk), _p(ch), fl, _p(out), rs,
                         C.c_double(omega))
    return (d, xk) if IN == 4 else out


def slab_push_chain_skewed(P, n, prm, theta, x0, b, sweeps=5, rs=4, omega=0.8, seed=1):
    """LL halo protocol under skew: P free-running emulated ranks (one host thread each, random delays between their
    kernels) run `sweeps` fused-push Jacobi sweeps and a residual.  Returns the assembled residual."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros_like(b)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    load().emu_slab_push_chain_skewed(P, n, _p(prm), _p(th), _p(x0), _p(b), _p(out), rs, C.c_double(omega), sweeps, seed)
    return out


def slab_poisson_chain(P, n, prm, theta, p0, b, sweeps=4, rs=4, omega=0.8, seed=1):
    """k_poisson on P free-running emulated slabs: push, `sweeps` fused-push GtG sweeps, residual."""
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros_like(b)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    load().emu_slab_poisson_chain(P, n, _p(prm), _p(th), _p(p0), _p(b), _p(out), rs, C.c_double(omega), sweeps, seed)
    return out
