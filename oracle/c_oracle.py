"""ctypes front end of oracle/mpbp_oracle_c.c, the OpenMP C restatement of the reference's hot path.

TEST / BASELINE INFRASTRUCTURE ONLY (see the C file's header).  `build()` compiles it with gcc; the
library lands in oracle/_build/ (git-ignored, travels to the GPU box with the snapshot)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mpbp_oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libmpbp_oracle_c.so")
_lib = None
_dp = C.POINTER(C.c_double)


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stderr)
    return OUT


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.oc_create.restype = C.c_void_p
        lib.oc_create.argtypes = [C.c_int] + [C.c_double] * 7 + [_dp] + [C.c_int] * 5 + [C.c_double] + [C.c_int] * 4 + \
                                 [C.c_double] * 2 + [C.c_int]
        lib.oc_destroy.argtypes = [C.c_void_p]
        lib.oc_num_threads.restype = C.c_int
        lib.oc_set_num_threads.argtypes = [C.c_int]
        lib.oc_set_noise.argtypes = [C.c_double, C.c_ulonglong]
        lib.oc_set_form.argtypes = [C.c_int]
        for name in ("oc_apply_A", "oc_apply_F", "oc_apply_G", "oc_apply_D", "oc_apply_GtG", "oc_precond"):
            getattr(lib, name).argtypes = [C.c_void_p, _dp, _dp]
        lib.oc_vcycle.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        lib.oc_solve.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        lib.oc_fgmres.restype = C.c_int
        lib.oc_fgmres.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int, _dp, C.POINTER(C.c_int)]
        _lib = lib
    return _lib


def set_threads(t=None):
    """Use `t` OpenMP threads (default: every core this process may run on, whatever OMP_NUM_THREADS says)."""
    t = int(t) if t else len(os.sched_getaffinity(0))
    load().oc_set_num_threads(t)
    return load().oc_num_threads()


def set_noise(amp=0.0, seed=0):
    """Conditioning probe: ~amp relative noise on b and on every A.x / M.v inside COracle.fgmres (0 = off)."""
    load().oc_set_noise(float(amp), int(seed))


def set_form(form=0):
    """0: the reference's coefficient table (default); 1: the same rows evaluated differences-first (flux form)."""
    load().oc_set_form(int(form))


def _p(a):
    return a.ctypes.data_as(_dp)


class COracle:
    """Same sub-solver vocabulary as mpbp_oracle.SubSolverConfig; F and P cycles may differ."""

    def __init__(self, n, xi, eta_n, eta_s, c, d_u, d_p=1.0, d_div=-1.0, theta=None, kind="mg", F_cycles=4, P_cycles=2,
                 F_sweeps=20, P_sweeps=20, omega=0.8, nu1=2, nu2=2, n_coarse=4, cheb=True, lmin=0.75, lmax=1.2,
                 project=True):
        self.lib = load()
        self.n, self.N = n, n * n
        th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        self._th = th
        self.h = self.lib.oc_create(n, xi, eta_n, eta_s, c, d_u, d_p, d_div, None if th is None else _p(th),
                                    {"jacobi": 0, "mg": 1}[kind], F_cycles, P_cycles, F_sweeps, P_sweeps, omega, nu1, nu2,
                                    n_coarse, int(cheb), lmin, lmax, int(project))
        if not self.h:
            raise RuntimeError("oc_create failed (singular coarse operator?)")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oc_destroy(self.h)
            self.h = None

    @property
    def threads(self):
        return self.lib.oc_num_threads()

    def _call(self, name, x, out_len):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(out_len)
        getattr(self.lib, name)(self.h, _p(x), _p(y))
        return y

    def apply_A(self, x): return self._call("oc_apply_A", x, 5 * self.N)
    def apply_F(self, x): return self._call("oc_apply_F", x, 4 * self.N)
    def apply_G(self, p): return self._call("oc_apply_G", p, 4 * self.N)
    def apply_D(self, w): return self._call("oc_apply_D", w, self.N)
    def apply_GtG(self, p): return self._call("oc_apply_GtG", p, self.N)
    def precond(self, v): return self._call("oc_precond", v, 5 * self.N)

    def vcycle(self, which, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        self.lib.oc_vcycle(self.h, int(which == "F"), _p(b), _p(x))
        return x

    def solve(self, which, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        self.lib.oc_solve(self.h, int(which == "F"), _p(b), _p(x))
        return x

    def fgmres(self, b, tol=1e-8, restart=40, maxiter=150, use_pc=True):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        hist = np.zeros(maxiter)
        info = C.c_int(0)
        its = self.lib.oc_fgmres(self.h, _p(b), _p(x), tol, restart, maxiter, int(use_pc), _p(hist), C.byref(info))
        return x, info.value, hist[:its]
