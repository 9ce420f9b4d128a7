"""Generates tests/golden/large_*.npz: right-preconditioned FGMRES residual histories of the BENCH configuration
(bench.py: F 6 / GtG 2 Chebyshev-accelerated V(2,2) cycles, restart 40, rtol 1e-8) at BASELINE.json's larger sizes,
computed by the OpenMP C oracle (oracle/mpbp_oracle_c.c -- itself pinned against the numpy oracle and the reference's
golden vectors at small n, tests/test_c_oracle.py), together with the envelope of what another correct fp64
implementation may return (tests/conftest.py:hist_check): 16 re-runs with ~1-ulp noise on b and on every A.x / M.v,
and the run with the OTHER evaluation order of the viscous rows.

Evaluation order matters at these sizes: the reference's coefficient table (sum of coefficient x value terms, what its
dense matmul computes) and the differences-first form (oc_set_form(1), what the CUDA kernels compute) are the same
operator, but at 1024^2, contrast 1e3 the former needs 15 iterations and the latter 13 -- its rounding error, relative
to the values instead of their differences, shows up in the history from iteration 5 on.  The fixture's history is the
differences-first run (the smaller rounding error); the coefficient-table run is part of the envelope and recorded.

    python tests/golden/make_large.py            (about 20 minutes on 8 cores; the GPU box only reads the .npz)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import c_oracle  # noqa: E402
import mpbp_oracle as O  # noqa: E402
from bench import SUB  # noqa: E402  (the benchmarked sub-solver definition)

CASES = [(512, 1.0), (1024, 1.0e3), (1024, 1.0e4)]
RESTART, RTOL, PERTURBED = 40, 1e-8, 16


def sub_kwargs():
    kw = {k: SUB[k] for k in ("kind", "F_cycles", "P_cycles", "cheb", "nu1", "nu2", "omega", "n_coarse")}
    for k in ("lmin", "lmax"):
        if k in SUB:
            kw[k] = SUB[k]
    return kw


def main():
    c_oracle.set_threads()
    for n, eta in CASES:
        t0 = time.time()
        co = c_oracle.COracle(n, 1.0, eta, 1.0, 1.0, -1.0, **sub_kwargs())
        _, b = O.manufactured(n, 1.0, -1.0, 1.0, eta, 1.0)
        c_oracle.set_noise(0.0)
        c_oracle.set_form(1)
        x, info, hist = co.fgmres(b, tol=RTOL, restart=RESTART, maxiter=150)
        env = np.zeros(len(hist))
        its = [len(hist)]

        def fold(h):
            k = min(len(h), len(hist))
            env[:k] = np.maximum(env[:k], np.abs(h[:k] - hist[:k]) / hist[:k])
            if len(h) != len(hist):
                env[k:] = np.inf
            its.append(len(h))
        c_oracle.set_form(0)
        _, info_table, hist_table = co.fgmres(b, tol=RTOL, restart=RESTART, maxiter=150)
        fold(hist_table)
        for s in range(PERTURBED):
            c_oracle.set_form(1 if s % 2 == 0 else 0)
            c_oracle.set_noise(1.2e-16, 1000 + s)
            _, _, h = co.fgmres(b, tol=RTOL, restart=RESTART, maxiter=150)
            fold(h)
        c_oracle.set_noise(0.0)
        c_oracle.set_form(0)
        idx = np.random.default_rng(n).integers(0, len(x), 4096)
        tag = f"large_n{n}_eta{int(eta)}"
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), params=np.array([n, 1.0, eta, 1.0, 1.0, -1.0]),
                            sub=np.array(repr(sub_kwargs())), restart=RESTART, rtol=RTOL, hist=hist, hist_env=env,
                            hist_coefficient_table=hist_table,
                            info=info, its_perturbed=np.array(its), x_norm=np.linalg.norm(x),
                            x_sample=x[idx], x_sample_idx=idx)
        print(f"{tag}: its {len(hist)} differences-first / {len(hist_table)} coefficient table (all runs {min(its)}..{max(its)}), info {info}, env max "
              f"{env[np.isfinite(env)].max():.2e}, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
