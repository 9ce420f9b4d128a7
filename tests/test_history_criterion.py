"""The residual-history parity criterion (tests/conftest.py:hist_check) checked on the CPU: it must ACCEPT a second,
independent correct implementation (the C oracle with an odd thread count, i.e. its own rounding in every
operation) against the golden histories produced through the reference, and REJECT histories that are wrong
(scaled by 1.5, shifted by one or two iterations) -- i.e. the envelope relaxation is not vacuous."""
import numpy as np
import pytest

import c_oracle
from conftest import golden, hist_check

SOLVES = [("solve_mgcheb_n16_eta100.npz", dict(kind="mg", F_cycles=4, P_cycles=4, cheb=True)),
          ("solve_mgplain_n16_eta100.npz", dict(kind="mg", F_cycles=2, P_cycles=2, cheb=False)),
          ("solve_jacobi_n16_eta100.npz", dict(kind="jacobi", F_sweeps=20, P_sweeps=20, omega=0.8)),
          ("solve_mgcheb_n32_eta1.npz", dict(kind="mg", F_cycles=4, P_cycles=4, cheb=True))]


@pytest.mark.parametrize("fx,kw", SOLVES)
def test_criterion_accepts_independent_implementation_and_rejects_wrong_histories(fx, kw):
    g = golden(fx)
    n, xi, eta_n, eta_s, c, d = g["params"]
    c_oracle.set_threads(5)
    co = c_oracle.COracle(int(n), xi, eta_n, eta_s, c, d, **kw)
    _, info, h = co.fgmres(g["b_vec"], tol=1e-8, restart=150, maxiter=150)
    c_oracle.set_threads()
    assert info == 0
    worst, kstar = hist_check(h, g["hist"], g["hist_sens"], label=f"C oracle vs {fx}")
    assert kstar >= 6  # the reproducible prefix (held to 10 x the oracle's own scatter) is never trivial
    for wrong in (h * 1.5, np.concatenate([h[:1], h[:-1]]), np.concatenate([h[:2], h[:-2]]), h[:-2]):
        with pytest.raises(AssertionError):
            hist_check(wrong, g["hist"], g["hist_sens"], verbose=False)


def test_envelopes_have_sixteen_runs_and_strict_entries():
    """Where the oracle's history is reproducible (Jacobi sub-solves, eta=1) the criterion is the plain 1e-10."""
    for fx in ("solve_jacobi_n16_eta100.npz", "solve_mgcheb_n32_eta1.npz"):
        env = golden(fx)["hist_sens"]
        assert np.isfinite(env).all() and env.max() < 1e-8
