import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.fixture(scope="session")
def mp():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mp_block_preconditioners_b200 as pkg
    return pkg


def hist_check(h, ref, env, label="", strict_rel=1e-10, env_factor=10.0, ill=1e-2, verbose=True):
    """Residual-history parity criterion (BASELINE.json: 1e-10 relative, iteration count within +-1).

    `env[k]` is the oracle's OWN reproducibility at entry k: the largest relative change of its history over
    >= 16 runs with ~1-ulp perturbations of b (recorded in the fixture).  Three classes of entries:
      * well conditioned (env <= 1e-11): |h-ref|/ref <= 1e-10, the BASELINE figure, nothing relaxed;
      * conditioned (1e-11 < env <= 1e-2): |h-ref|/ref <= 10 x env;
      * ill conditioned (env > 1e-2 or the perturbed runs changed length): the oracle itself moves by percents
        there (GMRES plateaus where one Ritz value is about to converge), so a relative figure is meaningless;
        instead h[k] must satisfy |h-ref| <= 1e-10 (absolute, histories are relative to ||b||) OR lie between the
        oracle's neighbouring entries, ref[k+1] <= h[k] <= ref[k-1] (GMRES residuals are monotone: "the same
        drop, at most one iteration early or late").
    Returns (max relative deviation over the first two classes, number of ill-conditioned entries)."""
    h, ref, env = np.asarray(h, float), np.asarray(ref, float), np.asarray(env, float)
    assert abs(len(h) - len(ref)) <= 1, f"{label}: iteration count {len(h)} vs oracle {len(ref)}"
    k = min(len(h), len(ref))
    h, r, e = h[:k], ref[:k], env[:k]
    rel = np.abs(h - r) / r
    bad = ~np.isfinite(e) | (e > ill)
    allowed = np.where(e <= 0.1 * strict_rel, strict_rel, np.maximum(strict_rel, env_factor * e))
    ok = rel <= allowed
    lo = np.append(r[1:], 0.0) * (1 - 1e-6)
    hi = np.insert(r[:-1], 0, np.inf) * (1 + 1e-6)
    ok_bad = (np.abs(h - r) <= 1e-10) | ((h >= lo) & (h <= hi))
    good = np.where(bad, ok_bad, ok)
    worst = float(rel[~bad].max()) if (~bad).any() else 0.0
    if verbose:
        print(f"[hist] {label}: {k} entries, max rel dev (conditioned entries) {worst:.2e}, "
              f"ill-conditioned entries {int(bad.sum())} (max rel dev there {float(rel[bad].max()) if bad.any() else 0.0:.2e})")
    assert good.all(), (f"{label}: history deviates at entries {np.nonzero(~good)[0].tolist()}: rel {rel[~good]}, "
                        f"allowed {allowed[~good]}, env {e[~good]}")
    return worst, int(bad.sum())
