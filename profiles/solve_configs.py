"""Full preconditioned solves at BASELINE.json's configurations (run under gpurun): iterations to
rtol 1e-8, wall time, true residual and discretisation error vs the manufactured solution."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mp_block_preconditioners_b200 as mp
from mp_block_preconditioners_b200.utils import device_norms, manufactured_device

# (n, eta_n, F cycles, GtG cycles, Chebyshev interval)
CONFIGS = [(512, 1.0, 6, 2, (0.75, 1.2)), (2048, 1.0e3, 6, 2, (0.75, 1.2)), (4096, 1.0e4, 6, 2, (0.75, 1.2)),
           (4096, 1.0e4, 6, 2, (0.78, 1.17)), (2048, 1.0e4, 6, 2, (0.75, 1.2)), (1024, 1.0e4, 6, 2, (0.75, 1.2))]
import bench  # noqa: E402  (the benchmarked hierarchy: n_coarse)
if len(sys.argv) > 1:
    CONFIGS = [c for c in CONFIGS if c[0] <= int(sys.argv[1])]
out = []
for n, eta_n, kF, kP, (lmin, lmax) in CONFIGS:
    sub = mp.SubSolver(kind="mg", F_cycles=kF, P_cycles=kP, cheb=True, lmin=lmin, lmax=lmax, n_coarse=bench.SUB["n_coarse"])
    bp = mp.MultiphaseBlockPreconditioner(n, 1.0, eta_n, 1.0, sub_solver=sub)
    A = bp.get_big_A_matrix(1.0, -1.0)[0]
    M = bp.approx_schur_operator(1.0, -1.0)
    u, b = manufactured_device(A.plan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = mp.fgmres(A, b, M=M, tol=1e-8, restart=40, maxiter=10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    hist = mp.fgmres.last_history
    true_rel = float(torch.linalg.norm(b - A @ x) / torch.linalg.norm(b))
    L1, L2, mx = device_norms(A.plan, x, u, (1.0 / n) ** 2)
    rec = dict(n=n, eta_n=eta_n, F_cycles=kF, P_cycles=kP, interval=[lmin, lmax], restart=40, info=info, iterations=len(hist), seconds=dt,
               its_per_s=len(hist) / dt, true_rel_residual=true_rel, err_L1=L1, err_L2=L2, err_max=mx,
               history=[float(h) for h in hist])
    out.append(rec)
    print({k: v for k, v in rec.items() if k != "history"}, flush=True)
    A.plan.close()
    del A, M, u, b, x, bp
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/solve_configs.json", "w"), indent=1)
