"""precond_apply / V-cycle time at 4096^2 under the fusion and coarse-kernel knobs (run under gpurun).
usage: python profiles/fuse_sweep.py [n]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import mp_block_preconditioners_b200 as mp
from mp_block_preconditioners_b200._cabi import check
from mp_block_preconditioners_b200.utils import manufactured_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = bench.WORKLOAD
peak, _ = bench.measured_peak()
out = []


def timeit(fn, reps=5):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


KNOBS = ("MPBP_FUSE", "MPBP_GRAPH", "MPBP_COARSE", "MPBP_PF", "MPBP_WAVE", "MPBP_RS")
for env in [dict(), dict(MPBP_WAVE="1"), dict(MPBP_FUSE="1"), dict(MPBP_FUSE="3"), dict(MPBP_FUSE="5"),
            dict(MPBP_GRAPH="0"), dict(MPBP_COARSE="64"), dict(MPBP_PF="2"), dict(MPBP_PF="4"), dict(MPBP_PF="5"),
            dict(MPBP_RS="16"), dict(MPBP_RS="24"), dict(MPBP_RS="48")]:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    bp = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=mp.SubSolver(**bench.SUB))
    A = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])[0]
    p, lib = A.plan, A.plan.lib
    N = p.N
    u, b = manufactured_device(p)
    x4 = torch.zeros(4 * N, dtype=torch.float64, device="cuda")
    y5 = torch.empty(5 * N, dtype=torch.float64, device="cuda")
    st = p.stream()
    ms_j = timeit(lambda: check(lib.mpbp_jacobi_F(p.h, b.data_ptr(), x4.data_ptr(), 10, 0.8, st))) / 10
    ms_a = timeit(lambda: check(lib.mpbp_apply_A(p.h, b.data_ptr(), y5.data_ptr(), st)), reps=10)
    ms_v = timeit(lambda: check(lib.mpbp_vcycle_F(p.h, b.data_ptr(), x4.data_ptr(), st)))
    ms_s = timeit(lambda: check(lib.mpbp_solve_F(p.h, b.data_ptr(), x4.data_ptr(), st)))
    ms_p = timeit(lambda: check(lib.mpbp_precond_apply(p.h, b.data_ptr(), y5.data_ptr(), st)))
    by = p.precond_bytes()
    rec = dict(env=env, jacobi_F_ms=ms_j, jacobi_frac=104 * N / ms_j / 1e6 / peak, apply_A_ms=ms_a,
               apply_A_frac=88 * N / ms_a / 1e6 / peak, vcycle_F_ms=ms_v, solve_F_ms=ms_s, precond_ms=ms_p, precond_GB=by / 1e9, gbs=by / ms_p / 1e6,
               frac=by / ms_p / 1e6 / peak)
    out.append(rec)
    print(rec, flush=True)
    p.close()
    del bp, A, p, u, b, x4, y5
    torch.cuda.empty_cache()
json.dump(out, open(f"gpurun_out/fuse_sweep_n{n}.json", "w"), indent=1)
