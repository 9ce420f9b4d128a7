"""Per-kernel achieved GB/s at level 0 (CUDA events, 20 launches each after warm-up), run under gpurun.
usage: python profiles/kernel_table.py [n]"""
import ctypes as C
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
bp = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=mp.SubSolver(**bench.SUB))
A = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])[0]
p, lib = A.plan, A.plan.lib
N = p.N
peak, _ = bench.measured_peak()
u, b = manufactured_device(p)
x4 = torch.zeros(4 * N, dtype=torch.float64, device="cuda")
y5 = torch.empty(5 * N, dtype=torch.float64, device="cuda")
x1 = torch.zeros(N, dtype=torch.float64, device="cuda")
st = p.stream()
bp_ptr = b.data_ptr() + 4 * N * 8


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []


def add(name, bytes_per_cell, fn, div=1):
    ms = timeit(fn) / div
    gbs = bytes_per_cell * N / ms / 1e6
    rows.append((name, bytes_per_cell, ms, gbs, gbs / peak))
    print(f"{name:34s} {bytes_per_cell:5.0f} N  {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {gbs / peak:5.2f}", flush=True)


add("apply_A  k_stokes<0,true>", 88, lambda: check(lib.mpbp_apply_A(p.h, b.data_ptr(), y5.data_ptr(), st)))
add("apply_F  k_stokes<0,false>", 72, lambda: check(lib.mpbp_apply_F(p.h, b.data_ptr(), y5.data_ptr(), st)))
add("jacobi_F k_stokes<2,false>", 104, lambda: check(lib.mpbp_jacobi_F(p.h, b.data_ptr(), x4.data_ptr(), 10, 0.8, st)), div=10)
add("jacobi_P k_poisson<2>", 32, lambda: check(lib.mpbp_jacobi_P(p.h, bp_ptr, x1.data_ptr(), 10, 0.8, st)), div=10)
add("apply_GtG k_poisson<0>", 24, lambda: check(lib.mpbp_apply_GtG(p.h, bp_ptr, x1.data_ptr(), st)))
add("apply_D  k_div", 48, lambda: check(lib.mpbp_apply_D(p.h, b.data_ptr(), None, x1.data_ptr(), st)))
add("apply_G  k_grad", 48, lambda: check(lib.mpbp_apply_G(p.h, bp_ptr, y5.data_ptr(), st)))
add("apply_GtFG chain", 168, lambda: check(lib.mpbp_apply_GtFG(p.h, bp_ptr, x1.data_ptr(), st)))
r = C.c_double()
add("dot (5N)", 80, lambda: check(lib.mpbp_dot(p.h, b.data_ptr(), u.data_ptr(), 5 * N, C.byref(r), st)))
add("nrm2 (5N)", 40, lambda: check(lib.mpbp_nrm2(p.h, b.data_ptr(), 5 * N, C.byref(r), st)))
add("axpy (5N)", 120, lambda: check(lib.mpbp_axpy(p.h, 0.5, b.data_ptr(), y5.data_ptr(), 5 * N, st)))
pb = p.precond_bytes() / N
add("vcycle_F", 0, lambda: check(lib.mpbp_vcycle_F(p.h, b.data_ptr(), x4.data_ptr(), st)))
add("vcycle_P", 0, lambda: check(lib.mpbp_vcycle_P(p.h, bp_ptr, x1.data_ptr(), st)))
add("precond_apply", pb, lambda: check(lib.mpbp_precond_apply(p.h, b.data_ptr(), y5.data_ptr(), st)), div=1)
json.dump([dict(kernel=a, bytes_per_cell=bb, ms=c, gbs=d, frac=e) for a, bb, c, d, e in rows],
          open(os.path.join("gpurun_out", f"kernel_table_n{n}.json"), "w"), indent=1)
