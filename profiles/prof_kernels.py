"""Profiling driver (run under ncu via gpurun): at the bench workload (4096^2, contrast 1e4) launches
  1. three level-0 damped-Jacobi sweeps on F   (k_stokes<2,false>: the dominant kernel)
  2. one A.x                                    (k_stokes<0,true>)
  3. one GtG Jacobi sweep, one D, one G apply
  4. one full preconditioner apply + one A.x + dot/axpy (one GMRES step's worth of kernels)
so that `-k regex:k_stokes -c 3` captures the dominant kernel and the launch list covers a whole step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import mp_block_preconditioners_b200 as mp
from mp_block_preconditioners_b200._cabi import check
from mp_block_preconditioners_b200.utils import manufactured_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.WORKLOAD["n"]
w = bench.WORKLOAD
bp = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=mp.SubSolver(**bench.SUB))
A, S, F, D, G = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])
M = bp.approx_schur_operator(c=w["c"], d_u=w["d_u"])
p, lib = A.plan, A.plan.lib
N = p.N
u, b = manufactured_device(p)
x = torch.zeros(4 * N, dtype=torch.float64, device="cuda")
z = torch.empty_like(b)
torch.cuda.synchronize()
# everything before this point (plan creation: ~1300 launches for the 16x16 dense coarse inverse) is not profiled
# when ncu runs with --profile-from-start off
torch.cuda.profiler.start()
check(lib.mpbp_jacobi_F(p.h, b.data_ptr(), x.data_ptr(), 3, 0.8, p.stream()))
check(lib.mpbp_apply_A(p.h, b.data_ptr(), z.data_ptr(), p.stream()))
check(lib.mpbp_jacobi_P(p.h, b.data_ptr() + 4 * N * 8, x.data_ptr(), 1, 0.8, p.stream()))
check(lib.mpbp_apply_D(p.h, b.data_ptr(), None, x.data_ptr(), p.stream()))
check(lib.mpbp_apply_G(p.h, b.data_ptr() + 4 * N * 8, x.data_ptr(), p.stream()))
torch.cuda.synchronize()
l0 = p.launches
check(lib.mpbp_precond_apply(p.h, b.data_ptr(), z.data_ptr(), p.stream()))
check(lib.mpbp_apply_A(p.h, z.data_ptr(), u.data_ptr(), p.stream()))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in one precond apply + A.x:", p.launches - l0)
