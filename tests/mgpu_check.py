"""Multi-GPU slab parity check, launched under torchrun (one rank per GPU):
every rank applies the distributed operators / preconditioner / GMRES to its row slab and the
results are compared with the single-GPU plan evaluated on the same global vectors.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py [n]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mp_block_preconditioners_b200 as mp
    from mp_block_preconditioners_b200.parallel import gather_slabs, scatter_slab
    from mp_block_preconditioners_b200.preconditioner import DivergenceOperator, GtFGOperator, GtGOperator
    from mp_block_preconditioners_b200.solve import _krylov
    from mp_block_preconditioners_b200._cabi import SIDE_LEFT, SIDE_RIGHT

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    xi, eta_n, eta_s, c, d = 1.0, 100.0, 1.0, 1.0, -1.0
    dmin = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    # third argument "nokrylov": operators, V-cycles and the apply only.  Used for the 256 / dist_min_n = 16 stress
    # hierarchy, whose small distributed levels consist almost entirely of slab-edge rows (general instantiation of the
    # marching kernels, other multiply-add contraction): the eight single-GPU variants that build the history envelope
    # below cannot reproduce that, and the histories are compared on the default hierarchy instead (1024, 0).
    with_krylov = not (len(sys.argv) > 3 and sys.argv[3] == "nokrylov")
    sub = mp.SubSolver(kind="mg", F_cycles=3, P_cycles=2, cheb=True, dist_min_n=dmin)
    ok = True

    def report(name, err, tol):
        nonlocal ok
        t = torch.tensor([err], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        good = float(t) < tol
        ok = ok and good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} {name}: max err {float(t):.3e} (tol {tol:.0e})", flush=True)

    # global reference on every rank with a single-GPU plan
    bp1 = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub)
    A1, _, F1, D1, G1 = bp1.get_big_A_matrix(c, d)
    M1 = bp1.approx_schur_operator(c, d)
    bpd = mp.MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s, sub_solver=sub, distributed=True)
    Ad, _, Fd, Dd, Gd = bpd.get_big_A_matrix(c, d)
    Md = bpd.approx_schur_operator(c, d)
    p = Ad.plan
    assert p.rows == n // world
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(5 * n * n, dtype=torch.float64, device="cuda", generator=gen)
    x[4 * n * n:] -= x[4 * n * n:].mean()
    xs = scatter_slab(x, n, 5, rank, world)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    y1 = A1 @ x
    report("apply_A", rel(Ad @ xs, scatter_slab(y1, n, 5, rank, world)), 1e-13)
    N = n * n
    report("apply_F", rel(Fd @ xs[:4 * p.N], scatter_slab(F1 @ x[:4 * N], n, 4, rank, world)), 1e-13)
    report("apply_D", rel(Dd @ xs[:4 * p.N], scatter_slab(D1 @ x[:4 * N], n, 1, rank, world)), 1e-13)
    report("apply_G", rel(Gd @ xs[4 * p.N:], scatter_slab(G1 @ x[4 * N:], n, 4, rank, world)), 1e-13)
    report("apply_GtG", rel(GtGOperator(p) @ xs[4 * p.N:], scatter_slab(GtGOperator(A1.plan) @ x[4 * N:], n, 1, rank, world)), 1e-13)
    report("apply_GtFG", rel(GtFGOperator(p) @ xs[4 * p.N:], scatter_slab(GtFGOperator(A1.plan) @ x[4 * N:], n, 1, rank, world)), 1e-12)
    v1 = A1.plan.call("mpbp_vcycle_F", x[:4 * N], 4 * N, 4 * N)
    report("vcycle_F", rel(p.call("mpbp_vcycle_F", xs[:4 * p.N], 4 * p.N, 4 * p.N), scatter_slab(v1, n, 4, rank, world)), 1e-11)
    v1 = A1.plan.call("mpbp_vcycle_P", x[4 * N:], N, N)
    report("vcycle_P", rel(p.call("mpbp_vcycle_P", xs[4 * p.N:], p.N, p.N), scatter_slab(v1, n, 1, rank, world)), 1e-11)
    z1 = M1 @ x
    zd = Md @ xs
    report("precond_apply", rel(zd, scatter_slab(z1, n, 5, rank, world)), 1e-10)
    zg = gather_slabs(zd, n, 5, world)
    report("gather_roundtrip", rel(zg, z1), 1e-10)
    # Krylov: same residual history and iterate as the single-GPU solve
    from mp_block_preconditioners_b200.utils import manufactured_device
    u1, b1 = manufactured_device(A1.plan)
    ud, bd = manufactured_device(p)
    report("manufactured_rhs", rel(bd, scatter_slab(b1, n, 5, rank, world)), 1e-15)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import hist_check

    def krylov_pair(tag, eta, A1_, M1_, Ad_, Md_, b1_, bd_, sides):
        # conditioning of the history: the same single-GPU solve with other (still deterministic) summation orders in
        # the dot products -- the only arithmetic difference a slab run introduces.  Envelope over eight of them.
        # A slab run also changes WHICH code body computes a row (the first / last strip of a slab runs the general
        # instantiation of the marching kernels, whose fused multiply-adds are contracted differently: V-cycles differ
        # by an ulp), so two of the variants also use other strip heights.
        variants = []
        for rb, rs in (("211", None), ("307", "6"), ("401", "10"), ("593", None), ("149", "14"), ("257", None),
                       ("449", "22"), ("1009", "8")):
            os.environ["MPBP_RED_BLOCKS"] = rb
            if rs:
                os.environ["MPBP_RS"] = rs
            bps = mp.MultiphaseBlockPreconditioner(n, xi, eta, eta_s, sub_solver=sub)
            variants.append((bps.get_big_A_matrix(c, d)[0], bps.approx_schur_operator(c, d)))
            os.environ.pop("MPBP_RS", None)
        del os.environ["MPBP_RED_BLOCKS"]
        for side, name in sides:
            mi = 60 if side == SIDE_RIGHT else 5
            xa, ia, ha = _krylov(A1_, b1_, M1_, None, 1e-8, 30, mi, side)
            xb, ib, hb = _krylov(Ad_, bd_, Md_, None, 1e-8, 30, mi, side)
            env = np.zeros(len(ha))
            xdev = 0.0
            for As, Ms in variants:
                xs_, is_, hs_ = _krylov(As, b1_, Ms, None, 1e-8, 30, mi, side)
                k = min(len(ha), len(hs_))
                env[:k] = np.maximum(env[:k], np.abs(ha[:k] - hs_[:k]) / ha[:k])
                if len(hs_) != len(ha):
                    env[k:] = np.inf
                xdev = max(xdev, rel(xs_, xa))
            good = 1.0
            try:
                if rank == 0:
                    hist_check(hb, ha, env, label=f"{tag}{name} slab vs single GPU")
                else:
                    hist_check(hb, ha, env, verbose=False)
            except AssertionError as exc:
                good = 0.0
                if rank == 0:
                    print("   " + str(exc)[:400], flush=True)
            report(f"{tag}{name}_history (1e-10 rel, 10x reorder envelope, +-1 iteration)", 1.0 - good, 0.5)
            report(f"{tag}{name}_same_info", float(abs(ia - ib)), 0.5)
            xtol = max(1e-6, 1e1 * xdev)
            report(f"{tag}{name}_solution", rel(xb, scatter_slab(xa, n, 5, rank, world)), xtol)

    # eta_n = 100: right-preconditioned FGMRES (the left-preconditioned history is ill conditioned at this
    # contrast -- the oracle itself moves by tens of percent under 1-ulp perturbations, tests/golden)
    if not with_krylov:
        if rank == 0:
            print("MGPU_ALL_PASS" if ok else "MGPU_FAILED", flush=True)
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0 if ok else 1)
    krylov_pair("eta100_", eta_n, A1, M1, Ad, Md, b1, bd, ((SIDE_RIGHT, "fgmres"),))
    # eta_n = 1: both Krylov variants
    bp1b = mp.MultiphaseBlockPreconditioner(n, xi, 1.0, eta_s, sub_solver=sub)
    bpdb = mp.MultiphaseBlockPreconditioner(n, xi, 1.0, eta_s, sub_solver=sub, distributed=True)
    A1b, Adb = bp1b.get_big_A_matrix(c, d)[0], bpdb.get_big_A_matrix(c, d)[0]
    M1b, Mdb = bp1b.approx_schur_operator(c, d), bpdb.approx_schur_operator(c, d)
    _, b1b = manufactured_device(A1b.plan)
    _, bdb = manufactured_device(Adb.plan)
    krylov_pair("eta1_", 1.0, A1b, M1b, Adb, Mdb, b1b, bdb, ((SIDE_RIGHT, "fgmres"), (SIDE_LEFT, "gmres_left")))
    if rank == 0:
        print("MGPU_ALL_PASS" if ok else "MGPU_FAILED", flush=True)
    sys.stdout.flush()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
