#!/usr/bin/env python
"""bench.py -- GMRES iterations/s of the block-preconditioned two-phase Stokes solve (BASELINE.json
metric) with the roofline of the dominant kernel and the CPU oracle timed beside it.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                    (CPU arm: the oracle port on the host cores)

A "step" is one inner iteration of right-preconditioned FGMRES (solve.py:285): one preconditioner apply
(approx_schur_op, solve.py:257-277), one A.x, and the Arnoldi orthogonalisation at that basis size
(restart 40).  Workload: 2D 4096^2 MAC grid, 10^4 viscosity contrast (BASELINE.json configs[3]).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 must print exactly ONE line (the JSON): NCCL writes its "NCCL version ..." banner (NCCL_DEBUG=VERSION/WARN)
# to stdout unless told otherwise, so route NCCL's log to stderr before any communicator exists
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

WORKLOAD = dict(n=4096, eta_n=1.0e4, eta_s=1.0, xi=1.0, c=1.0, d_u=-1.0, restart=40)
# F: 6 / GtG: 2 Chebyshev-accelerated V(2,2) cycles: the configuration that converges at 4096^2, contrast 1e4
# (40 iterations to rtol 1e-8, profiles/r1_solve_configs.json; with 4 F-cycles it needs 318)
SUB = dict(kind="mg", F_cycles=6, P_cycles=2, cheb=True, nu1=2, nu2=2, omega=0.8, n_coarse=16)
METRIC = "gmres_iterations_per_second"
UNIT = "its/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy/scipy) on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------
def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _host_free_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def cpu_port_its_per_s(steps, warmup, n_sample=None):
    """Times `steps` FGMRES iterations of the CPU oracle port (same contrast, same sub-solver definition and
    restart) on the host cores.  The sample is the REAL 4096^2 workload whenever the host has the memory for it
    (about 20 GB: hierarchy + the few Krylov bases the bounded sample touches); only otherwise a smaller grid whose
    its/s is scaled by the cell ratio -- the line says which.  The port is the OpenMP C restatement
    (oracle/mpbp_oracle_c.c) on every core this process may use (torchrun's OMP_NUM_THREADS=1 is overridden)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import mpbp_oracle as O
    w = WORKLOAD
    if n_sample is None:
        free = _host_free_gb()
        n_sample = w["n"] if free >= 40 else (2048 if free >= 12 else 1024)
    scale = (n_sample / w["n"]) ** 2
    _, b = O.manufactured(n_sample, w["c"], w["d_u"], w["xi"], w["eta_n"], w["eta_s"])
    from c_oracle import COracle, set_threads
    threads = set_threads(_host_threads())
    co = COracle(n_sample, w["xi"], w["eta_n"], w["eta_s"], w["c"], w["d_u"], kind="mg", F_cycles=SUB["F_cycles"],
                 P_cycles=SUB["P_cycles"], cheb=True, omega=SUB["omega"], nu1=SUB["nu1"], nu2=SUB["nu2"],
                 n_coarse=SUB["n_coarse"], lmin=SUB.get("lmin", 0.75), lmax=SUB.get("lmax", 1.2))
    if warmup > 0:
        co.fgmres(b, tol=0.0, restart=w["restart"], maxiter=1)
    t0 = time.perf_counter()
    _, _, hist = co.fgmres(b, tol=0.0, restart=w["restart"], maxiter=steps)
    dt = time.perf_counter() - t0
    its = len(hist)
    what = "OpenMP C oracle (oracle/mpbp_oracle_c.c)"
    if n_sample == w["n"]:
        sample = (f"{its} FGMRES iterations (Arnoldi bases 1..{its}) of the {what} on the full {n_sample}^2 workload, "
                  f"{dt:.1f} s on {threads} host threads (host has {os.cpu_count()} cores); measured, not extrapolated")
    else:
        sample = (f"{its} FGMRES iterations of the {what} on a {n_sample}^2 grid ({dt:.1f} s on {threads} host threads; "
                  f"host has {os.cpu_count()} cores, {_host_free_gb():.0f} GB free: too little for 4096^2), its/s scaled "
                  f"by the cell ratio {scale:.5f} to 4096^2 (EXTRAPOLATED)")
    return dict(value=its / dt * scale, unit=UNIT, cores=threads, kind="port", sample=sample, steps=its, seconds=dt,
                ms_per_step=dt / its / scale * 1e3, n_sample=n_sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))  # a bounded sample: ~7 s per iteration at 4096^2 on 16 cores
    res = cpu_port_its_per_s(steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": res["steps"], "steps_requested": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus), "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def config_dict(ngpu):
    w = WORKLOAD
    return {"workload": f"2D {w['n']}^2 MAC grid, viscosity contrast eta_n/eta_s = 1e4, right-preconditioned FGMRES "
                        f"(restart {w['restart']}), BFBt block preconditioner with Chebyshev-accelerated V(2,2) "
                        f"multigrid sub-solves (F: {SUB['F_cycles']} cycles, GtG: {SUB['P_cycles']} cycles)",
            "n": w["n"], "unknowns": 5 * w["n"] ** 2, "eta_n": w["eta_n"], "eta_s": w["eta_s"], "restart": w["restart"],
            "sub_solver": SUB, "parallelism": f"row-slabs x{ngpu}" if ngpu > 1 else "single GPU",
            "l2_policy": "inputs_exceed_l2 (one 5N vector = 671 MB vs 126 MB L2)"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mp_block_preconditioners_b200 as mp
    from mp_block_preconditioners_b200._cabi import SIDE_RIGHT, check
    from mp_block_preconditioners_b200.solve import _krylov
    from mp_block_preconditioners_b200.utils import manufactured_device

    w = dict(WORKLOAD)
    if args.n:
        w["n"] = args.n
    n = w["n"]
    sub = mp.SubSolver(**SUB)
    bp = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=sub, distributed=world > 1)
    A, S, F, D, G = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])
    M = bp.approx_schur_operator(c=w["c"], d_u=w["d_u"])
    p = A.plan
    lib = p.lib
    N = p.N
    u_dev, b_dev = manufactured_device(p)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        """device time of fn() in ms: CUDA events on the launching stream, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    restart = w["restart"]

    def solve_steps(k):
        return _krylov(A, b_dev, M, None, 1e-8, restart, 10 ** 6, SIDE_RIGHT, force_iters=k)

    # ---- warm-up (>= 3 iterations) then the timed region ----
    solve_steps(max(3, args.warmup))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = p.launches
    ms_total = timed(lambda: solve_steps(args.steps))
    launches = p.launches - l0
    value = args.steps / (ms_total * 1e-3)

    # ---- dominant kernel: damped-Jacobi sweep on F at level 0 (k_stokes<2,false>), 104 N algorithmic bytes ----
    peak, peak_src = measured_peak()
    xF = torch.zeros(4 * N, dtype=torch.float64, device="cuda")
    bF = b_dev[:4 * N].contiguous()
    sweeps = 20

    def jac():
        check(lib.mpbp_jacobi_F(p.h, bF.data_ptr(), xF.data_ptr(), sweeps, 0.8, p.stream()))
    jac()
    ms_jac = timed(jac) / sweeps
    jac_bytes = 104.0 * N
    roof = {"bound": "hbm", "kernel": "k_stokes_x<IN 0, MODE 2, EP 0> (damped-Jacobi sweep on F, level 0)",
            "achieved": jac_bytes / (ms_jac * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "bytes_per_launch": jac_bytes, "ms_per_launch": ms_jac, "traffic": None}
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of THIS kernel from the committed ncu --set full capture
    # of the final code (profiles/r2_ncu_traffic.json, written by profiles/ncu_table.py); only quoted for the
    # configuration it was captured on
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            tr = json.load(f)
        if world == 1 and n == int(tr["n"]):
            roof["traffic"] = float(tr["dram_bytes_read"]) + float(tr["dram_bytes_write"])
            roof["traffic_source"] = tr.get("source", "profiles/r2_ncu_traffic.json")
    except Exception:
        pass
    roof["frac"] = roof["achieved"] / peak

    # ---- the preconditioner apply as a whole and the other hot kernels ----
    z = torch.empty_like(b_dev)

    def pc():
        check(lib.mpbp_precond_apply(p.h, b_dev.data_ptr(), z.data_ptr(), p.stream()))
    pc()
    reps = 3
    ms_pc = timed(lambda: [pc() for _ in range(reps)]) / reps
    pc_bytes = p.precond_bytes()

    def ax():
        check(lib.mpbp_apply_A(p.h, b_dev.data_ptr(), z.data_ptr(), p.stream()))
    ax()
    ms_ax = timed(lambda: [ax() for _ in range(10)]) / 10
    kernels = {
        "precond_apply": {"ms": ms_pc, "algorithmic_bytes": pc_bytes, "gbs": pc_bytes / ms_pc / 1e6,
                          "frac": pc_bytes / ms_pc / 1e6 / peak},
        "apply_A": {"ms": ms_ax, "algorithmic_bytes": 88.0 * N, "gbs": 88.0 * N / ms_ax / 1e6,
                    "frac": 88.0 * N / ms_ax / 1e6 / peak},
    }
    clocks = sampler.stop() if rank == 0 else None

    # ---- in-run parity check (outside every timed region): the record carries its own evidence ----
    parity = run_parity(mp, p, A, M, b_dev, z, n, w, sub, rank, world) if not args.no_parity else None

    # ---- end to end through the public Python API with HOST buffers: one FGMRES cycle per call ----
    del xF, z
    e2e_calls = max(1, min(2, args.steps // restart))
    torch.cuda.empty_cache()
    bh = torch.empty(5 * N, dtype=torch.float64, pin_memory=True)
    bh.copy_(b_dev)
    b_host = bh.numpy()

    def e2e_once():
        x, info = mp.fgmres(A, b_host, M=M, tol=0.0, restart=restart, maxiter=1, host=True)
        return x
    e2e_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        xh = e2e_once()
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": e2e_calls * restart / dt, "unit": UNIT,
           "h2d_bytes_per_step": 5 * N * 8 * world / restart, "d2h_bytes_per_step": 5 * N * 8 * world / restart,
           "note": f"host-to-host fgmres(A, b, M=...) calls of one {restart}-iteration cycle each: b copied H2D and x "
                   f"copied D2H inside every call ({5 * N * 8 * world} bytes each way per call)"}

    if rank == 0:
        cpu = cpu_port_its_per_s(2, 1) if (world == 1 and not args.no_cpu) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(world) if not args.n else {**config_dict(world), "n": n, "workload": f"override n={n}"},
                "roofline": roof, "kernels": kernels, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "parity": parity}
        if cpu:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        _emit(line)
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # leave without running CUDA/NCCL teardown from interpreter shutdown (a rank blocking in a
        # communicator destructor would hang the whole torchrun job)
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def run_parity(mp, p, A, M, b_dev, z, n, w, sub, rank, world):
    """Parity evidence carried by every bench line (never timed).
    N = 1: the GPU's approx_schur_op (solve.py:257-277) and A.x against the C oracle on a 1024^2 grid with the
           benchmarked contrast and sub-solver (the oracle finishes there in about a second);
    N > 1: the slab-distributed apply and A.x of the BENCHMARKED workload against a single-GPU plan of the same grid
           evaluated on the same global vector, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    if world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from c_oracle import COracle, set_threads
        import mpbp_oracle as O
        m = 1024
        set_threads(_host_threads())
        co = COracle(m, w["xi"], w["eta_n"], w["eta_s"], w["c"], w["d_u"],
                     **{k: v for k, v in SUB.items()})
        bp = mp.MultiphaseBlockPreconditioner(m, w["xi"], w["eta_n"], w["eta_s"], sub_solver=sub)
        Am = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])[0]
        Mm = bp.approx_schur_operator(c=w["c"], d_u=w["d_u"])
        rng = np.random.default_rng(1024)
        v = rng.standard_normal(5 * m * m)
        v[4 * m * m:] -= v[4 * m * m:].mean()
        _, b = O.manufactured(m, w["c"], w["d_u"], w["xi"], w["eta_n"], w["eta_s"])

        def nrel(a, b_):
            return float(np.abs(a - b_).max() / np.abs(b_).max())
        zb = co.precond(b)
        # the smooth manufactured rhs is an ill-conditioned input of M (BFBt amplifies rounding in the rough modes by
        # ~eta_n/h^2 relative to the smooth content): the oracle's own M b moves by `sens` under 1-ulp perturbations
        sens = max(nrel(co.precond(b * (1.0 + 1.2e-16 * rng.standard_normal(b.shape))), zb) for _ in range(3))
        out = {"against": "C oracle (oracle/mpbp_oracle_c.c), 1024^2, benchmarked contrast and sub-solver",
               "apply_A_relerr": nrel(Am @ v, co.apply_A(v)),
               "precond_apply_relerr": nrel(Mm @ v, co.precond(v)),
               "precond_apply_rhs_relerr": nrel(Mm @ b, zb),
               "precond_apply_rhs_oracle_sensitivity_1ulp": sens,
               "tolerance": {"apply_A": 1e-13, "precond_apply": 1e-9, "precond_apply_rhs": "max(1e-9, 10 x sensitivity)"}}
        out["ok"] = bool(out["apply_A_relerr"] < 1e-13 and out["precond_apply_relerr"] < 1e-9
                         and out["precond_apply_rhs_relerr"] < max(1e-9, 10.0 * sens))
        Am.plan.close()
        return out
    from mp_block_preconditioners_b200.parallel import scatter_slab
    from mp_block_preconditioners_b200.utils import manufactured_device
    bp1 = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=sub)
    A1 = bp1.get_big_A_matrix(c=w["c"], d_u=w["d_u"])[0]
    M1 = bp1.approx_schur_operator(c=w["c"], d_u=w["d_u"])
    _, b1 = manufactured_device(A1.plan)
    e_b = rel(b_dev, scatter_slab(b1, n, 5, rank, world))
    # a well-conditioned input: the same seeded random global vector on every rank (zero-mean pressure part)
    g = torch.Generator(device="cuda")
    g.manual_seed(4096)
    v1 = torch.randn(5 * n * n, dtype=torch.float64, device="cuda", generator=g)
    v1[4 * n * n:] -= v1[4 * n * n:].mean()
    vd = scatter_slab(v1, n, 5, rank, world)
    e_A = rel(A @ vd, scatter_slab(A1 @ v1, n, 5, rank, world))
    e_M = rel(M @ vd, scatter_slab(M1 @ v1, n, 5, rank, world))
    # the smooth manufactured rhs is an ill-conditioned input of M (see the N = 1 block): its sensitivity is measured on
    # the single-GPU plan with 1-ulp perturbations of b and bounds the slab-vs-single difference
    z1 = M1 @ b1
    e_Mb = rel(M @ b_dev, scatter_slab(z1, n, 5, rank, world))
    sens = 0.0
    for s_ in range(2):
        g.manual_seed(7 + s_)
        pert = 1.0 + 1.2e-16 * torch.randn(b1.numel(), dtype=torch.float64, device="cuda", generator=g)
        sens = max(sens, rel(M1 @ (b1 * pert), z1))
        del pert
    del z1, b1, v1, vd
    A1.plan.close()
    torch.cuda.empty_cache()
    t = torch.tensor([e_b, e_A, e_M, e_Mb], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e_b, e_A, e_M, e_Mb = (float(v) for v in t)
    return {"against": f"single-GPU plan of the same {n}^2 workload on the same global vectors (max over {world} ranks)",
            "rhs_relerr": e_b, "apply_A_relerr": e_A, "precond_apply_relerr": e_M,
            "precond_apply_rhs_relerr": e_Mb, "precond_apply_rhs_sensitivity_1ulp": sens,
            "tolerance": {"apply_A": 1e-13, "precond_apply": 1e-9, "precond_apply_rhs": "max(1e-9, 10 x sensitivity)"},
            "ok": bool(e_A < 1e-13 and e_M < 1e-9 and e_Mb < max(1e-9, 10.0 * sens))}


def run_apply_sweep(args):
    """BASELINE.json configs[4]: preconditioner-apply + SpMV throughput on an 8192^2 grid (contrast 1e4, bench
    sub-solver) across the ranks of the launch, with the halo-exchange and all-reduce latencies split out.
    value = algorithmic GB/s of the preconditioner apply summed over ranks (total bytes / max-over-ranks time)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mp_block_preconditioners_b200 as mp
    from mp_block_preconditioners_b200._cabi import check
    from mp_block_preconditioners_b200.utils import manufactured_device
    w = dict(WORKLOAD)
    n = args.n or 8192
    bp = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=mp.SubSolver(**SUB),
                                          distributed=world > 1)
    A = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])[0]
    p, lib = A.plan, A.plan.lib
    N = p.N
    u_dev, b_dev = manufactured_device(p)
    z = torch.empty_like(b_dev)
    stream = torch.cuda.current_stream()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        check(lib.mpbp_precond_apply(p.h, b_dev.data_ptr(), z.data_ptr(), p.stream()))
    l0 = p.launches
    steps = max(1, min(args.steps, 10))
    ms_pc = timed(lambda: check(lib.mpbp_precond_apply(p.h, b_dev.data_ptr(), z.data_ptr(), p.stream())), steps)
    launches = p.launches - l0
    ms_ax = timed(lambda: check(lib.mpbp_apply_A(p.h, b_dev.data_ptr(), z.data_ptr(), p.stream())), 20)
    clocks = sampler.stop() if rank == 0 else None
    hu, au = C.c_double(0.0), C.c_double(0.0)
    check(lib.mpbp_comm_probe(p.h, 200, C.byref(hu), C.byref(au), p.stream()))
    peak, peak_src = measured_peak()
    pc_bytes = p.precond_bytes() * world          # every rank moves its slab's share
    ax_bytes = 88.0 * N * world
    if rank == 0:
        gbs = pc_bytes / ms_pc / 1e6
        line = {"metric": "precond_apply_GBps", "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_pc, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"2D {n}^2 MAC grid, contrast 1e4: preconditioner apply (approx_schur_op) + A.x "
                                       f"throughput sweep, row slabs over {world} GPU(s)", "n": n, "sub_solver": SUB,
                           "l2_policy": f"inputs_exceed_l2 (one 5N vector = {5 * n * n * 8 / 1e9:.2f} GB)"},
                "roofline": {"bound": "hbm", "kernel": "precond_apply (all kernels)", "achieved": gbs / world,
                             "peak": peak, "unit": "GB/s per GPU", "peak_source": peak_src, "frac": gbs / world / peak,
                             "traffic": None, "bytes_per_launch": pc_bytes / world, "ms_per_launch": ms_pc},
                "kernels": {"precond_apply": {"ms": ms_pc, "algorithmic_bytes": pc_bytes, "gbs": gbs,
                                              "frac_per_gpu": gbs / world / peak},
                            "apply_A": {"ms": ms_ax, "algorithmic_bytes": ax_bytes, "gbs": ax_bytes / ms_ax / 1e6,
                                        "frac_per_gpu": ax_bytes / ms_ax / 1e6 / world / peak}},
                "comm": {"halo_exchange_us": hu.value, "allreduce_us": au.value,
                         "note": "one level-0 halo exchange of a 5-field vector (push kernel + fetching consumer kernel, "
                                 "LL peer-memory protocol) and one scalar all-reduce (one-block reduction kernel with the "
                                 "fused peer-memory all-reduce), 200 repetitions each replayed from a CUDA graph"},
                "clocks": clocks, "gpu_launches": int(launches)}
        _emit(line)
    sys.stdout.flush()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


_REAL_STDOUT = None


def _capture_stdout():
    """Everything any library prints to fd 1 during the run (NCCL banners, warnings) goes to stderr; the one
    JSON line is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override the grid size (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block")
    ap.add_argument("--workload", default="solve4096", choices=["solve4096", "apply8192"],
                    help="solve4096: the headline metric (default); apply8192: BASELINE configs[4] throughput sweep")
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "apply8192":
        run_apply_sweep(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
