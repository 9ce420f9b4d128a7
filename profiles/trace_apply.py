"""Kernel timeline of the preconditioner apply on rank 0 of an N-rank run (CUPTI activity records through
torch.profiler; works with the whole-apply CUDA graph).  Aggregates, per (kernel, grid), the busy time and the idle
gap in front of each launch -- the evidence for where a slab-distributed apply spends its time.

usage: [torchrun ...] python profiles/trace_apply.py [n] [tag]       -> gpurun_out/trace_apply_<tag>.json/.txt
Numbers taken under the profiler are diagnostics, never bench values."""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch
import torch.distributed as dist

import bench

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import mp_block_preconditioners_b200 as mp
from mp_block_preconditioners_b200._cabi import check
from mp_block_preconditioners_b200.utils import manufactured_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tag = sys.argv[2] if len(sys.argv) > 2 else f"{world}gpu"
w = bench.WORKLOAD
bp = mp.MultiphaseBlockPreconditioner(n, w["xi"], w["eta_n"], w["eta_s"], sub_solver=mp.SubSolver(**bench.SUB),
                                      distributed=world > 1)
A = bp.get_big_A_matrix(c=w["c"], d_u=w["d_u"])[0]
p, lib = A.plan, A.plan.lib
u, b = manufactured_device(p)
z = torch.empty_like(b)


def pc():
    check(lib.mpbp_precond_apply(p.h, b.data_ptr(), z.data_ptr(), p.stream()))


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for _ in range(3):
    pc()
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    pc()
e1.record()
sync()
ms_plain = e0.elapsed_time(e1) / 3

from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pc()
    pc()
    sync()
path = f"/tmp/trace_{rank}.json"
prof.export_chrome_trace(path)
if rank == 0:
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    # the second apply only: everything after the midpoint k_combine
    comb = [i for i, e in enumerate(ev) if "k_combine" in e["name"]]
    if len(comb) >= 2:
        ev = ev[comb[-2] + 1: comb[-1] + 1]

    def short(nm):
        nm = re.sub(r"^void ", "", nm)
        nm = re.sub(r"\(.*$", "", nm)
        return nm.replace("mpbp::", "")[:70]
    agg = {}
    prev_end = None
    t_first, t_last = ev[0]["ts"], ev[-1]["ts"] + ev[-1]["dur"]
    for e in ev:
        g = e.get("args", {}).get("grid", [0, 0, 0])
        key = (short(e["name"]), tuple(g))
        a = agg.setdefault(key, dict(count=0, busy_us=0.0, gap_us=0.0))
        a["count"] += 1
        a["busy_us"] += e["dur"]
        if prev_end is not None:
            a["gap_us"] += max(0.0, e["ts"] - prev_end)
        prev_end = max(prev_end or 0.0, e["ts"] + e["dur"])
    rows = sorted(((k, v) for k, v in agg.items()), key=lambda kv: -(kv[1]["busy_us"] + kv[1]["gap_us"]))
    span = t_last - t_first
    busy = sum(v["busy_us"] for _, v in rows)
    gaps = sum(v["gap_us"] for _, v in rows)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/trace_apply_{tag}.txt", "w") as f:
        f.write(f"n={n} ranks={world}: apply {ms_plain:.3f} ms without the profiler; traced span {span / 1e3:.3f} ms, "
                f"kernel busy {busy / 1e3:.3f} ms, idle gaps {gaps / 1e3:.3f} ms, {len(ev)} launches\n")
        f.write(f"{'kernel':72s} {'grid':>16s} {'count':>6s} {'busy us':>10s} {'avg us':>8s} {'gap us':>10s} {'avg gap':>8s}\n")
        for (nm, g), v in rows:
            f.write(f"{nm:72s} {str(list(g)):>16s} {v['count']:6d} {v['busy_us']:10.1f} {v['busy_us'] / v['count']:8.2f} "
                    f"{v['gap_us']:10.1f} {v['gap_us'] / v['count']:8.2f}\n")
    json.dump(dict(n=n, ranks=world, ms_plain=ms_plain, span_us=span, busy_us=busy, gap_us=gaps,
                   rows=[dict(kernel=k[0], grid=list(k[1]), **v) for k, v in rows]),
              open(f"gpurun_out/trace_apply_{tag}.json", "w"), indent=1)
    print(open(f"gpurun_out/trace_apply_{tag}.txt").read()[:6000])
sys.stdout.flush()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)
