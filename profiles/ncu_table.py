"""Compact per-kernel table from an .ncu-rep: for every distinct kernel name the launch with the LARGEST grid (the
level-0 instance of a multigrid kernel), with the metrics the roofline discussion uses.
    python profiles/ncu_table.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] [--traffic profiles/r2_ncu_traffic.json]
--traffic writes dram__bytes_read/write of the level-0 damped-Jacobi sweep (bench.py's roofline kernel) for bench.py."""
import csv
import json
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rdMB", None), ("dram__bytes_write.sum", "wrMB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("launch__registers_per_thread", "regs", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1),
        ("lts__t_sector_hit_rate.pct", "L2hit%", 1), ("l1tex__t_sector_hit_rate.pct", "L1hit%", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long", 1),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", 1),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", 1),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", 1),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_nsel", 1),
        ("smsp__inst_executed.sum", "Minst", 1e-6)]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    out = {}
    args = list(sys.argv[1:])
    traffic = None
    if "--traffic" in args:
        i = args.index("--traffic")
        traffic = args[i + 1]
        del args[i:i + 2]
    for rep in args:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        best = {}
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
            g = r[idx["Grid Size"]]
            size = 1
            for t in re.findall(r"\d+", g):
                size *= int(t)
            if name not in best or size > best[name][0]:
                best[name] = (size, r)
        for name, (size, r) in best.items():
            rec = {"grid": r[idx["Grid Size"]]}
            for key, short, scale in COLS:
                if key not in idx:
                    continue
                if scale is None:
                    rec[short] = to_bytes(r[idx[key]], units[idx[key]]) / 1e6
                else:
                    v = float(r[idx[key]].replace(",", ""))
                    if key == "gpu__time_duration.sum":
                        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[idx[key]], 1e-3)
                        scale = 1
                    rec[short] = v * scale
            out[name] = rec
    names = [s for _, s, _ in COLS]
    print(f"{'kernel':58s} " + " ".join(f"{s:>8s}" for s in names))
    for name, rec in out.items():
        print(f"{name[:58]:58s} " + " ".join(f"{rec.get(s, float('nan')):8.1f}" for s in names) + "  grid " + rec["grid"])
    json.dump(out, open(args[0] + ".table.json", "w"), indent=1)
    if traffic:
        for name, rec in out.items():
            if re.search(r"k_stokes_x<0, 2, (false|0), 0, (false|0)", name) and "rdMB" in rec:
                g = [int(t) for t in re.findall(r"\d+", rec["grid"])]
                json.dump({"kernel": name, "grid": rec["grid"], "n": 4096 if g[:2] == [35, 128] else None,
                           "dram_bytes_read": rec["rdMB"] * 1e6, "dram_bytes_write": rec["wrMB"] * 1e6,
                           "time_us_under_ncu": rec.get("us"), "source": "ncu --set full capture " + " ".join(args)},
                          open(traffic, "w"), indent=1)
                break


if __name__ == "__main__":
    main()
