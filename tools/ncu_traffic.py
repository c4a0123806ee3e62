"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep, read here with `ncu -i ... --page raw --csv`) into
profiles/: one CSV row per profiled launch (duration, DRAM bytes, throughputs, occupancy, top stall reasons) and
profiles/r02_traffic.json = measured DRAM bytes per cell and launch of every kernel, which bench.py scales by the
cell count for `roofline.traffic` (with the source file named in the record).
usage: ncu_traffic.py <cells of the profiled mesh> <tag> <rep> [<rep> ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6,
        "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("impes::", "").strip()
    return n


def main():
    ncell, tag, reps = int(sys.argv[1]), sys.argv[2], sys.argv[3:]
    rows_out, traffic = [], {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            rec = {"file": os.path.basename(rep), "kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["launch__grid_size"]]}
            for k in KEYS:
                if k in idx:
                    v = float(r[idx[k]].replace(",", "") or 0) * UNIT.get(units[idx[k]], 1.0)
                    rec[k] = v
            rows_out.append(rec)
            name = rec["kernel"].split("<")[0]
            variant = rec["kernel"]
            b = rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0)
            t = traffic.setdefault(variant, {"bytes": [], "us": [], "source": f"profiles/{tag}_ncu_launches.csv ({os.path.basename(rep)}, {ncell} cells)"})
            t["bytes"].append(b)
            t["us"].append(rec.get("gpu__time_duration.sum", 0))
            t.setdefault("fp64", []).append(rec.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "kernel", "grid"] + KEYS)
        for rec in rows_out:
            w.writerow([rec["file"], rec["kernel"], rec["grid"]] + [f"{rec.get(k, 0):.6g}" for k in KEYS])
    table = {}
    for variant, t in traffic.items():
        # keep the FULL variant of kernels that also have a cheap one (first jacobi sweep): the largest traffic
        b = max(t["bytes"])
        i = t["bytes"].index(b)
        table[variant] = {"bytes_per_cell": b / ncell, "us_at_profile": t["us"][i], "fp64_pipe_pct": round(t["fp64"][i], 1),
                          "source": t["source"]}
    with open(os.path.join(ROOT, "profiles", f"{tag}_traffic_by_variant.json"), "w") as f:
        json.dump(table, f, indent=1)
    # names as bench.py's rooflines use them
    by = {}
    for variant, rec in table.items():
        base = variant.split("<")[0]
        if base not in by or rec["bytes_per_cell"] > by[base]["bytes_per_cell"]:
            by[base] = rec
    out = {}
    for name in ("k1q_sat", "k1_mg_jacobi", "k1_mg_restrict", "k1_darcy_apply", "k_darcy_apply", "k_sat_stream"):
        if name in by:
            out[name] = by[name]
    if "k1q_apply_a" in by and "k1q_apply_b" in by:
        out["k1q_apply_a+b"] = {"bytes_per_cell": by["k1q_apply_a"]["bytes_per_cell"] + by["k1q_apply_b"]["bytes_per_cell"],
                                "fp64_pipe_pct": by["k1q_apply_a"]["fp64_pipe_pct"], "source": by["k1q_apply_a"]["source"]}
    bk = [k for k in by if k.startswith("k_minres_b")]
    if bk and "k_minres_c" in by:
        out["k_minres_b+c"] = {"bytes_per_cell": by[bk[0]]["bytes_per_cell"] + by["k_minres_c"]["bytes_per_cell"],
                               "source": by["k_minres_c"]["source"]}
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(f"{k}: {v['bytes_per_cell']:.1f} B/cell")


if __name__ == "__main__":
    main()
