#!/usr/bin/env python3
"""Condense `ncu --page raw --csv` exports into the tracked summaries under profiles/.
usage: python profiles/summarize_ncu.py <traverse_raw.csv> [<shade_raw.csv>]   (run where the .ncu-rep files were read)"""
import csv, json, sys, os

KEYS = [
    ("duration_us", "gpu__time_duration.sum", None), ("warp_inst", "smsp__inst_executed.sum", 1), ("thread_inst", "smsp__thread_inst_executed.sum", 1),
    ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("achieved_occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("registers", "launch__registers_per_thread", 1),
    ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct", 1), ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1),
    ("dram_read_bytes", "dram__bytes_read.sum", None), ("dram_write_bytes", "dram__bytes_write.sum", None),
    ("dram_throughput_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("branch_uniform_pct", "smsp__sass_average_branch_targets_threads_uniform.pct", 1),
    ("cycles_active_avg", "smsp__cycles_active.avg", 1), ("cycles_elapsed_max", "sm__cycles_elapsed.max", 1),
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for name, metric, scale in KEYS:
            if metric not in hdr:
                continue
            i = hdr.index(metric)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            d[name] = v * (UNIT.get(units[i], 1) if scale is None else scale)
        out.append(d)
    return out


if __name__ == "__main__":
    trav = load(sys.argv[1])
    labels = ["closest-hit wave 0 (2 073 600 primary rays)", "any-hit wave 0 (shadow rays of the primary hits)", "closest-hit wave 1 (mirror / dielectric children)", "any-hit wave 1"]
    for d, l in zip(trav, labels):
        d["launch"] = l
    closest = [d for d in trav if "<0" in d["kernel"] or "false" in d["kernel"] or d.get("launch", "").startswith("closest")]
    closest = [d for d in trav if d.get("launch", "").startswith("closest")]
    summary = {
        "source": "profiles/%s (ncu --set full --clock-control none, DT_SYNC_WAVES=1 so each kernel runs alone; one config-2 frame)" % os.path.basename(sys.argv[1]),
        "traffic_bytes_per_launch": sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in closest) / max(1, len(closest)),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the %d captured closest-hit launches (waves 0 and 1 of 7)" % len(closest),
        "ncu": {"launches": trav},
    }
    if len(sys.argv) > 2:
        sh = load(sys.argv[2])
        for d, l in zip(sh, ["shade wave 0", "shade wave 1"]):
            d["launch"] = l
        summary["ncu"]["shade"] = sh
    json.dump(summary, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traverse_summary.json"), "w"), indent=1)
    for d in trav + summary["ncu"].get("shade", []):
        print(json.dumps(d))
