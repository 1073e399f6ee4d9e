#!/usr/bin/env python3
"""Condense an `ncu --page raw --csv` export of the two traversal launches of a config-5 wave into profiles/ncu_config5_summary.json
(read by bench.py for the issue-slot roofline: warp instructions per ray x live rays/s against 4 x #SM x SM clock).
usage: python profiles/summarize_ncu_cfg.py <raw.csv> <rays of the closest-hit launch> <rays of the any-hit launch> <source note> [out.json]"""
import csv, json, os, sys

M = {"duration_us": "gpu__time_duration.sum", "warp_inst": "smsp__inst_executed.sum", "warp_inst_issued": "smsp__inst_issued.sum",
     "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "registers": "launch__registers_per_thread", "grid": "launch__grid_size", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
     "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum", "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
     "branch_uniform_pct": "smsp__sass_average_branch_targets_threads_uniform.pct", "cycles_elapsed_max": "sm__cycles_elapsed.max",
     "local_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_st_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    rays = {"closest": float(sys.argv[2]), "shadow": float(sys.argv[3]), "shade": float(sys.argv[2])}      # k_shade (if captured) shades the closest-hit wave
    out = {"source": sys.argv[4], "kernels": {}}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        which = "shade" if "k_shade" in name else ("shadow" if "k_traverse_dyn<1" in name or "k_traverse_dyn<true" in name else ("closest" if "k_traverse_dyn" in name else None))
        if which is None:
            continue
        d = {"kernel": name.split("(")[0]}
        for k, m in M.items():
            if m in hdr and r[hdr.index(m)]:
                i = hdr.index(m)
                d[k] = float(r[i].replace(",", "")) * (UNIT.get(units[i], 1) if k in ("duration_us", "dram_read_bytes", "dram_write_bytes") else 1)
        d["rays"] = rays[which]
        d["warp_inst_per_ray"] = d["warp_inst_issued"] / rays[which]
        d["traffic_bytes_per_launch"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        d["dram_bytes_per_ray"] = d["traffic_bytes_per_launch"] / rays[which]
        d["mrays_per_s"] = rays[which] / d["duration_us"]
        d["issue_slot_frac_chip"] = d["warp_inst_issued"] / (d["cycles_elapsed_max"] * 4 * 148)
        stalls = sorted(((float(r[i].replace(",", "")), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for i in range(len(hdr)) if "smsp__average_warps_issue_stalled" in hdr[i] and hdr[i].endswith("per_issue_active.ratio") and r[i]), reverse=True)
        d["top_stalls_warps_per_issue"] = {h: v for v, h in stalls[:6]}
        out["kernels"][which] = d
    path = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_config5_summary.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
