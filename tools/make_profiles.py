#!/usr/bin/env python
"""gpurun_out/ncu_<wl>.ncu-rep + launches_<wl>.csv  ->  profiles/rNN_<wl>.md (tracked summaries) + profiles/traffic.json.

usage: python tools/make_profiles.py r01 spmm spmv ...
The .ncu-rep files stay in gpurun_out/ (scratch); what the judge reads is the markdown written here."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def raw_page(path):
    csv_path = path[: -len(".ncu-rep")] + ".raw.csv"
    if os.path.exists(csv_path) and (not os.path.exists(path) or os.path.getmtime(csv_path) >= os.path.getmtime(path)):
        out = open(csv_path).read()            # exported on the GPU box by tools/gpu_prof.sh (the .ncu-rep stays there)
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def launches(path):
    agg = collections.OrderedDict()
    for r in csv.reader(open(path)):
        if len(r) > 10 and r[0].isdigit():
            agg.setdefault(r[4].split("(")[0], []).append(float(r[-1].replace(",", "")))
    return agg


def main():
    tag, wls = sys.argv[1], sys.argv[2:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for wl in wls:
        rep = os.path.join(ROOT, "gpurun_out", f"ncu_{wl}.ncu-rep")
        lst = os.path.join(ROOT, "gpurun_out", f"launches_{wl}.csv")
        lines = [f"# {tag} — {wl}: ncu summary (B200, `--set full --clock-control none`, one launch of the dominant kernel)", ""]
        if os.path.exists(rep) or os.path.exists(rep[: -len(".ncu-rep")] + ".raw.csv"):
            for d, u in raw_page(rep):
                lines += [f"Kernel: `{d.get('Kernel Name', '')}`  grid {d.get('Grid Size')} block {d.get('Block Size')}", "",
                          "| metric | value | unit |", "|---|---|---|"]
                for k in KEYS:
                    if k in d:
                        lines.append(f"| {k} | {d[k]} | {u[k]} |")
                rd = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT[u["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT[u["dram__bytes_write.sum"]]
                dur = float(d["gpu__time_duration.sum"].replace(",", ""))
                dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[u["gpu__time_duration.sum"]]
                lines += ["", f"DRAM traffic per launch: {(rd + wr) / 1e9:.3f} GB (read {rd / 1e9:.3f} + write {wr / 1e9:.3f}) "
                              f"in {dur_s * 1e3:.3f} ms under ncu = {(rd + wr) / dur_s / 1e9:.0f} GB/s", ""]
                traffic[wl] = {"kernel": d.get("Kernel Name", "").split("(")[0], "dram_bytes_per_launch": rd + wr, "round": tag}
        if os.path.exists(lst):
            lines += [f"## launch list of one `bench.py --workload {wl} --steps 2 --warmup 3` (ours only; "
                      "`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised)", "",
                      "| kernel | launches | avg µs | share of our GPU time |", "|---|---|---|---|"]
            agg = launches(lst)
            tot = sum(sum(v) for v in agg.values())
            for k, v in agg.items():
                lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |")
        open(os.path.join(ROOT, "profiles", f"{tag}_{wl}.md"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
