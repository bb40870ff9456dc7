#!/usr/bin/env python
"""gpurun_out/{step,ncu,launches}_<wl>.csv (tools/gpu_prof2.sh)  ->  profiles/<tag>_<wl>.md + profiles/traffic.json.

usage: python tools/make_profiles2.py r02 <commit> spmm spmv mttkrp ...
traffic.json[wl] = {dram_bytes_per_launch (summed over every kernel of one step), kernel (the kernels of the step), prof_name (the
library's profile slot bench.py reads), round, commit}: bench.py refuses the figure when prof_name does not match what it launched."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def step_rows(path):
    """ncu --csv log with several metrics: one line per (launch, metric) -> list of launches {name, metric: value_in_base_units}"""
    out = collections.OrderedDict()
    hdr = None
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Metric Name" in r:
                hdr = {h: i for i, h in enumerate(r)}
            continue
        if len(r) < len(hdr):
            continue
        lid = r[hdr["ID"]]
        d = out.setdefault(lid, {"name": r[hdr["Kernel Name"]].split("(")[0]})
        v = float(r[hdr["Metric Value"]].replace(",", "") or 0)
        d[r[hdr["Metric Name"]]] = v * UNIT.get(r[hdr["Metric Unit"]], 1.0)
    return list(out.values())


def main():
    tag, commit, wls = sys.argv[1], sys.argv[2], sys.argv[3:]
    import bench
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for wl in wls:
        g = os.path.join(ROOT, "gpurun_out")
        lines = [f"# {tag} — {wl}: ncu summary (B200, `--clock-control none`, commit {commit})", ""]
        sp = os.path.join(g, f"step_{wl}.csv")
        if os.path.exists(sp):
            rows = step_rows(sp)
            tot_t = sum(r.get("gpu__time_duration.sum", 0) for r in rows)
            tot_rd = sum(r.get("dram__bytes_read.sum", 0) for r in rows)
            tot_wr = sum(r.get("dram__bytes_write.sum", 0) for r in rows)
            lines += [f"## every kernel of one warmed-up step (`bench.py --workload {wl}`; serialised by ncu, so overlapped kernels add up)", "",
                      "| kernel | µs | DRAM read MB | DRAM write MB | L2 hit % | L1 hit % |", "|---|---|---|---|---|---|"]
            for r in rows:
                lines.append(f"| `{r['name']}` | {r.get('gpu__time_duration.sum', 0) * 1e6:.1f} | {r.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
                             f"{r.get('dram__bytes_write.sum', 0) / 1e6:.1f} | {r.get('lts__t_sector_hit_rate.pct', 0):.1f} | {r.get('l1tex__t_sector_hit_rate.pct', 0):.1f} |")
            lines += [f"| **step total** | **{tot_t * 1e6:.1f}** | **{tot_rd / 1e6:.1f}** | **{tot_wr / 1e6:.1f}** | | |", "",
                      f"DRAM traffic per step: {(tot_rd + tot_wr) / 1e9:.3f} GB (read {tot_rd / 1e9:.3f} + write {tot_wr / 1e9:.3f}) in {tot_t * 1e3:.3f} ms "
                      f"under ncu = {(tot_rd + tot_wr) / max(tot_t, 1e-12) / 1e9:.0f} GB/s", ""]
            traffic[wl] = {"dram_bytes_per_launch": tot_rd + tot_wr, "kernel": sorted(set(r["name"] for r in rows)),
                           "prof_name": bench.DOMINANT[wl], "round": tag, "commit": commit}
        rp = os.path.join(g, f"ncu_{wl}.raw.csv")
        if os.path.exists(rp) and os.path.getsize(rp) > 0:
            rows = list(csv.reader(open(rp)))
            if len(rows) > 2:
                hdr, units = rows[0], rows[1]
                for r in rows[2:]:
                    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
                    lines += [f"## `--set full`, dominant kernel: `{d.get('Kernel Name', '')[:160]}`  grid {d.get('Grid Size')} block {d.get('Block Size')}", "",
                              "| metric | value | unit |", "|---|---|---|"]
                    lines += [f"| {k} | {d[k]} | {u[k]} |" for k in KEYS if k in d]
                    lines.append("")
        lp = os.path.join(g, f"launches_{wl}.csv")
        if os.path.exists(lp):
            agg = collections.OrderedDict()
            for r in csv.reader(open(lp)):
                if len(r) > 10 and r[0].isdigit():
                    agg.setdefault(r[4].split("(")[0], []).append(float(r[-1].replace(",", "")))
            tot = sum(sum(v) for v in agg.values()) or 1.0
            lines += [f"## launch list of `bench.py --workload {wl} --steps 2 --warmup 3` (ours only; cold-cache, serialised)", "",
                      "| kernel | launches | avg µs | share of our GPU time |", "|---|---|---|---|"]
            lines += [f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |" for k, v in agg.items()]
        open(os.path.join(ROOT, "profiles", f"{tag}_{wl}.md"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: (v["dram_bytes_per_launch"], v.get("round")) for k, v in traffic.items()}, indent=1))


if __name__ == "__main__":
    main()
