"""Turn ncu artefacts from gpurun_out/ into tracked summaries under profiles/.

  python tools/summarize_profiles.py launches <in.csv> <out.md> "<title / command>"
  python tools/summarize_profiles.py full <in.ncu-rep> <out.md> "<title>" [kernel-regex]

`launches`: per-kernel table of a `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` CSV.
`full`: the metrics that matter for a roofline reading, for every captured launch of a `--set full` report.
"""
import collections, csv, os, re, subprocess, sys


def launches(src, dst, title):
    rows = list(csv.reader(open(src, errors="replace")))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ik, im, iv, iu = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in data:
        if len(r) > iv:
            name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("kvm::", "")
            agg[name][r[im]].append(float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0))
    tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
    lines = [f"# ncu launch list — {title}", "",
             "Cold-cache, serialised launches under the profiler: compare SHARES, not absolute times.", "",
             "| kernel | launches | avg µs | share of profiled time | avg DRAM read MB | avg DRAM write MB |", "|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        t = v["gpu__time_duration.sum"]
        rd = sum(v.get("dram__bytes_read.sum", [0])) / len(t)
        wr = sum(v.get("dram__bytes_write.sum", [0])) / len(t)
        lines.append(f"| `{k}` | {len(t)} | {sum(t) / len(t):.1f} | {sum(t) / tot:.3f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def full(src, dst, title, pattern=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, rows = rr[0], rr[1], rr[2:]
    ik = hdr.index("Kernel Name")
    lines = [f"# ncu --set full — {title}", ""]
    for r in rows:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
        if pattern and not re.search(pattern, name):
            continue
        m = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in m:
                lines.append(f"| {w} | {m[w][1]} | {m[w][0]} |")
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
