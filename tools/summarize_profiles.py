"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <round-tag> <launches.csv> <walker.ncu-rep>"""
import collections, csv, json, os, subprocess, sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

rows = list(csv.reader(open(launches)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[h], rows[h + 1:]
ik, im, iv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in data:
    if len(r) > iv:
        agg[r[ik].split("(")[0].replace("void ", "")][r[im]].append(float(r[iv].replace(",", "")))
tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
lines = [f"# ncu launch list — {tag}", "",
         "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
         "--csv python bench.py --steps 10 --warmup 3 --cpu-queries 0 --no-query-set` (cold-cache, serialised: compare SHARES).", "",
         "| kernel | launches | avg µs | share of step | avg DRAM read MB | avg DRAM write MB |", "|---|---|---|---|---|---|"]
walker_traffic = None
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = v["gpu__time_duration.sum"]
    rd, wr = sum(v["dram__bytes_read.sum"]) / len(t), sum(v["dram__bytes_write.sum"]) / len(t)
    if "cnsm_relay" in k:
        walker_traffic = rd + wr
    lines.append(f"| {k} | {len(t)} | {sum(t) / len(t) / 1e3:.1f} | {sum(t) / tot:.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} |")
open(os.path.join(out_dir, f"launches_{tag}.md"), "w").write("\n".join(lines) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, vals = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
m = {hh: (u, v) for hh, u, v in zip(hdr, units, vals)}
wl = [f"# ncu --set full: cnsm_relay_kernel — {tag}", "",
      "One launch of the statistics walker inside `bench.py` (n=1e8, m=1024, chunk 6144, 16276 chains, 509 CTAs x 192 threads).", "",
      "| metric | value | unit |", "|---|---|---|"]
for w in want:
    if w in m:
        wl.append(f"| {w} | {m[w][1]} | {m[w][0]} |")
open(os.path.join(out_dir, f"walker_{tag}.md"), "w").write("\n".join(wl) + "\n")
json.dump({"round": tag, "n_per_gpu": 100_000_000, "chain_chunk": 6144,
           "walker_dram_bytes_per_launch": walker_traffic,
           "source": f"profiles/launches_{tag}.md (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over launches)"},
          open(os.path.join(out_dir, f"roofline_{tag}.json"), "w"), indent=1)
print("\n".join(lines[-6:]))
print("\n".join(wl[-20:]))
