#!/usr/bin/env python3
"""Turn the gpurun_out ncu captures into the tracked text summaries under profiles/.
   python tools/summarize_profiles.py <round-tag> <launches.csv> <prof.ncu-rep> <bench.json>"""
import collections, csv, json, os, subprocess, sys
tag, launches, rep, bench = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list: per-kernel share of the step (cold-cache, serialised: compare SHARES)
rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    agg.setdefault(name, []).append(float(r[ix["Metric Value"]]))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, "%s_launch_list.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (python bench.py --steps 2 --warmup 3 ...)\n")
    f.write("# %d launches captured; per-launch times are cold-cache and serialised -> compare SHARES with bench.py's kernel_ms_per_step\n" % (len(rows) - 1))
    f.write("%-28s %8s %12s %12s %8s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-28s %8d %12.1f %12.2f %7.1f%%\n" % (k, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    d = json.load(open(bench))
    ks = d["kernel_ms_per_step"]; st = sum(ks.values())
    f.write("\n# bench.py (CUDA events, warm, L2 flushed between steps) shares for comparison\n")
    for k, v in sorted(ks.items(), key=lambda kv: -kv[1]):
        f.write("%-28s %31.2f us/step %7.1f%%\n" % (k, v * 1e3, 100 * v / st))

# ---- full capture: key metrics per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]; ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
with open(os.path.join(out, "%s_ncu_full_summary.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on ; python tools/profile_kernels.py 16777216 32 1\n")
    f.write("# (16,777,216 PCI records / 32 pci.ids images: inputs larger than L2)\n")
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        seen[name.split("(")[0]] += 1
        f.write("\n== %s  [launch #%d of this kernel]\n" % (name[:110], seen[name.split("(")[0]]))
        for w in want:
            if w in ix:
                f.write("   %-62s %s %s\n" % (w, r[ix[w]], units[ix[w]]))
        vals = sorted(((float(r[ix[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:5]
        f.write("   top stalls (warps per issue-active): " + ", ".join("%s %.2f" % (h, v) for v, h in vals) + "\n")
        try:
            t = float(r[ix["gpu__time_duration.sum"]]); rd = float(r[ix["dram__bytes_read.sum"]]); wr = float(r[ix["dram__bytes_write.sum"]])
            u = units[ix["dram__bytes_read.sum"]]; mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tu = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[units[ix["gpu__time_duration.sum"]]]
            f.write("   DRAM traffic %.1f MB in %.1f us = %.0f GB/s (under the profiler; not a bench number)\n" % ((rd + wr) * mult / 1e6, t * tu * 1e6, (rd + wr) * mult / (t * tu) / 1e9))
        except Exception as e:
            f.write("   (traffic summary unavailable: %s)\n" % e)
print("wrote", os.listdir(out))
