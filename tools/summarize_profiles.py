#!/usr/bin/env python3
"""Turn the gpurun_out ncu captures into the tracked text summaries under profiles/.

   python tools/summarize_profiles.py <round-tag> <launches.csv> <bench.json> <capture.ncu-rep>[=label] ...

Writes profiles/<tag>_launch_list.txt (per-launch times of the bench steps captured with
`ncu --metrics gpu__time_duration.sum`, next to bench.py's event-timed shares), profiles/<tag>_ncu_full_summary.txt
(key metrics + top stalls of every kernel in the `--set full` captures) and profiles/<tag>_ncu_traffic.json
(DRAM bytes per launch of the kernels bench.py reports a roofline for; bench.py reads it back as `roofline.traffic`)."""
import collections, csv, json, os, re, subprocess, sys

tag, launches, bench = sys.argv[1:4]
reps = sys.argv[4:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("kvg::", "")
    return re.sub(r"\(.*", "", name)


# ---- launch list: per-kernel share of the step (cold-cache, serialised: compare SHARES)
rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
seq = []
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = short(r[ix["Kernel Name"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    agg.setdefault(name, []).append(v)
    seq.append((name, v))
tot = sum(sum(v) for k, v in agg.items() if k != "k_fill")
d = json.loads(open(bench).read().strip().splitlines()[-1])
with open(os.path.join(out, "%s_launch_list.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-extra ...\n")
    f.write("# %d launches captured; per-launch times are cold-cache and serialised (no overlap between the parse stream and the\n"
            "# scan stream, no programmatic dependent launch) -> compare SHARES with bench.py's kernel_ms_per_step; k_fill is the\n"
            "# untimed L2 flush between steps\n" % len(seq))
    f.write("%-44s %8s %12s %10s %8s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-44s %8d %12.1f %10.2f %7.1f%%\n" % (k[:44], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3,
                                                       100 * sum(v) / tot if k != "k_fill" else 0.0))
    # the last complete step, launch by launch
    fills = [i for i, s_ in enumerate(seq) if s_[0] == "k_fill"]
    # the last COMPLETE step: the launches between the last two L2 flushes (or behind the last one)
    if len(fills) >= 2 and len(seq) - 1 - fills[-1] < fills[-1] - fills[-2] - 1:
        step = seq[fills[-2] + 1:fills[-1]]
    else:
        step = seq[fills[-1] + 1:] if fills else seq
    f.write("\n# one captured step, in launch order (us)\n")
    for name, v in step:
        f.write("   %-44s %8.2f\n" % (name[:44], v / 1e3))
    f.write("   %-44s %8.2f   (sum of serialised launches; the step itself: %.1f us with overlap)\n"
            % ("total", sum(v for _, v in step) / 1e3, d["ms_per_step"] * 1e3))
    ks = d["kernel_ms_per_step"]; st = sum(ks.values())
    f.write("\n# bench.py (CUDA events around every launch, warm, L2 flushed between steps) for comparison; ~5 us of event\n"
            "# overhead per launch is inside these figures\n")
    for k, v in sorted(ks.items(), key=lambda kv: -kv[1]):
        f.write("   %-28s %10.2f us/step %7.1f%%\n" % (k, v * 1e3, 100 * v / st))

# ---- full captures: key metrics per kernel
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per LAUNCH from the committed ncu --set full captures "
                       "(profiles/%s_ncu_full_summary.txt); bench.py copies these into roofline.traffic when the workload "
                       "matches.  Lines still in L2 when a kernel ends are not in these figures." % tag}
with open(os.path.join(out, "%s_ncu_full_summary.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on (one capture per section); times under the profiler are\n"
            "# cold-cache and serialised: not bench numbers\n")
    for spec in reps:
        rep, _, label = spec.partition("=")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        if len(rr) < 3:
            f.write("\n#### %s: empty capture\n" % os.path.basename(rep))
            continue
        h, units = rr[0], rr[1]; jx = {x: i for i, x in enumerate(h)}
        stall = [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
        f.write("\n#### %s  %s\n" % (os.path.basename(rep), label))
        seen = collections.Counter()
        for r in rr[2:]:
            name = short(r[jx["Kernel Name"]])
            seen[name] += 1
            f.write("\n== %s  [launch #%d of this kernel in the capture]\n" % (r[jx["Kernel Name"]][:120], seen[name]))
            for w in want:
                if w in jx:
                    f.write("   %-62s %s %s\n" % (w, r[jx[w]], units[jx[w]]))
            vals = sorted(((float(r[jx[x]].replace(",", "") or 0), x.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                           for x in stall), reverse=True)[:5]
            f.write("   top stalls (warps per issue-active): " + ", ".join("%s %.2f" % (x, v) for v, x in vals) + "\n")
            try:
                t = float(r[jx["gpu__time_duration.sum"]].replace(",", "")) * TUNIT[units[jx["gpu__time_duration.sum"]]]
                rd = float(r[jx["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[jx["dram__bytes_read.sum"]]]
                wr = float(r[jx["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[jx["dram__bytes_write.sum"]]]
                f.write("   DRAM traffic %.2f MB in %.1f us = %.0f GB/s (under the profiler)\n" % ((rd + wr) / 1e6, t * 1e6, (rd + wr) / t / 1e9))
                key = None
                if label.startswith("parse256") and name == "k_pciids_scan":
                    key = "pciids_parse@256"
                elif label.startswith("big") and name.startswith("k_classify_ragged") and seen[name] == 1:
                    key = "classify_compact@16777216"
                elif label.startswith("cfg2"):
                    fam = {"k_order_scatter": "order_scatter", "k_order_hist": "order_hist", "k_order_final": "order_final",
                           "k_classify_oneshot": "classify_compact", "k_order_tilescan": "order_tilescan"}
                    for pfx, famname in fam.items():
                        if name.startswith(pfx) and seen[name] == 1:
                            key = famname + "@config2"
                if key and key not in traffic:
                    traffic[key] = {"bytes_per_launch": int(rd + wr), "launch": r[jx["Kernel Name"]][:100],
                                    "us_under_ncu": round(t * 1e6, 2), "capture": os.path.basename(rep)}
            except Exception as e:  # noqa: BLE001
                f.write("   (traffic summary unavailable: %s)\n" % e)
json.dump(traffic, open(os.path.join(out, "%s_ncu_traffic.json" % tag), "w"), indent=1)
print("wrote", sorted(x for x in os.listdir(out) if x.startswith(tag)))
