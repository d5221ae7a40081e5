"""Per-kernel summary of an `ncu --set full` report: python tools/ncu_summary.py REPORT.ncu-rep [TRAFFIC.json] [NOTE]
Prints duration, registers, occupancy, issue activity, DRAM bytes and the top stall reasons of every profiled launch;
with a second argument also writes the per-launch DRAM traffic bench.py reports as roofline.traffic."""
import csv, json, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1]
    vals = {k: r[col[k]] for k in keep if k in col}
    def num(k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return 0.0
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    unit_r, unit_w = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(unit_r, 1.0); wr *= scale.get(unit_w, 1.0)
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
           h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
            try:
                stalls.append((round(float(r[col[h]]), 2), h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")))
            except Exception:
                pass
    stalls.sort(reverse=True)
    print(name, vals)
    print("   dram read %.1f MB write %.1f MB; stalls per issue: %s" % (rd / 1e6, wr / 1e6, stalls[:6]))
    traffic[name] = {"traffic_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "duration_us": num("gpu__time_duration.sum") / (1e3 if rows[1][col["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0),
                     "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "registers": int(num("launch__registers_per_thread"))}
if len(sys.argv) > 2:
    traffic["source"] = sys.argv[3] if len(sys.argv) > 3 else rep
    json.dump(traffic, open(sys.argv[2], "w"), indent=1)
