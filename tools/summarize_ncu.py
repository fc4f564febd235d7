"""Summarise ncu outputs brought back in gpurun_out/ into tracked text files under profiles/.

    python tools/summarize_ncu.py <tag> [launches.csv|-] [prof.ncu-rep] [--traffic-json]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launch_csv = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
rep = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

if os.path.exists(launch_csv):
    lines = [l for l in open(launch_csv) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for row in r:
        name = row[ix["Kernel Name"]].split("(")[0][:70]
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {os.path.basename(launch_csv)}\n")
        f.write(f"# total kernel time {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write(f"{'kernel':72s} {'n':>5s} {'total_ms':>10s} {'avg_us':>10s} {'max_us':>10s} {'share':>7s}\n")
        for k, (n, s, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:72s} {n:5d} {s / 1e3:10.3f} {s / n:10.1f} {m:10.1f} {s / tot:7.1%}\n")
    print(open(os.path.join(out_dir, f"{tag}_launches.txt")).read())

if rep and os.path.exists(rep):
    # a .ncu-rep, or its raw page already exported as csv (the report itself may be too large to bring back)
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
    with open(os.path.join(out_dir, f"{tag}_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on: {os.path.basename(rep)} (per captured launch)\n")
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k} [{units[i]}]: " + " | ".join(row[i][:60] for row in data) + "\n")
    print(open(os.path.join(out_dir, f"{tag}_full.txt")).read())
    try:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        tr = [float(row[ir]) * scale[units[ir]] + float(row[iw]) * scale[units[iw]] for row in data]
        if "--traffic-json" in sys.argv:
            json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "source": f"profiles/{tag}_full.txt"},
                      open(os.path.join(out_dir, "predict_var_traffic.json"), "w"))
    except Exception as e:  # noqa: BLE001
        print("traffic:", e)
