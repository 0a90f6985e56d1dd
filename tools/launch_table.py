"""Per-launch table from an ncu --csv launch list (duration, both tensor-pipe metrics, issue slots, DRAM bytes) and,
with --traffic-json, the measured DRAM traffic of the tcgen05 kernels that bench.py reports as roofline.traffic.

  python tools/launch_table.py gpurun_out/r2_launches_mb4096.csv [--md] [--traffic-json profiles/r2_traffic.json --mb 4096]
"""
import argparse
import csv
import json
import os
from collections import OrderedDict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--md", action="store_true", help="markdown table")
ap.add_argument("--traffic-json", default=None)
ap.add_argument("--mb", type=int, default=4096)
ap.add_argument("--command", default="")
args = ap.parse_args()

rows = [l for l in open(args.csv) if l.startswith('"')]
launches = OrderedDict()
for r in csv.DictReader(rows):
    d = launches.setdefault(r["ID"], {"kernel": r["Kernel Name"].replace("void ", ""), "grid": r["Grid Size"]})
    try:
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        pass

cols = [("gpu__time_duration.sum", "us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%", 1.0),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tc%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%", 1.0),
        ("dram__bytes_read.sum", "rd MB", 1e-6), ("dram__bytes_write.sum", "wr MB", 1e-6)]
present = [c for c in cols if any(c[0] in d for d in launches.values())]
sep = " | " if args.md else "  "
hdr = ["kernel", "grid"] + [c[1] for c in present]
if args.md:
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
tot_us, tot_bytes, conv_bytes, conv_n = 0.0, 0.0, 0.0, 0
for d in launches.values():
    vals = ["%.1f" % (d.get(c[0], float("nan")) * c[2]) for c in present]
    name = d["kernel"][:40]
    line = sep.join(["%-40s" % name, "%-14s" % d["grid"]] + ["%9s" % v for v in vals])
    print(("| " + line + " |") if args.md else line)
    tot_us += d.get("gpu__time_duration.sum", 0.0) * 1e-3
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot_bytes += b
    if any(k in d["kernel"] for k in ("conv3x3_tc_kernel", "conv3x3_rows_kernel", "l01_fused_kernel")):
        conv_bytes += b
        conv_n += 1
print("total %.1f us, DRAM %.1f MB; tcgen05 kernels: %d launches, %.1f MB" % (tot_us, tot_bytes * 1e-6, conv_n, conv_bytes * 1e-6))
if args.traffic_json and conv_n:
    out = json.load(open(args.traffic_json)) if os.path.exists(args.traffic_json) else {}
    out["command"] = args.command or out.get("command", "")
    out[str(args.mb)] = {"bytes_per_conv_launch": conv_bytes / conv_n, "conv_launches_per_chunk_pair": conv_n,
                         "dram_bytes_per_chunk_pair_all_kernels": tot_bytes, "dram_bytes_per_pair": tot_bytes / args.mb,
                         "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, " + os.path.basename(args.csv)}
    json.dump(out, open(args.traffic_json, "w"), indent=1)
    print("wrote", args.traffic_json, "DRAM bytes per pair %.0f" % (tot_bytes / args.mb))
