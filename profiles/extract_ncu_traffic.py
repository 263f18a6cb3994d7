#!/usr/bin/env python
"""Per-kernel DRAM traffic and pipe utilisation from `ncu --set full` reports -> profiles/r02_ncu_traffic.json (read by bench.py's
roofline.traffic) and a markdown summary.
usage: python profiles/extract_ncu_traffic.py category=report.ncu-rep[@launch] | category=raw.csv:kernel_regex[@nth] ...
A report is exported with `ncu -i report --page raw --csv`; a raw CSV (exported on the GPU box, the reports are too large to
travel) is read directly and the nth launch whose kernel name matches the regex is used (default: the median one by duration)."""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "lts__t_bytes.sum": "l2_bytes",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def read(report, launch=0, regex=None):
    import re
    if report.endswith(".csv"):
        rows = list(csv.reader(open(report)))
    else:
        txt = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    if regex:
        ik, it = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
        cand = [r for r in rows[2:] if len(r) > ik and re.search(regex, r[ik])]
        if not cand:
            raise SystemExit("no launch matches %r in %s" % (regex, report))
        cand.sort(key=lambda r: float(r[it].replace(",", "")))
        vals = cand[launch] if launch is not None else cand[len(cand) // 2]
        n_match = len(cand)
    else:
        vals, n_match = rows[2 + (launch or 0)], 1
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    stalls = {}
    for i, h in enumerate(hdr):
        try:
            v = float(vals[i].replace(",", ""))
        except ValueError:
            continue
        if h in WANT:
            out[WANT[h]] = v * SCALE.get(units[i], 1.0)
        elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and v >= 0.3:
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 2)
    out["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
    out["dram_bytes_per_launch"] = out.get("dram_read", 0.0) + out.get("dram_write", 0.0)
    out["launches_matched"] = n_match
    return out


def main():
    res = {}
    for arg in sys.argv[1:]:
        cat, rep = arg.split("=", 1)
        launch, regex = None, None
        if "@" in rep:
            rep, launch = rep.rsplit("@", 1)
            launch = int(launch)
        if ".csv:" in rep:
            rep, regex = rep.split(".csv:", 1)
            rep += ".csv"
        elif launch is None:
            launch = 0
        r = read(rep, launch, regex)
        r["report"] = os.path.basename(rep)
        res[cat] = r
    json.dump(res, open(os.path.join(HERE, "r02_ncu_traffic.json"), "w"), indent=1)
    print("| category | kernel | time | DRAM read+write | warp instr | issue active | alu / fma pipe | warps active | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|")
    for cat, r in res.items():
        print("| %s | `%s` | %.3f ms | %.1f MB | %.1f M | %.1f %% | %.1f / %.1f %% | %.1f %% | %s |" % (
            cat, r["kernel"].split("(")[0][:60], r.get("duration", 0) * 1e3, r["dram_bytes_per_launch"] / 1e6, r.get("warp_instructions", 0) / 1e6, r.get("issue_active_pct", 0),
            r.get("pipe_alu_pct", 0), r.get("pipe_fma_pct", 0), r.get("warps_active_pct", 0), ", ".join("%s %.2f" % kv for kv in r["top_stalls"].items())))


if __name__ == "__main__":
    main()
