#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py r1 [--batch 2048 --dim 16 --precision 32]

Reads  gpurun_out/<round>_launches.csv     (ncu --metrics gpu__time_duration.sum ... --csv)
       gpurun_out/<round>_full.ncu-rep     (ncu --set full --import-source on)
Writes profiles/<round>_launches.csv       (copy) and profiles/<round>_launches_summary.md (per-kernel share)
       profiles/<round>_full_raw.csv       (ncu -i ... --page raw --csv)
       profiles/<round>_full_summary.md    (duration, DRAM bytes, L2 hit rate, occupancy, top stalls per kernel)
       profiles/<round>_traffic.json       (dram bytes per launch per kernel; bench.py's roofline.traffic)
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def launches(rnd):
    src = os.path.join(GO, f"{rnd}_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PR, f"{rnd}_launches.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
        k = short(r["Kernel Name"])
        tot[k] += us
        cnt[k] += 1
    total = sum(tot.values())
    with open(os.path.join(PR, f"{rnd}_launches_summary.md"), "w") as f:
        f.write(f"# {rnd}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | avg us | share of the step |\n|---|---|---|---|\n")
        for k in sorted(tot, key=lambda x: -tot[x]):
            f.write(f"| {k} | {cnt[k]} | {tot[k] / cnt[k]:.2f} | {100 * tot[k] / total:.1f} % |\n")
    print(open(os.path.join(PR, f"{rnd}_launches_summary.md")).read())


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
])


def full(rnd, name, meta):
    rep = os.path.join(GO, f"{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(PR, f"{name}_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    out = [f"# {name}: ncu --set full --clock-control none ({meta})\n"]
    for r in rows[2:]:
        k = short(r[hdr.index("Kernel Name")])
        out.append(f"\n## {k}  ({r[hdr.index('Kernel Name')].strip()})\n")
        out.append("| metric | value |\n|---|---|")
        rd = wr = 0.0
        for m, label in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {label} (`{m}`) | {r[i]} {units[i]} |")
                if m == "dram__bytes_read.sum":
                    rd = to_bytes(r[i], units[i])
                if m == "dram__bytes_write.sum":
                    wr = to_bytes(r[i], units[i])
        traffic[k] = rd + wr
        out.append(f"| **DRAM traffic per launch** | {(rd + wr) / 1e6:.3f} MB |")
        # stall reasons from the source page
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", k], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        h = next((x for x in srows if "Source" in x and "# Samples" in x), None)
        if h:
            st = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            agg = defaultdict(int)
            for x in srows[srows.index(h) + 1:]:
                for i, c in st:
                    try:
                        agg[c] += int(x[i])
                    except Exception:
                        pass
            tot = sum(agg.values()) or 1
            top = sorted(agg.items(), key=lambda kv: -kv[1])[:5]
            out.append("| top warp stall reasons (samples) | " + ", ".join(f"{c[6:]} {100 * v / tot:.0f} %" for c, v in top if v) + " |")
    open(os.path.join(PR, f"{name}_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    return traffic


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
    args = dict(batch=2048, dim=16, precision=32)
    for i, a in enumerate(sys.argv):
        if a.startswith("--") and a[2:] in args:
            args[a[2:]] = int(sys.argv[i + 1])
    os.makedirs(PR, exist_ok=True)
    launches(rnd)
    tr = full(rnd, f"{rnd}_full", f"python bench.py --steps 20 --warmup 3, one step of the timed region; batch {args['batch']}, dim {args['dim']}, fp{args['precision']}")
    if tr:
        json.dump(dict(args, dram_bytes_per_launch=tr, source=f"profiles/{rnd}_full_raw.csv"), open(os.path.join(PR, f"{rnd}_traffic.json"), "w"), indent=1)
    for extra in sorted(os.listdir(GO)):
        if extra.startswith(rnd + "_") and extra.endswith(".ncu-rep") and extra != f"{rnd}_full.ncu-rep":
            full(rnd, extra[:-8], "see profiles/README.md for the command")


if __name__ == "__main__":
    main()
