#!/usr/bin/env python
"""Per-source-line warp-stall samples of each kernel in an ncu report (needs -lineinfo and --import-source on):

    python tools/ncu_hotspots.py gpurun_out/r1_full.ncu-rep profiles/r1_full_hotspots.md [top_n]

Runs `ncu -i <rep> --page source --print-source cuda,sass --csv` and lists the lines with the most stall
samples (where the warps of that kernel wait), with their share of the kernel's samples."""
import csv
import io
import os
import re
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    # blocks: "File Path",<file> / "Function Name",<kernel> / header / rows; rows with a line number are CUDA source
    # lines carrying the sum over their SASS instructions (which follow with an empty line number)
    per_kernel = {}
    fpath = func = None
    header = None
    for row in csv.reader(io.StringIO(txt)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath, header = row[1], None
        elif row[0] == "Function Name":
            func = row[1]
        elif row[0] == "Line No":
            header = row
            col = next(i for i, c in enumerate(row) if c.startswith("Warp Stall Sampling (All"))
        elif header is not None and row[0] != "" and func is not None:
            try:
                n = int(float(row[col].replace(",", "")))
            except (ValueError, IndexError):
                continue
            per_kernel.setdefault(func, []).append((n, os.path.basename(fpath or "?"), row[0], row[1].strip()))
    md = ["# per-source-line warp-stall samples (ncu --set full --import-source on; `--page source --print-source cuda,sass`)\n",
          f"Report: `{rep}` (one step of `python bench.py --steps 20 --warmup 3`, B = 2048, dim 16, fp32; kernels serialised by "
          "ncu). `samples` = warp-stall sampling hits attributed to the CUDA source line (all samples, summed over its SASS "
          "instructions and inlined call sites); the share is of the kernel's total. Written by `tools/ncu_hotspots.py`.\n"]
    for func, agg in per_kernel.items():
        total = sum(n for n, *_ in agg) or 1
        agg.sort(key=lambda x: -x[0])
        short = re.search(r"(k_[a-z0-9_]+)", func)
        md.append(f"\n## {short.group(1) if short else func}  ({total} samples)\n")
        md.append("| samples | share | file:line | source |")
        md.append("|---|---|---|---|")
        for n, f, ln, src in agg[:top_n]:
            if n == 0:
                break
            md.append(f"| {n} | {100.0 * n / total:.0f} % | `{f}:{ln}` | `{src[:140].replace('|', '/')}` |")
    open(out, "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
