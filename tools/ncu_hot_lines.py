#!/usr/bin/env python
"""Aggregates the source page of an .ncu-rep (captured with --import-source on, code built with -lineinfo) by source line:
share of the warp samples and executed instructions per line, top N. usage: python tools/ncu_hot_lines.py in.ncu-rep out.txt [N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    raw = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    if not raw.strip():
        raw = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if any("Sampl" in c for c in r))
    hdr = rows[hdr_i]

    def col(*names):
        for n in names:
            for i, c in enumerate(hdr):
                if c.strip() == n:
                    return i
        for n in names:
            for i, c in enumerate(hdr):
                if n in c:
                    return i
        return None

    c_src, c_smp, c_ins = col("Source"), col("# Samples", "Warp Stall Sampling (All Samples)", "Samples"), col("Instructions Executed", "Executed")
    c_file = col("File Path", "File") if col("File Path", "File") is not None else None
    c_line = col("Line") if col("Line") is not None else None
    agg = defaultdict(lambda: [0.0, 0.0])
    for r in rows[hdr_i + 1:]:
        if len(r) <= max(c_src, c_smp):
            continue
        key = r[c_src].strip()
        if c_file is not None and c_line is not None and len(r) > max(c_file, c_line):
            key = f"{r[c_file].split('/')[-1]}:{r[c_line]}  {key}"
        try:
            agg[key][0] += float(r[c_smp] or 0)
            if c_ins is not None:
                agg[key][1] += float(r[c_ins] or 0)
        except ValueError:
            continue
    total = sum(v[0] for v in agg.values()) or 1.0
    with open(dst, "w") as f:
        f.write(f"total samples {int(total)}  (columns: {hdr[c_smp]!r}, {hdr[c_ins] if c_ins is not None else None!r})\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            f.write(f"{100 * v[0] / total:5.1f}% ins={int(v[1]):>11d} {k[:150]}\n")
    print(f"wrote {dst}: {len(agg)} lines, {int(total)} samples")


if __name__ == "__main__":
    main()
