#!/usr/bin/env python3
"""Stall samples and executed instructions per CUDA source line of one kernel, from an ncu report taken with
`--set full --import-source on` of a `-lineinfo` build.

  python tools/ncu_by_line.py gpurun_out/r1z.ncu-rep vc_carve_bricks [top_n]

Prints the top_n lines by warp-stall samples in source order (share of the kernel's samples / of its instructions)."""
import csv
import io
import subprocess
import sys


def main(report, kernel, top_n=40):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{kernel}", "--launch-count", "1"], capture_output=True, text=True).stdout
    agg, cur = {}, "?"
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        elif len(r) > 8 and r[0].isdigit():  # a CUDA line (its SASS rows follow with an empty first column)
            try:
                samples, inst = int(r[6]), int(r[7])
            except ValueError:
                continue
            a = agg.setdefault((cur, int(r[0])), [r[1].strip()[:120], 0, 0])
            a[1] += samples
            a[2] += inst
    ts, ti = sum(a[1] for a in agg.values()) or 1, sum(a[2] for a in agg.values()) or 1
    print(f"{kernel}: {ts} samples, {ti} warp instructions")
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]
    for (f, ln), (src, s, i) in sorted(top):
        print(f"{f}:{ln:<5d} samples {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%  {src}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
