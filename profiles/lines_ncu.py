#!/usr/bin/env python
"""Per-source-line stall table of an ``ncu --set full --import-source on`` capture (needs -lineinfo builds).

Usage: python profiles/lines_ncu.py report.ncu-rep [top_n]   -> lines ranked by warp-stall samples, with the share of
long_scoreboard (waiting on a load) among them and the warp instructions executed."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg = {}
    fn = path = ""
    hdr = None
    total = 0
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            fn = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue                       # SASS rows have an empty line number: the line row carries their sum
        samples = int(r[hdr["# Samples"]] or 0)
        lsb = int(r[hdr["stall_long_sb"]] or 0)
        ex = int(r[hdr["Instructions Executed"]] or 0)
        key = (path, int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += samples; a[1] += lsb; a[2] += ex
        total += samples
    print(f"# {rep}: {total} stall samples attributed to source lines")
    print(f"{'samples':>8s} {'share':>6s} {'long_sb':>7s} {'exec(M)':>8s}  file:line  source")
    for (path, line, text), (s, l, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{s:8d} {100.0 * s / total:5.1f}% {100.0 * l / max(1, s):6.0f}% {ex / 1e6:8.1f}  {path}:{line}  {text}")


if __name__ == "__main__":
    main()
