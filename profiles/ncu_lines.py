#!/usr/bin/env python
"""Instructions and stall samples per CUDA source line for the FIRST launch in an .ncu-rep (captured with
--import-source on, kernels built with -lineinfo).  Usage: python profiles/ncu_lines.py x.ncu-rep [top_n] [file-substr]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    only = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname, hdr, seen_files, rows = None, None, set(), []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            if fname in seen_files:  # second launch starts
                break
            seen_files.add(fname)
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            try:
                rows.append((fname, int(r[0]), r[1].strip(), int(r[hdr.index("# Samples")] or 0),
                             int(r[hdr.index("Instructions Executed")] or 0)))
            except ValueError:
                pass
    tot_i = sum(x[4] for x in rows) or 1
    tot_s = sum(x[3] for x in rows) or 1
    print(f"{tot_i / 1e6:.1f} M warp instructions, {tot_s} stall samples")
    cum = 0.0
    for f, ln, src, s, i in sorted(rows, key=lambda x: -x[4])[:top]:
        if only and only not in f:
            continue
        cum += 100 * i / tot_i
        print(f"{100 * i / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% smp (cum {cum:5.1f})  {f}:{ln}  {src[:100]}")


if __name__ == "__main__":
    main()
