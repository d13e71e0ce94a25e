#!/usr/bin/env python
"""Per-source-line hot spots of an .ncu-rep captured with --import-source on (kernels built with -lineinfo):
stall samples and executed warp instructions aggregated by CUDA source line.
Usage: python profiles/ncu_hotspots.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    kernel, fname, hdr = None, None, None
    rows = {}  # kernel -> list of (file, line, source, samples, inst)
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kernel = r[1].split("(")[0].split("::")[-1] + ("<" + r[1].split("<(int)")[1].split(">")[0] + ">" if "<(int)" in r[1] else "")
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            try:
                samples = int(r[hdr.index("# Samples")] or 0)
                inst = int(r[hdr.index("Instructions Executed")] or 0)
            except ValueError:
                continue
            rows.setdefault(kernel, []).append((fname, int(r[0]), r[1].strip(), samples, inst))
    for k, lst in rows.items():
        tot_s = sum(x[3] for x in lst) or 1
        tot_i = sum(x[4] for x in lst) or 1
        print(f"== {k}: {tot_i / 1e6:.1f} M warp instructions, {tot_s} stall samples")
        print("   by stall samples:")
        for f, ln, src, s, i in sorted(lst, key=lambda x: -x[3])[:top]:
            print(f"   {100 * s / tot_s:5.1f}% smp {100 * i / tot_i:5.1f}% inst  {f}:{ln}  {src[:90]}")
        print("   by instructions:")
        for f, ln, src, s, i in sorted(lst, key=lambda x: -x[4])[:top]:
            print(f"   {100 * i / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% smp  {f}:{ln}  {src[:90]}")


if __name__ == "__main__":
    main()
