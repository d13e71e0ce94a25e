#!/usr/bin/env python
"""Instructions and stall samples per PHASE of a kernel: source lines of the first launch in an .ncu-rep are summed
between the '// ----' phase comments of the file named on the command line (default scan_small.cu); other files
(device headers) are listed by file.  Usage: python profiles/ncu_phase_table.py x.ncu-rep [file.cu] [units]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    main_file = sys.argv[2] if len(sys.argv) > 2 else "scan_small.cu"
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname, hdr, seen, rows = None, None, set(), []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            if fname in seen:
                break
            seen.add(fname)
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            try:
                rows.append((fname, int(r[0]), r[1], int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0)))
            except ValueError:
                pass
    tot_i = sum(x[4] for x in rows) or 1
    tot_s = sum(x[3] for x in rows) or 1
    # the full listing (comment lines carry no instructions and are missing above): line -> phase
    src_out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"],
                             capture_output=True, text=True).stdout
    phase_of, cur_file, phase = {}, None, "prologue"
    for r in csv.reader(io.StringIO(src_out)):
        if not r:
            continue
        if r[0] == "File Name":
            cur_file = r[1].split("/")[-1]
            phase = "prologue"
        elif r[0].isdigit() and cur_file == main_file:
            m = re.search(r"// ---- (.{0,60})", r[1])
            if m:
                phase = m.group(1).strip(" -")[:52]
            else:
                m = re.search(r"// (insert: a slot|accumulate relative to|cells are emitted in)", r[1])
                if m:
                    phase = "voxel: " + m.group(1)
            phase_of[int(r[0])] = phase
    acc = {}
    for f, ln, src, s, i in rows:
        key = phase_of.get(ln, "?") if f == main_file else "(" + f + ")"
        a = acc.setdefault(key, [0, 0])
        a[0] += i
        a[1] += s
    print(f"{tot_i / 1e6:.1f} M warp instructions, {tot_s} stall samples")
    for k, (i, s) in sorted(acc.items(), key=lambda x: -x[1][0]):
        extra = f"  {i * 32 / units:6.1f} thread-instr/unit" if units else ""
        print(f"{100 * i / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% smp  {k}{extra}")


if __name__ == "__main__":
    main()
