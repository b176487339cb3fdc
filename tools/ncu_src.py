#!/usr/bin/env python
"""Per-source-line table of one kernel launch from an .ncu-rep captured with --import-source on:
   python tools/ncu_src.py report.ncu-rep kernel_regex [launch_skip] [top] [samples]
("samples" as the fifth argument sorts by stall samples instead of executed instructions: latency-bound kernels)
columns: line, warp-instructions executed (share), avg active threads, stall samples (share), source text."""
import csv, io, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     stdout=subprocess.PIPE, text=True).stdout
files, cur, hdr = {}, None, None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] in ("File Path", "File Name"):
        cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = {h: i for i, h in enumerate(row)}; continue
    if hdr is None or not row[0].strip().isdigit():
        continue
    def num(name):
        try:
            return float(row[hdr[name]].replace(",", ""))
        except (ValueError, KeyError, IndexError):
            return 0.0
    ie, te, sm = num("Instructions Executed"), num("Thread Instructions Executed"), num("# Samples")
    files.setdefault(cur, []).append((int(row[0]), ie, te, sm, row[1].strip()))
tot_i = sum(r[1] for v in files.values() for r in v)
tot_s = sum(r[3] for v in files.values() for r in v)
print("total warp-instructions %.4g, samples %d" % (tot_i, tot_s))
rows = [(f,) + r for f, v in files.items() for r in v]
rows.sort(key=lambda r: -(r[4] if len(sys.argv) > 5 and sys.argv[5] == 'samples' else r[2]))
for f, ln, ie, te, sm, src in rows[:top]:
    print("%-18s %5d  winst %5.1f%%  lanes %4.1f  samples %5.1f%%  %s" % (f, ln, 100 * ie / max(tot_i, 1), te / max(ie, 1), 100 * sm / max(tot_s, 1), src[:110]))
