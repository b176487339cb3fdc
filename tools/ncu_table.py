#!/usr/bin/env python
"""Compact per-launch table from an .ncu-rep (needs ncu on PATH):  python tools/ncu_table.py file.ncu-rep [regex]"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"), ("smsp__inst_executed.sum", "winst"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("sass__inst_executed_local_loads", "lld"), ("sass__inst_executed_local_stores", "lst"),
        ("sm__inst_executed_pipe_fp64.sum", "fp64"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_alu.sum", "alu"), ("sm__inst_executed_pipe_lsu.sum", "lsu")]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected", "selected", "barrier",
          "lg_throttle", "mio_throttle", "branch_resolving", "no_instructions", "dispatch_stall", "imc_miss", "drain", "membar", "sleeping", "tex_throttle"]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        if pat and not pat.search(name):
            continue
        parts = []
        for c, label in COLS:
            if c in ix:
                v = r[ix[c]]
                u = units[ix[c]]
                try:
                    f = float(v.replace(",", ""))
                    if label == "us":                      # ncu picks the time unit per report
                        f *= {"s": 1e6, "second": 1e6, "ms": 1e3, "msecond": 1e3, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3}.get(u, 1.0)
                    v = ("%.3g" % f)
                except ValueError:
                    pass
                parts.append("%s=%s%s" % (label, v, (u if u in ("Mbyte", "Kbyte", "Gbyte", "byte") else "")))
        st = []
        tot = 0.0
        for s in STALLS:
            k = "smsp__pcsamp_warps_issue_stalled_" + s
            if k in ix:
                try:
                    f = float(r[ix[k]].replace(",", ""))
                except ValueError:
                    continue
                st.append((f, s)); tot += f
        st.sort(reverse=True)
        print(name, " ".join(parts))
        print("    stalls: " + ", ".join("%s %.0f%%" % (s, 100 * f / tot) for f, s in st[:6] if tot > 0))


if __name__ == "__main__":
    main()
