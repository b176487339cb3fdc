#!/usr/bin/env python
"""Per-source-region summary of one kernel in an .ncu-rep (needs ncu + nvdisasm and the built .so):
   python tools/ncu_lines.py report.ncu-rep kernel_substring [launch_index] [cubin_name]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "slam-visual-odometry_b200", "lib", "libb200vo.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
addr2line = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", cub], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    if kname not in txt:
        continue
    cur, infn = None, False
    for line in txt.splitlines():
        m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\.text\.(\S+):', line)
        if m:
            infn = kname in m.group(1); continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', line)
        if m and infn:
            addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname, "--launch-skip", str(launch), "--launch-count", "1"],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ia, ie, it, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
ino, iw, il = h.index("stall_no_inst"), h.index("stall_wait"), h.index("stall_long_sb")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    f, l = addr2line.get(a - base) or ('?', 0)
    b = (f, l // 10 * 10)
    v = agg[b]
    v[0] += int(r[ie]); v[1] += int(r[it]); v[2] += int(r[isamp]); v[3] += int(r[ino]); v[4] += int(r[iw]); v[5] += int(r[il])
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print("total warp-instructions %d, samples %d" % (tot, tots))
for k, v in sorted(agg.items(), key=lambda x: -x[1][2])[:int(os.environ.get("TOP", "40"))]:
    print("%-28s winst %5.1f%% lanes %4.1f samples %5.1f%%  no_inst %4.1f%% wait %4.1f%% long_sb %4.1f%%" % (
        "%s:%d" % k, 100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / tots, 100 * v[3] / max(v[2], 1), 100 * v[4] / max(v[2], 1), 100 * v[5] / max(v[2], 1)))
