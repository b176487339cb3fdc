#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches_r01.csv > profiles/launches_r01_summary.md
Times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).
Kernels of the synthetic-frame renderer (torch `at::` kernels, outside the timed region) are
listed separately from the library's own kernels (namespace b200vo).
"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += v
    # ncu prints the namespace only in some modes (not with -k filters): recognise the library's kernels by name too
    lib = re.compile(r"b200vo::|^(resize_tma|resize_strip|resize|repack_gray|fill32|blur_strip|bgr2gray|fast|fast_levels|select|orient|keypoint_finalize|describe|blur|brief|match|match_knn2|filter|"
                     r"five_point_group|five_point_deferred|score|score_models|finalize|pose_decompose|pose_vote|pose_select|pose_mask|"
                     r"clear_detect_state|adaptive_update|stamp_rounds|popc_rate)_kernel")
    ours = {k: v for k, v in agg.items() if lib.search(k)}
    other = {k: v for k, v in agg.items() if not lib.search(k)}
    tot = sum(v[1] for v in ours.values())
    print("# ncu launch list summary: %s\n" % path)
    print("Library kernels (namespace b200vo): %d launches, %.1f us total under ncu\n" % (sum(v[0] for v in ours.values()), tot))
    print("| kernel | launches | total us | share of library time |\n|---|---|---|---|")
    for k, (c, t) in sorted(ours.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, c, t, 100 * t / tot))
    print("\nOther kernels in the capture (torch ops of the synthetic renderer / tensor plumbing, outside the timed region): "
          "%d launches, %.1f us" % (sum(v[0] for v in other.values()), sum(v[1] for v in other.values())))


if __name__ == "__main__":
    main(sys.argv[1])
