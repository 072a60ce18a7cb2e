#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel of an ncu report
(needs -lineinfo and --import-source on):  python tools/ncu_lines.py rep.ncu-rep [kernel-substring] [top]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fn, hdr, seen = None, None, set()
    agg = []
    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            take = pat in (fn or "")
            continue
        if hdr is None or not take or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))

        def num(x):
            try:
                return int(x)
            except (TypeError, ValueError):
                return 0
        agg.append((num(d.get("Instructions Executed")), num(d.get("# Samples")), cur_file.split("/")[-1], int(r[0]), r[1].strip()[:110]))
    tot_i = sum(a[0] for a in agg) or 1
    tot_s = sum(a[1] for a in agg) or 1
    print("# %s  kernel~%r  total warp-instructions %d, samples %d" % (rep, pat, tot_i, tot_s))
    for a in sorted(agg, key=lambda a: -a[1])[:top]:
        print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100.0 * a[1] / tot_s, 100.0 * a[0] / tot_i, a[2], a[3], a[4]))


if __name__ == "__main__":
    main()
