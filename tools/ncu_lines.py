#!/usr/bin/env python
"""Per-source-line executed instructions / stall samples of one kernel of an .ncu-rep
(usage: ncu_lines.py report.ncu-rep kernel_regex [top])."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out, tot, files, cur = [], 0, set(), None
for r in rows:
    if r and r[0] == "File Path":
        if r[1] in files:
            break   # the same file again: the next launch - first launch only
        files.add(r[1]); cur = r[1].rsplit("/", 1)[-1]
    if len(r) >= 8 and r[0].isdigit():
        try:
            n, s = int(r[7]), int(r[4])
        except ValueError:
            continue
        out.append((n, s, int(r[0]), (cur + ": " if cur and not cur.startswith("orb_extract") else "") + r[1]))
        tot += n
stot = sum(o[1] for o in out) or 1
print("total warp instructions", tot)
for n, s, l, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  L%-4d %s" % (100.0 * n / tot, 100.0 * s / stot, l, src.strip()[:100]))
