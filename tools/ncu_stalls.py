#!/usr/bin/env python
"""Print the top warp-stall reasons (pc sampling) per kernel of an .ncu-rep."""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]; kn = hdr.index("Kernel Name")
for r in rows[2:]:
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try: st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1
    print("==", r[kn][:50], " ".join("%s=%.0f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:7]))
