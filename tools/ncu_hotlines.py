"""Rank source lines of one kernel launch by warp-stall samples and executed instructions.
usage: python tools/ncu_hotlines.py REP [launch-skip] [top]   (needs --import-source on at capture time and -lineinfo at compile time)"""
import csv
import io
import os
import subprocess
import sys

rep, skip, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip),
                               "--launch-count", "1"], stderr=subprocess.DEVNULL).decode()
rows, cur_file, hdr, kernel = [], None, None, None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
    elif r[0] == "Function Name":
        kernel = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":   # a source line (SASS rows carry an address in column 2)
        d = dict(zip(hdr, r))
        rows.append((int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0), cur_file, int(r[0]), r[1].strip()[:100]))
ts, ti = sum(r[0] for r in rows) or 1, sum(r[1] for r in rows) or 1
print(f"# {kernel}\n# {ts} samples, {ti} warp instructions")
for s, i, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{100 * s / ts:5.1f}% smp {100 * i / ti:5.1f}% inst  {f}:{ln}  {src}")
