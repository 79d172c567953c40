"""Aggregate one kernel launch of an .ncu-rep by source file and by (file, 20-line bucket): executed warp instructions and stall samples
other than barrier waits.  usage: python tools/ncu_byfile.py REP [launch-skip]"""
import collections
import csv
import io
import os
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip),
                               "--launch-count", "1"], stderr=subprocess.DEVNULL).decode()
byfile, bybucket = collections.Counter(), collections.Counter()
sfile, sbucket = collections.Counter(), collections.Counter()
cur, hdr = None, None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        inst = int(d.get("Instructions Executed") or 0)
        smp = int(d.get("# Samples") or 0) - int(d.get("stall_barrier") or d.get("Stall Barrier") or 0)
        byfile[cur] += inst
        bybucket[(cur, int(r[0]) // 20 * 20)] += inst
        sfile[cur] += smp
        sbucket[(cur, int(r[0]) // 20 * 20)] += smp
ti, ts = sum(byfile.values()) or 1, sum(sfile.values()) or 1
print("header columns:", [h for h in (hdr or []) if "tall" in h or "ampl" in h][:12])
print(f"# {ti} warp instructions, {ts} samples (barrier waits excluded where the column exists)")
for f, n in byfile.most_common(12):
    print(f"{100 * n / ti:5.1f}% inst {100 * sfile[f] / ts:5.1f}% smp  {f}")
print()
for (f, b), n in bybucket.most_common(40):
    print(f"{100 * n / ti:5.1f}% inst {100 * sbucket[(f, b)] / ts:5.1f}% smp  {f}:{b}-{b + 19}")
