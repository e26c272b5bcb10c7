#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line using nvdisasm -g line info.
usage: ncu_by_line.py <src.csv from `ncu -i rep --page source --csv`> <cubin> <kernel-substring> [top]"""
import csv, re, subprocess, sys, collections
srccsv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
off2line, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
ia, iall, iex = hdr.index("Address"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
recs = []
for r in rows[2:]:
    try:
        recs.append((int(r[ia], 16), int(r[iall] or 0), int(r[iex] or 0)))
    except Exception:
        pass
base = min(a for a, _, _ in recs)
agg = collections.Counter(); ex = collections.Counter()
for a, s, e in recs:
    k = off2line.get(a - base)
    agg[k] += s; ex[k] += e
tot = sum(agg.values())
src = {}
for k in agg:
    if k and k[0] not in src:
        try: src[k[0]] = open("/root/repo/ntm_tracker_b200/csrc/" + k[0]).read().splitlines()
        except Exception: src[k[0]] = []
print("total stall samples", tot)
for k, s in agg.most_common(top):
    text = ""
    if k and src.get(k[0]) and k[1] - 1 < len(src[k[0]]):
        text = src[k[0]][k[1] - 1].strip()[:95]
    print("%5d %5.1f%% inst=%9d %-22s %s" % (s, 100.0 * s / tot, ex[k], "%s:%d" % k if k else "?", text))
