#!/usr/bin/env python
"""Joins ncu's per-SASS-instruction page with nvdisasm line info (same build) and prints the hottest source lines.
usage: ncu_lines.py <rep> <kernel regex> <cubin-disassembly (nvdisasm -g -c)> <section substring> [launch-skip]"""
import csv, io, re, subprocess, sys, collections
rep, kre, dis, sect = sys.argv[1:5]
skip = sys.argv[5] if len(sys.argv) > 5 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for k in range(2, len(rows)):                 # some ncu versions print the kernel twice
    if rows[k] and rows[k][0] == "Kernel Name":
        rows = rows[:k]; break
hdr = rows[1]
col = {k: i for i, k in enumerate(hdr)}
ins = rows[2:]
# line info
cur = None; line = None; lines = []
for l in open(dis):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m: cur = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if cur and sect in cur and re.match(r'\s+/\*[0-9a-f]+\*/', l):
        lines.append(line)
print("sass rows", len(ins), "disasm rows", len(lines))
n = min(len(ins), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for i in range(n):
    r = ins[i]
    ie = float(r[col["Instructions Executed"]] or 0); te = float(r[col["Thread Instructions Executed"]] or 0); sm = float(r[col["# Samples"]] or 0)
    a = agg[lines[i]]; a[0] += ie; a[1] += te; a[2] += sm
    tot[0] += ie; tot[1] += te; tot[2] += sm
print("total warp-inst %.3g thread-inst %.3g avg threads %.1f samples %d" % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
byfunc = collections.defaultdict(lambda: [0, 0, 0])
for (k, a) in agg.items():
    b = byfunc[(k[0], k[1] // 10 * 10) if k else None]
    for j in range(3): b[j] += a[j]
print("---- hottest 10-line buckets by stall samples")
for k, a in sorted(byfunc.items(), key=lambda x: -x[1][2])[:45]:
    print(k, "samples %.1f%% inst %.1f%% threads/inst %.1f" % (100 * a[2] / tot[2], 100 * a[0] / tot[0], a[1] / max(a[0], 1)))
