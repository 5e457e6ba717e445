#!/usr/bin/env python
"""Like ncu_lines.py, but per source line of one file: ncu_lines_exact.py <rep> <kernel regex> <disassembly> <section substring> <file> <line lo> <line hi> [launch-skip]"""
import csv, io, re, subprocess, sys, collections
rep, kre, dis, sect, fname, lo, hi = sys.argv[1:8]
lo, hi = int(lo), int(hi)
skip = sys.argv[8] if len(sys.argv) > 8 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for k in range(2, len(rows)):
    if rows[k] and rows[k][0] == "Kernel Name":
        rows = rows[:k]; break
hdr = rows[1]; col = {k: i for i, k in enumerate(hdr)}; ins = rows[2:]
cur = None; line = None; lines = []
for l in open(dis):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m: cur = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if cur and sect in cur and re.match(r'\s+/\*[0-9a-f]+\*/', l): lines.append(line)
n = min(len(ins), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); tot = [0, 0, 0]
for i in range(n):
    r = ins[i]
    ie = float(r[col["Instructions Executed"]] or 0); te = float(r[col["Thread Instructions Executed"]] or 0); sm = float(r[col["# Samples"]] or 0)
    tot[0] += ie; tot[1] += te; tot[2] += sm
    a = agg[lines[i]]; a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
src = open([p for p in ("actinon_b200/csrc/" + fname,) ][0]).read().splitlines()
for ln in range(lo, hi + 1):
    a = agg.get((fname, ln))
    if a: print("%4d %5.1f%% smp %5.1f%% inst %3d sass thr/inst %4.1f | %s" % (ln, 100 * a[2] / tot[2], 100 * a[0] / tot[0], a[3], a[1] / max(a[0], 1), src[ln - 1][:110]))
    else: print("%4d %35s | %s" % (ln, "", src[ln - 1][:110]))
