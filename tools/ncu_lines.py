#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.
usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [top_n]
Joins `ncu --page source --print-source=sass` (in instruction order) with `nvdisasm -g`
line markers of the same cubin."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin") and "index_build" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".text.") and kern in l)
cur, seq = ("?", 0), []
for l in dis[start + 1:]:
    s = l.strip()
    if s.startswith(".text."): break
    m = re.match(r'//## File "([^"]+)", line (\d+)', s)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
    if m: seq.append(cur)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg, aggs = collections.Counter(), collections.Counter()
stalls = collections.defaultdict(collections.Counter)
n = min(len(seq), len(data))
for k in range(n):
    agg[seq[k]] += int(data[k][ci]); aggs[seq[k]] += int(data[k][cs])
    for i in stall_cols:
        v = int(data[k][i] or 0)
        if v: stalls[seq[k]][hdr[i]] += v
tot, tots = sum(agg.values()), sum(aggs.values())
src = {}
base = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc")
for f in os.listdir(base):
    src[f] = open(os.path.join(base, f)).read().split("\n")
print("instructions %d  samples %d  (%d sass instrs, %d csv rows)" % (tot, tots, len(seq), len(data)))
allst = collections.Counter()
for k in stalls: allst.update(stalls[k])
print("stall mix:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / max(1, sum(allst.values()))) for k, v in allst.most_common(8)))
def text(f, ln):
    return src[f][ln - 1].strip()[:84] if f in src and ln - 1 < len(src[f]) else ""
print("--- by stall samples")
for (f, ln), v in aggs.most_common(top):
    top_st = ",".join("%s:%d%%" % (k.replace("stall_", ""), 100 * c / max(1, v)) for k, c in stalls[(f, ln)].most_common(2))
    print("%5.1f%% smp %5.1f%% inst %s:%d [%s] %s" % (100 * v / tots, 100 * agg[(f, ln)] / tot, f, ln, top_st, text(f, ln)))
print("--- by instructions")
for (f, ln), v in agg.most_common(top):
    print("%5.1f%% inst %5.1f%% smp %s:%d %s" % (100 * v / tot, 100 * aggs[(f, ln)] / tots, f, ln, text(f, ln)))
