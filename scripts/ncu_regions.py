"""Warp-stall samples and executed instructions per REGION of one_sim (call-site line inside one_sim_impl, following
the inline chain outwards): python scripts/ncu_regions.py <rep> <lib.so> <kernel-substr> name:line,name:line,...,end:line
A region runs from its first line of csrc/mcts.cuh (as compiled into the profiled library) to the next region's."""
import csv, re, subprocess, sys, tempfile, os, io
from collections import Counter
rep, lib, kern = sys.argv[1:4]
marks = [(int(x.split(":")[1]), x.split(":")[0]) for x in sys.argv[4].split(",")]
lo = marks[0][0]; hi = marks[-1][0]; marks = marks[:-1]
def region_of(line):
    name = None
    for ln, nm in marks:
        if ln <= line: name = nm
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l][0]
lines = []; group = []; chain = ()
for l in txt[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        if lines: break
    m = re.match(r'\s*//## File "(.*?)", line (\d+)(?: inlined at "(.*?)", line (\d+))?', l)
    if m:
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        if m.group(3): group.append((m.group(3).split("/")[-1], int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if group: chain = tuple(group); group = []
        lines.append((m.group(2).strip(), chain))
assert len(lines) == len(data), (len(lines), len(data))
samp = Counter(); inst = Counter(); tot_s = tot_i = 0
for (ins, chain), r in zip(lines, data):
    n = int(r[ix["# Samples"]]); e = int(r[ix["Instructions Executed"]])
    reg = "outside one_sim"
    for f, ln in reversed(chain):  # outermost frame first
        if f == "mcts.cuh" and lo <= ln < hi:
            reg = region_of(ln); break
    samp[reg] += n; inst[reg] += e; tot_s += n; tot_i += e
print(f"total samples {tot_s}, instructions {tot_i/1e9:.2f} G")
for reg, n in samp.most_common():
    print(f"{100*n/tot_s:5.1f}% samples  {100*inst[reg]/tot_i:5.1f}% inst ({inst[reg]/1e6:8.0f} M)  {reg}")
