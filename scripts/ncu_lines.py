"""Attribute ncu per-SASS 'Instructions Executed' to CUDA source lines (innermost inline frame).
usage: python scripts/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, io
from collections import Counter
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
execs = [int(r[ix["Instructions Executed"]]) for r in data]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l][0]
lines = []; group = []; 
for l in txt[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        if lines: break
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m:
        group.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if group: cur = group[0]; chain = tuple(group); group = []
        lines.append((m.group(2).strip(), cur, chain))
assert len(lines) == len(data), (len(lines), len(data))
tot = sum(execs); c = Counter(); ops = {}
for (ins, src, chain), n in zip(lines, execs):
    c[src] += n
srcs = {}
for k, v in c.most_common(top):
    f, ln = k
    path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
    try: text = open(path).read().split("\n")[ln - 1].strip()[:90]
    except Exception: text = ""
    print(f"{100*v/tot:5.1f}% {v/1e6:10.0f}M {f}:{ln}  {text}")
