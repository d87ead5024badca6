"""Warp-stall samples per CUDA source line: python scripts/ncu_stalls.py <rep> <lib.so> <kernel-substr> [top]"""
import csv, re, subprocess, sys, tempfile, os, io
from collections import Counter, defaultdict
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l][0]
lines = []; group = []
for l in txt[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        if lines: break
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m: group.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if group: cur = group[0]; group = []
        lines.append((m.group(2).strip(), cur))
assert len(lines) == len(data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; per = Counter(); reason = defaultdict(Counter); allr = Counter()
for (ins, src), r in zip(lines, data):
    n = int(r[ix["# Samples"]]); tot += n; per[src] += n
    for h in stall_cols:
        v = int(r[ix[h]]); reason[src][h] += v; allr[h] += v
print("total samples", tot)
print("overall:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in allr.most_common(9)))
for src, n in per.most_common(top):
    f, ln = src
    try: text = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)).read().split("\n")[ln - 1].strip()[:70]
    except Exception: text = ""
    rs = ", ".join(f"{h[6:]} {100*v/n:.0f}%" for h, v in reason[src].most_common(3))
    print(f"{100*n/tot:5.1f}% {f}:{ln} [{rs}] {text}")
