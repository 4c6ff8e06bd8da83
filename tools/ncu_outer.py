"""Aggregate ncu per-instruction counters of one kernel by the OUTER source line of a chosen file: every SASS instruction is
attributed to the most recent line of <file> seen in address order (inlined callee code follows its call site), then lines are
bucketed by the ranges given on the command line.
usage: ncu_outer.py <rep> <kernel-substring> <nvdisasm -g output> <file> name:lo-hi [name:lo-hi ...]"""
import collections
import csv
import re
import subprocess
import sys

rep, kname, sassfile, fname = sys.argv[1:5]
buckets = []
for b in sys.argv[5:]:
    n, r = b.split(":")
    lo, hi = r.split("-")
    buckets.append((n, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()[1:]))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
text = open(sassfile).read()
start = text.index(".text." + [m for m in re.findall(r"\.text\.(\S+):", text) if kname in m][0] + ":")
seg = text[start:]
end = seg.find("//--------------------- .", 10)
seg = seg[:end] if end > 0 else seg
outer = None
addr2 = {}
for ln in seg.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if m.group(1).endswith(fname):
            outer = int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr2[int(m.group(1), 16)] = outer
inst = collections.Counter()
smp = collections.Counter()
thr = collections.Counter()
base = None
tot_i = tot_s = 0
for r in rows[1:]:
    try:
        a = int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else int(r[idx["Address"]])
    except Exception:
        continue
    if base is None:
        base = a
    line = addr2.get(a - base)
    name = "other"
    if line is not None:
        for n, lo, hi in buckets:
            if lo <= line <= hi:
                name = n
                break
        else:
            name = f"line{line}"
    ie = int(r[idx["Instructions Executed"]] or 0)
    s = int(r[idx["# Samples"]] or 0)
    te = int(r[idx["Thread Instructions Executed"]] or 0) if "Thread Instructions Executed" in idx else 0
    inst[name] += ie
    smp[name] += s
    thr[name] += te
    tot_i += ie
    tot_s += s
print("total warp instructions", tot_i, "samples", tot_s)
for n, v in inst.most_common(40):
    print(f"{100.0 * v / max(tot_i, 1):6.2f}% inst {100.0 * smp[n] / max(tot_s, 1):6.2f}% samples  threads/inst {thr[n] / max(v, 1):5.1f}  {n}")
