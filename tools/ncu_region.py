"""Stall breakdown of one source-line range of a kernel (outer line of <file>, as tools/ncu_outer.py): where do the warps that
execute this region spend their time?  usage: ncu_region.py <rep> <kernel> <nvdisasm -g output> <file> lo hi"""
import collections
import csv
import re
import subprocess
import sys

rep, kname, sassfile, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
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
        addr2[int(m.group(1), 16)] = (outer, m.group(2))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
perline = collections.defaultdict(lambda: [0, 0])
base = None
ninst = nsmp = 0
for r in rows[1:]:
    try:
        a = int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else int(r[idx["Address"]])
    except Exception:
        continue
    if base is None:
        base = a
    line, sass = addr2.get(a - base, (None, ""))
    if line is None or not (lo <= line <= hi):
        continue
    ie = int(r[idx["Instructions Executed"]] or 0)
    s = int(r[idx["# Samples"]] or 0)
    ninst += ie
    nsmp += s
    perline[line][0] += ie
    perline[line][1] += s
    for h in stalls:
        tot[h] += int(r[idx[h]] or 0)
print(f"lines {lo}-{hi}: warp instructions {ninst}, samples {nsmp}, samples per 1000 instructions {1000.0 * nsmp / max(ninst, 1):.2f}")
for h, v in tot.most_common(10):
    print(f"  {100.0 * v / max(nsmp, 1):6.2f}%  {h}")
print("hottest lines (samples, instructions):")
for line, (ie, s) in sorted(perline.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  line {line}: {100.0 * s / max(nsmp, 1):5.1f}% samples {100.0 * ie / max(ninst, 1):5.1f}% inst")
