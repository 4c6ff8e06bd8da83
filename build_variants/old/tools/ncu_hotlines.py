"""Aggregate ncu warp-stall samples of one kernel by CUDA source line.
usage: ncu_hotlines.py <rep> <kernel-substring> <nvdisasm -g output>
Maps SASS addresses (ncu --page source --csv) to the //## File ... line N markers of nvdisasm -g."""
import csv, re, subprocess, sys, collections
rep, kname, sassfile = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the csv has a first line with the kernel name, then a header, then rows
lines = raw.splitlines()
rows = list(csv.reader(lines[1:]))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
# address -> (file, line) from nvdisasm
text = open(sassfile).read()
start = text.index(".text." + [m for m in re.findall(r"\.text\.(\S+):", text) if kname in m][0] + ":")
seg = text[start:]
end = seg.find("//--------------------- .", 10)
seg = seg[:end] if end > 0 else seg
cur = ("?", 0); addr2line = {}; stack = []
for ln in seg.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
agg = collections.Counter(); ins = collections.Counter(); tot = 0; totins = 0
base = None
for r in rows[1:]:
    try:
        a = int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else int(r[idx["Address"]])
    except Exception:
        continue
    if base is None: base = a
    off = a - base
    smp = int(r[idx["# Samples"]] or 0); ie = int(r[idx["Instructions Executed"]] or 0)
    key = addr2line.get(off, (("?", 0), ""))[0]
    agg[key] += smp; ins[key] += ie; tot += smp; totins += ie
print("total samples", tot, "total warp instructions", totins)
for key, v in agg.most_common(45):
    print("%6.2f%% samples  %6.2f%% inst   %s:%d" % (100.0 * v / tot, 100.0 * ins[key] / max(totins, 1), key[0], key[1]))
