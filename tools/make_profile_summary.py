"""Summarise an evidence pass (tools/gpu_profile.sh <tag>) into profiles/: launch-list shares, ncu --set full key metrics per
kernel (averaged over the captured launches) and profiles/traffic.json (DRAM bytes per launch, read by bench.py).
usage: python tools/make_profile_summary.py <tag>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_list(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1000 if u.startswith("n") else (v * 1000 if u.startswith("m") else v)
        k = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("cgrt::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def raw_page(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = collections.OrderedDict()
    for r in rows[2:]:
        k = r[idx["Kernel Name"]].split("(")[0]
        out.setdefault(k, []).append(r)
    return idx, units, out


WANT = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / warp inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__inst_executed.sum", "warp inst"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_bytes.sum", "L1 bytes"), ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors from L1"),
        ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local-memory load sectors (L1)"), ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local-memory store sectors (L1)"),
        ("smsp__inst_executed_op_local_ld.sum", "local loads (warp inst)"), ("smsp__inst_executed_op_local_st.sum", "local stores (warp inst)"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__cycles_active.avg", "SM sub-partition cycles active (avg)"), ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math")]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3,
         "nsecond": 1e-3}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


lines = [f"# {tag} - evidence pass on one B200 (tools/gpu_profile.sh {tag})", ""]
bj = os.path.join(G, f"bench_{tag}.json")
if os.path.exists(bj):
    j = json.loads([l for l in open(bj) if l.startswith("{")][-1])
    lines += [f"Bench line (`python bench.py`, exit 0 without ncu): **{j['value']:.1f} Mrays/s**, {j['ms_per_step']:.3f} ms/frame, e2e "
              f"{j['e2e']['value']:.1f} Mrays/s; roofline {j['roofline']['kernel']} frac {j['roofline']['frac']:.3f}; CPU reference "
              f"{j.get('cpu_baseline', {}).get('value', float('nan')):.2f} Mrays/s on {j.get('cpu_baseline', {}).get('cores', '?')} threads; "
              f"clocks {j['clocks']}.", ""]
ll = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(ll):
    agg = launch_list(ll)
    tot = sum(v[1] for v in agg.values())
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, same command; cold-cache, serialised)",
              "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f} % |")
    lines.append("")
    os.replace(ll, os.path.join(P, f"{tag}_launches.csv")) if not os.path.exists(os.path.join(P, f"{tag}_launches.csv")) else None
traffic = {}
for part in ("wave", "trace", "rest"):
    rp = os.path.join(G, f"prof_{tag}_{part}_raw.csv")
    if not os.path.exists(rp):
        continue
    idx, units, ker = raw_page(rp)
    lines += [f"## `ncu --set full --clock-control none` ({part}): mean over the captured launches", ""]
    for k, rows in ker.items():
        lines.append(f"### {k} ({len(rows)} launches)")
        lines += ["| metric | mean | min | max |", "|---|---|---|---|"]
        dram = 0.0
        extra = {}
        for m, label in WANT:
            if m not in idx:
                continue
            u = units[idx[m]]
            vals = [num(r[idx[m]]) * SCALE.get(u, 1.0) for r in rows]
            mean = sum(vals) / len(vals)
            unit = "us" if m.startswith("gpu__time") else ("B" if "bytes" in m else "")
            lines.append(f"| {label} | {mean:,.2f} {unit} | {min(vals):,.2f} | {max(vals):,.2f} |")
            if m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                dram += mean
            key = {"smsp__inst_executed.sum": "warp_inst_per_launch", "lts__t_bytes.sum": "l2_bytes_per_launch",
                   "l1tex__t_bytes.sum": "l1_bytes_per_launch", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_inst",
                   "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "gpu__time_duration.sum": "ncu_time_us"}.get(m)
            if key:
                extra[key] = mean
        traffic[k] = {"dram_bytes_per_launch": dram, "launches_captured": len(rows), "source": f"profiles/{tag}_summary.md", **extra}
        lines.append("")
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
if traffic:
    tp = os.path.join(P, "traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update(traffic)  # kernels not captured in this pass keep their previous entry (its `source` says which pass)
    json.dump(old, open(tp, "w"), indent=1)
print("\n".join(lines[:60]))
