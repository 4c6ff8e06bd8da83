"""profiles/<tag>_configs.md from the lines tools/configs_multi.py printed on 1 / 2 / 4 / 8 GPUs (gpurun_out/configs_n{1,2,4,8}.jsonl).
usage: python tools/make_configs_table.py <tag>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "gpurun_out", f"configs_n{n}.jsonl")
    if not os.path.exists(p):
        continue
    for ln in open(p):
        if ln.startswith("{"):
            j = json.loads(ln)
            lst = rows.setdefault(j["config"], {}).setdefault(n, [])
            lst[:] = [o for o in lst if o.get("exchange") != j.get("exchange")] + [j]  # a later line replaces an earlier one
out = [f"# {tag} - every BASELINE.json configuration on 1 / 2 / 4 / 8 B200 (tools/configs_multi.py, table by tools/make_configs_table.py)", "",
       "`python tools/configs_multi.py` (one GPU) and `python -m torch.distributed.run --nproc-per-node N tools/configs_multi.py` (N = 2, 4, 8),",
       "one call of `gpurun --gpus N` each. Frames: CUDA events per frame on the launching stream, 20 frames after 3 warm-up frames, L2 flushed",
       "(256 MiB fill) between frames outside the event pair, max over ranks, exchange (NVLink peer stores + flags, or the NCCL gather)",
       "inside the timed region. `Mrays/s` = logical rays of the frame (primary + one shadow ray per hit and light + bounce) / that time.",
       "C4 = batch queries on device-resident rays, contiguous 1/N chunks per rank, no exchange, best of 3, CUDA events, max over ranks.",
       "`== 1 GPU`: the frame rank 0 delivers is bit-identical to the frame it renders alone. CPU reference = oracle/_ref (the reference's",
       "own ray_tracing.cpp / bounding_volume_hierarchy.cpp) on all host threads of the same box, full frame, same run (N = 1 column).",
       "Pipeline: 3 = persistent wavefront `k_wave`, 2 = round pipeline (chosen per frame, DESIGN 4.4).", "",
       "| config | rays / frame (primary / shadow / bounce) | CPU reference Mrays/s | 1 GPU ms (Mrays/s) [pipeline] | 2 GPUs | 4 GPUs | 8 GPUs | speed-up at 8 | frames == 1 GPU | frame vs CPU reference |",
       "|---|---|---|---|---|---|---|---|---|---|"]


def cell(js):
    parts = []
    for j in js:
        t = f"{j['ms_per_frame']:.3f} ms ({j['Mrays_s']:.0f}) [{j['pipeline']}]"
        if len(js) > 1:
            t += f" {j['exchange']}"
        parts.append(t)
    return " / ".join(parts)


for name, by in rows.items():
    if name.startswith("C4"):
        continue
    j1 = by.get(1, [None])[0]
    same = [str(j["frame_equals_single_gpu"]) for n in (2, 4, 8) for j in by.get(n, [])]
    ok = "yes" if same and all(s == "True" for s in same) else ("n/a" if not same else "NO")
    cpu = f"{j1['cpu_reference_Mrays_s']} ({j1['cpu_threads']} thr)" if j1 and "cpu_reference_Mrays_s" in j1 else ""
    vs = (f"rays equal: {j1['cpu_rays_equal']}, max abs diff {j1['max_abs_vs_cpu']:.1g}, bit-equal pixels {100 * j1['bit_equal_frac_vs_cpu']:.3f} %"
          if j1 and "max_abs_vs_cpu" in j1 else "")
    any_j = (by.get(1) or by.get(2) or by.get(4) or by.get(8))[0]
    rays = f"{any_j['rays']:,.0f} ({any_j['primary']:,.0f} / {any_j['shadow']:,.0f} / {any_j['bounce']:,.0f})"
    sp = ""
    if j1 and 8 in by:
        sp = f"{j1['ms_per_frame'] / min(j['ms_per_frame'] for j in by[8]):.2f}x"
    out.append(f"| {name} | {rays} | {cpu} | {cell(by.get(1, []))} | {cell(by.get(2, []))} | {cell(by.get(4, []))} | {cell(by.get(8, []))} | {sp} | {ok} | {vs} |")
c4 = [by for name, by in rows.items() if name.startswith("C4")]
if c4:
    by = c4[0]
    j1 = by.get(1, [None])[0]
    out += ["", "| C4 (1 M-triangle soup, 16 777 216 incoherent rays) | CPU reference | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs | speed-up at 8 |", "|---|---|---|---|---|---|---|"]
    for key, label in (("closest", "closest hit"), ("any_inf", "any hit, range = inf"), ("any_u01", "any hit, range ~ U(0,1)")):
        cpu = ""
        if key == "closest" and j1 and "cpu_reference_Mrays_s" in j1:
            cpu = (f"{j1['cpu_reference_Mrays_s']} Mrays/s ({j1['cpu_threads']} thr, {j1['cpu_sample']}; GPU hit records bit-identical on it: "
                   f"{j1['sample_bit_identical']})")
        cells = [f"{by[n][0][key + '_ms']:.2f} ms ({by[n][0][key + '_Mrays_s']:.0f} Mrays/s)" if n in by else "" for n in (1, 2, 4, 8)]
        sp = f"{by[1][0][key + '_ms'] / by[8][0][key + '_ms']:.2f}x" if 1 in by and 8 in by else ""
        out.append(f"| {label} | {cpu} | " + " | ".join(cells) + f" | {sp} |")
timeouts = sum(j.get("handoff_timeouts", 0) or 0 for by in rows.values() for js in by.values() for j in js)
out += ["", f"Hand-off timeouts in all runs: {timeouts}."]
open(os.path.join(ROOT, "profiles", f"{tag}_configs.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
