"""A/B of the two production pipelines on one GPU: the persistent wavefront (k_wave, all three scheduling modes) must deliver
the frames of the round pipeline bit for bit, with the same ray counters. Prints device ms per frame for each.
    python tools/wave_check.py            # driver: one subprocess per pipeline / mode (the choice is read once per process)
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [("soup", 64, 64, 2), ("cornell", 512, 512, 2), ("monkey", 640, 360, 1), ("dragon", 480, 270, 5), ("dragon", 1920, 1080, 5),
         ("dodge", 1920, 1080, 2)]


def worker():
    import __graft_entry__ as ge
    from oracle import bindings as ob
    from conftest import load_golden
    capi = ge.load_package().capi
    out = {}
    scenes = {}
    for name, W, H, L in CASES:
        if name not in scenes:
            if name == "soup":
                flat, lights = ob.random_soup(3000, seed=7, scale=0.08, n_meshes=3), np.array([[0.0, 0.9, 0.0, 1, 1, 1]], np.float32)
            elif name == "dragon":
                flat, lights = ob.dragon_standin_fixture()
            else:
                g = load_golden(name)
                flat, lights = g.flat, g.lights
            scenes[name] = capi.Scene(flat, lights=lights, device=0)
        s = scenes[name]
        cam = capi.make_camera(W, H)
        rgb, st = s.render(cam, W, H, trace_limit=L)
        ms = []
        for _ in range(5):
            _, st2 = s.render(cam, W, H, trace_limit=L)
            ms.append(st2["device_ms"])
        out[f"{name}_{W}x{H}_L{L}"] = dict(sha=hashlib.sha1(np.ascontiguousarray(rgb).tobytes()).hexdigest(),
                                          rays=[st[k] for k in ("primary", "primary_hit", "shadow", "bounce")],
                                          replay=[st["replayed_closest"], st["replayed_shadow"]], ms=round(min(ms), 4),
                                          pipeline=st["pipeline"], launches=st["kernel_launches"])
    print("RESULT " + json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        return worker()
    runs = [("rounds", {"CGRT_PIPELINE": "rounds"}), ("wave auto", {"CGRT_PIPELINE": "wave", "CGRT_WAVE": "mode=0"}),
            ("wave lane", {"CGRT_PIPELINE": "wave", "CGRT_WAVE": "mode=1,switch=0"}), ("wave group", {"CGRT_PIPELINE": "wave", "CGRT_WAVE": "mode=2"}),
            ("automatic", {})]
    for extra in filter(None, os.environ.get("WAVE_CHECK_EXTRA", "").split(";")):  # e.g. "mode=1,fin=12;mode=2,fin=12"
        runs.append((f"wave {extra}", {"CGRT_PIPELINE": "wave", "CGRT_WAVE": extra}))
    res = {}
    for label, env in runs:
        e = dict(os.environ)
        e.update(env)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=e, capture_output=True, text=True, timeout=240)
        except subprocess.TimeoutExpired:
            print(f"[wave_check] {label}: TIMEOUT")
            continue
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if r.returncode != 0 or not line:
            print(f"[wave_check] {label}: FAILED rc={r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-3000:]}")
            continue
        res[label] = json.loads(line[0][7:])
    ok = True
    base = res.get("rounds")
    for label, r in res.items():
        for case, v in r.items():
            same = base is not None and v["sha"] == base[case]["sha"] and v["rays"] == base[case]["rays"]
            ok &= same
            print(f"[wave_check] {label:12s} {case:22s} pipeline {v['pipeline']} launches {v['launches']:3d} {v['ms']:8.4f} ms  rays {v['rays']} "
                  f"replay {v['replay']}  {'== rounds' if same else 'MISMATCH'}")
    print("[wave_check] " + ("OK" if ok and len(res) == len(runs) else "FAILED"))
    sys.exit(0 if ok and len(res) == len(runs) else 1)


if __name__ == "__main__":
    main()
