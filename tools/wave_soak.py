"""Soak of the persistent wavefront: many frames of mixed shapes, every one compared bit for bit with the first frame of its
shape (a protocol race would show as a watchdog error or a differing frame).  python tools/wave_soak.py [frames]
(knobs through CGRT_WAVE as usual, e.g. mode=1,switch=400000 to exercise the change-over with a backlog)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from conftest import load_golden  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 600
capi = ge.load_package().capi
flat, lights = ob.dragon_standin_fixture()
scenes = {"dragon": capi.Scene(flat, lights=lights, device=0)}
g = load_golden("cornell")
scenes["cornell"] = capi.Scene(g.flat, lights=g.lights, device=0)
only = os.environ.get("SOAK_ONLY")
shapes = [("dragon", 1920, 1080, 5, 0, 1), ("dragon", 640, 360, 5, 0, 1), ("dragon", 1920, 1080, 5, 3, 8), ("dragon", 1920, 1080, 3, 1, 2),
          ("cornell", 512, 512, 2, 0, 1), ("dragon", 333, 201, 4, 0, 1), ("dragon", 1280, 720, 5, 0, 1), ("cornell", 800, 800, 5, 0, 1)]
if only:
    shapes = [shapes[int(i)] for i in only.split(",")]
rng = np.random.default_rng(1)
first = {}
t0 = time.time()
bad = 0
for k in range(n_frames):
    name, W, H, L, rank, world = shapes[int(rng.integers(len(shapes)))] if k >= len(shapes) else shapes[k]
    s = scenes[name]
    cam = capi.make_camera(W, H)
    rgb, st = s.render(cam, W, H, trace_limit=L, rank=rank, world=world)
    key = (name, W, H, L, rank, world)
    sig = (rgb.view(np.uint32).sum(dtype=np.uint64), float(rgb.sum()), st["shadow"], st["bounce"], st["primary_hit"])
    if key not in first:
        first[key] = (rgb.copy(), sig)
    elif sig != first[key][1] or not np.array_equal(rgb.view(np.uint32), first[key][0].view(np.uint32)):
        bad += 1
        d = np.argwhere((rgb.view(np.uint32) != first[key][0].view(np.uint32)).any(axis=2))
        y, x = d[0]
        print("FRAME DIFFERS", key, "frame", k, "pixels", len(d), "first at (x, y)", int(x), int(y), "got", rgb[y, x], "want", first[key][0][y, x],
              "stats", sig[2:], "want", first[key][1][2:], "replays", st["replayed_closest"], st["replayed_shadow"], flush=True)
        for yy, xx in d[:6]:
            print("   ", int(xx), int(yy), rgb[yy, xx], first[key][0][yy, xx], flush=True)
print(f"soak: {n_frames} frames, {len(first)} shapes, {bad} differing, {time.time() - t0:.1f} s, CGRT_WAVE={os.environ.get('CGRT_WAVE', '')!r}")
sys.exit(1 if bad else 0)
