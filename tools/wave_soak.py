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
pinned = {}
if os.environ.get("SOAK_PINNED"):  # page-locked destinations: the overlapped delivery of cgrt_render (one buffer per frame size)
    import ctypes as C
    lib = capi.load_library()
    for (_, W, H, _, _, _) in shapes:
        if (W, H) not in pinned:
            ptr = C.c_void_p()
            capi.check(lib.cgrt_host_alloc_pinned(W * H * 12, C.byref(ptr)))
            pinned[(W, H)] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(H, W, 3))
first = {}
if os.environ.get("SOAK_STREAM"):  # streaming form: two frames (of different shapes) in flight, then wait and compare both
    import ctypes as C
    lib = capi.load_library()
    one = [sh for sh in shapes if sh[5] == 1]
    bufs = {}
    for k, (name, W, H, L, _, _) in enumerate(one):
        for slot in (0, 1):
            ptr = C.c_void_p()
            capi.check(lib.cgrt_host_alloc_pinned(W * H * 12, C.byref(ptr)))
            bufs[(k, slot)] = (ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(H, W, 3)))
    bad = 0
    t0 = time.time()
    for it in range(n_frames // 2):
        pair = [int(rng.integers(len(one))), int(rng.integers(len(one)))]
        same_scene = one[pair[0]][0] == one[pair[1]][0]
        for slot, k in enumerate(pair):
            name, W, H, L, _, _ = one[k]
            bufs[(k, slot)][1][:] = 0.5
            scenes[name].render_submit(capi.make_camera(W, H), capi.render_params(W, H, L), bufs[(k, slot)][0].value)
            if not same_scene:
                scenes[name].render_wait()
        scenes[one[pair[0]][0]].render_wait()
        for slot, k in enumerate(pair):
            rgb = bufs[(k, slot)][1]
            if k not in first:
                first[k] = rgb.copy()
            elif not np.array_equal(rgb.view(np.uint32), first[k].view(np.uint32)):
                bad += 1
                print("STREAMED FRAME DIFFERS", one[k], "iteration", it, "slot", slot, flush=True)
    print(f"soak (streaming): {2 * (n_frames // 2)} frames, {len(first)} shapes, {bad} differing, {time.time() - t0:.1f} s")
    sys.exit(1 if bad else 0)
t0 = time.time()
bad = 0
for k in range(n_frames):
    name, W, H, L, rank, world = shapes[int(rng.integers(len(shapes)))] if k >= len(shapes) else shapes[k]
    s = scenes[name]
    cam = capi.make_camera(W, H)
    out = pinned.get((W, H))
    if out is not None:
        out[:] = 0.25 if world == 1 else 0.0  # (the delivery must overwrite whatever the buffer held)
    rgb, st = s.render(cam, W, H, trace_limit=L, rank=rank, world=world, out=out)
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
