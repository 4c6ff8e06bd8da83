"""Debug: per-kernel scheduling counters of the persistent warps (library built with -DCGRT_INSTRUMENT)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
capi = ge.load_package().capi
lib = capi.load_library()
d = capi.dragon_standin()
s = capi.Scene(d, lights=d.lights)
W, H, L = 1920, 1080, 5
cam = capi.make_camera(W, H)
s.render(cam, W, H, trace_limit=L)
out = (C.c_ulonglong * 16)()
lib.cgrt_debug_instrumentation(out, 1)
tl = (C.c_uint * 512)()
lib.cgrt_debug_timeline(tl, 1)
sh = (C.c_uint * 64)()
lib.cgrt_debug_step_hist(sh, 1)
_, st = s.render(cam, W, H, trace_limit=L)
lib.cgrt_debug_instrumentation(out, 1)
v = [int(x) for x in out]
print("stats", st)
print("warps %d iterations %d avg running lanes/iter %.2f" % (v[15], v[0], v[1] / max(v[0], 1)))
for k, name in enumerate(("REF", "WIDE", "LEAF")):
    print("  class %-8s chosen %9d iterations (%.1f%%), avg lanes stepped %.2f" % (name, v[2 + k], 100.0 * v[2 + k] / max(v[0], 1), v[5 + k] / max(v[2 + k], 1)))
print("refill rounds %d lanes %d (%.1f/round); retire rounds %d lanes %d (%.1f/round)" % (v[8], v[9], v[9] / max(v[8], 1), v[10], v[11], v[11] / max(v[10], 1)))
lib.cgrt_debug_step_hist(sh, 1)
sh = np.array(list(sh)); cum = np.cumsum(sh) / max(sh.sum(), 1)
print("k_trace steps per ray (bucket of 8 steps: rays): " + " ".join("%d:%d" % (8 * b, sh[b]) for b in range(64) if sh[b]))
print("  mean steps (bucket mid) %.1f; 50/90/99/99.9%% below %s steps" % (float((sh * (np.arange(64) * 8 + 4)).sum() / max(sh.sum(), 1)), [int(8 * (np.searchsorted(cum, q) + 1)) for q in (0.5, 0.9, 0.99, 0.999)]))
tot = max(v[12] + v[13] + v[14], 1)
print("cycles: steps %.1f%% refill %.1f%% retire %.1f%%; per warp total %.0f cycles; per iteration %.0f cycles" % (100.0 * v[12] / tot, 100.0 * v[13] / tot, 100.0 * v[14] / tot, tot / max(v[15], 1), v[12] / max(v[0], 1)))

lib.cgrt_debug_timeline(tl, 1)
tl = np.array(list(tl)).reshape(2, 128, 2)
for k, name in enumerate(("k_paths", "k_shadow_all")):
    print(name, "timeline (25 us buckets): bursts | avg running lanes per bursting warp")
    last = max([b for b in range(128) if tl[k, b, 0] > 0] + [0])
    print("  " + " ".join("%d:%d|%.0f" % (b, tl[k, b, 0], tl[k, b, 1] / max(tl[k, b, 0], 1)) for b in range(last + 1)))
