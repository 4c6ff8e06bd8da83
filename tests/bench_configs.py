"""Device-time table for every BASELINE.json configuration (C1..C5) on one GPU. Not the contract bench (bench.py); used to
fill the per-config numbers in profiles/ and README."""
import os, sys, time, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # (lives under tests/: it uses the oracle's scene generators and the golden fixtures)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
capi = ge.load_package().capi
from conftest import load_golden
from oracle import bindings as ob

def frames(name, flat, lights, W, H, L, reps=10):
    s = capi.Scene(flat, lights=lights)
    cam = capi.make_camera(W, H)
    s.render(cam, W, H, trace_limit=L)
    ms = []
    for _ in range(reps):
        _, st = s.render(cam, W, H, trace_limit=L)
        ms.append(st["device_ms"])
    _, stp = s.render(cam, W, H, trace_limit=L, flags=capi.RENDER_PROFILE_ALL)
    _, stc = s.render(cam, W, H, trace_limit=L, flags=capi.RENDER_COUNT)
    rays = st["primary"] + st["shadow"] + st["bounce"]
    byt = 48 * rays + 32 * sum(stc["box_tests"]) + 48 * sum(stc["tri_tests"])
    best = min(ms)
    print(json.dumps(dict(config=name, W=W, H=H, trace_limit=L, rays=rays, primary_hit=st["primary_hit"], shadow=st["shadow"], bounce=st["bounce"],
                          device_ms=round(best, 4), Mrays_s=round(rays / best / 1e3, 1), bytes_per_ray=round(byt / rays, 1),
                          alg_GBps=round(byt / best / 1e6, 1), frac_hbm=round(byt / best / 1e6 / 6537.6, 4), launches=st["kernel_launches"],
                          class_ms=dict(zip(capi.class_names(stp), [round(v, 3) for v in stp["class_ms"]])),
                          replayed=[stp["replayed_closest"], stp["replayed_shadow"]])))
    s.close()
    if os.environ.get("CGRT_CONFIGS_EXACT_TOO"):  # the exact path pipeline on the same frame, for comparison
        s = capi.Scene(flat, lights=lights, exact_only=True)
        s.render(cam, W, H, trace_limit=L)
        ms = [s.render(cam, W, H, trace_limit=L)[1]["device_ms"] for _ in range(reps)]
        print(json.dumps(dict(config=name + " (exact-only path pipeline)", device_ms=round(min(ms), 4))))
        s.close()

g = load_golden("cornell"); frames("C1 cornell", g.flat, g.lights, 512, 512, 2)
g = load_golden("monkey"); frames("C2 monkey", g.flat, g.lights, 1920, 1080, 1)
d = capi.dragon_standin(); frames("C3 dragon stand-in", d, d.lights, 1920, 1080, 5)
g = load_golden("dodge"); frames("C5 dodge", g.flat, g.lights, 3840, 2160, 2)

# C4: 1 M-triangle soup, 16 M incoherent rays, device-resident batch queries
lib = capi.load_library()
flat = ob.random_soup(1_000_000, seed=1234, scale=0.01, smooth_normals=False)
s = capi.Scene(flat)
n = 16 * 1024 * 1024
rays = ob.random_rays(n, seed=5678)
dR, dH, dM, dO, dC = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
for p, b in ((dR, n * 32), (dH, n * 32), (dM, n * 4), (dO, n), (dC, n * 8)):
    capi.check(lib.cgrt_device_malloc(0, b, C.byref(p)))
capi.check(lib.cgrt_memcpy_h2d(0, dR, C.c_void_p(rays.ctypes.data), n * 32))
md = np.full(n, np.inf, np.float32)
capi.check(lib.cgrt_memcpy_h2d(0, dM, C.c_void_p(md.ctypes.data), n * 4))
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        lib.cgrt_device_synchronize(0); t0 = time.perf_counter(); fn(); lib.cgrt_device_synchronize(0)
        best = min(best, time.perf_counter() - t0)
    return best
tc = timed(lambda: capi.check(lib.cgrt_intersect_closest_device(s.h, dR, n, dH, None, None)))
ta = timed(lambda: capi.check(lib.cgrt_intersect_any_device(s.h, dR, dM, C.c_float(0.001), n, dO, None)))
capi.check(lib.cgrt_intersect_closest_device(s.h, dR, n, dH, dC, None)); lib.cgrt_device_synchronize(0)
cnt = np.zeros((n, 2), np.uint32); capi.check(lib.cgrt_memcpy_d2h(0, C.c_void_p(cnt.ctypes.data), dC, n * 8))
hits = np.zeros(n, capi.HIT_DTYPE); capi.check(lib.cgrt_memcpy_d2h(0, C.c_void_p(hits.ctypes.data), dH, n * 32))
byt = 48.0 * n + 32.0 * cnt[:, 0].sum(dtype=np.float64) + 48.0 * cnt[:, 1].sum(dtype=np.float64)
print(json.dumps(dict(config="C4 soup 1M tris / 16M incoherent rays", closest_ms=round(tc * 1e3, 2), closest_Mrays_s=round(n / tc / 1e6, 1),
                      any_ms=round(ta * 1e3, 2), any_Mrays_s=round(n / ta / 1e6, 1), hit_frac=round(float((hits["tri"] >= 0).mean()), 4),
                      box_per_ray=round(float(cnt[:, 0].mean()), 1), tri_per_ray=round(float(cnt[:, 1].mean()), 1),
                      bytes_per_ray=round(byt / n, 1), closest_alg_GBps=round(byt / tc / 1e9, 1), closest_frac_hbm=round(byt / tc / 1e9 / 6537.6, 3))))
