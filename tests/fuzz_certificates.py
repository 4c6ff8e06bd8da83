"""Long-running CPU fuzz campaign for the speculative traversal's certificates (not collected by pytest; run by hand:
python tests/fuzz_certificates.py [rounds]). Same harness as test_spec_certificate_cpu.py, with scaled / translated scenes,
far origins, tiny direction components and tiny ray bounds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import test_spec_certificate_cpu as T
from oracle import bindings as ob

lib = C.CDLL(os.path.join(T.HERE, "libspec_harness.so"))
lib.spec_run.restype = C.c_int
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tot = bad = 0
for rnd in range(rounds):
    rng = np.random.default_rng(1000 + rnd)
    for kind in T.SCENES:
        flat, grid = T.scene(kind)
        scale = np.float32(10.0 ** rng.integers(-3, 4))
        shift = (rng.uniform(-1, 1, 3) * scale * (10.0 if rng.random() < 0.3 else 0.0)).astype(np.float32)
        v = flat.vertices.copy()
        v[:, :3] = v[:, :3] * scale + shift
        f2 = ob.FlatScene(flat.vcount, flat.tcount, v, flat.triangles, flat.materials, flat.spheres)
        rays = T.ray_mix(flat, seed=rnd * 31 + len(kind), n=120000, grid=grid)
        rays["o"] = rays["o"] * scale + shift
        fin = rays["t"] < 1e30
        rays["t"][fin] *= scale
        # far origins and tiny direction components on a part of the rays
        k = len(rays) // 6
        idx = rng.choice(len(rays), k, replace=False)
        rays["o"][idx] -= rays["d"][idx] * np.float32(50.0) * scale
        idx = rng.choice(len(rays), k, replace=False)
        comp = rng.integers(0, 3, k)
        rays["d"][idx, comp] *= np.float32(10.0) ** rng.integers(-14, -3, k).astype(np.float32)
        idx = rng.choice(len(rays), k // 4, replace=False)
        rays["t"][idx] = (rng.uniform(0, 0.05, len(idx)) * scale).astype(np.float32)
        for sah in (True, False):
            ex, fa, cert, st = T.run(lib, f2, rays, mode=0, sah=sah)
            tot += st["rays"]; bad += st["mismatch"]
            line = f"round {rnd} {kind:12s} {'sah' if sah else 'ref'} scale {scale:g} rays {st['rays']} deferred {st['deferred'] / st['rays']:.5f} mismatch {st['mismatch']}"
            if st["mismatch"]:
                i = st["first"]
                line += f"  FIRST {i}: ray {rays[i]} exact {ex[i]} fast {fa[i]}"
            print(line, flush=True)
            far = rays.copy(); far["t"] = T.FLT_MAX
            md = (rng.uniform(0, 2, len(rays)) * scale).astype(np.float32)
            _, _, _, sa = T.run(lib, f2, far, mode=1, max_dist=md, eps=0.001, sah=sah)
            tot += sa["rays"]; bad += sa["mismatch"]
            if sa["mismatch"]:
                print("   ANY-HIT mismatch", sa, flush=True)
print("TOTAL rays", tot, "mismatches", bad)

# ---- BVH depths 1..33 on tiny / multi-mesh scenes, every traversal mode (0/1 speculative + certificate, 2/3 exact replay)
tot2 = bad2 = 0
cases = [("soup1", ob.random_soup(1, seed=1, scale=0.5)), ("soup2", ob.random_soup(2, seed=2, scale=0.5)),
         ("soup9x3", ob.random_soup(9, seed=3, scale=0.4, n_meshes=3)), ("soup300x300", ob.random_soup(300, seed=4, scale=0.3, n_meshes=300)),
         ("soup3000", ob.random_soup(3000, seed=5, scale=0.1)), ("boxes", T.box_walls()), ("grid", T.grid_planes(8)[0])]
for name, flat in cases:
    rays = T.ray_mix(flat, seed=11, n=60000)
    far = rays.copy(); far["t"] = T.FLT_MAX
    md = np.random.default_rng(3).uniform(0, 2, len(rays)).astype(np.float32)
    for depth in (1, 2, 3, 12, 20, 33):
        for sah in (True, False):
            for mode in (0, 2):
                st = T.run(lib, flat, rays, mode=mode, sah=sah, depth=depth)[3]
                tot2 += st["rays"]; bad2 += st["mismatch"]
            for mode in (1, 3):
                st = T.run(lib, flat, far, mode=mode, max_dist=md, sah=sah, depth=depth)[3]
                tot2 += st["rays"]; bad2 += st["mismatch"]
    print("depth sweep", name, "mismatches so far", bad2, flush=True)
print("DEPTH SWEEP rays", tot2, "mismatches", bad2)
