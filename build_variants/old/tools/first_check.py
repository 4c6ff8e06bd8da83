"""Ad-hoc GPU bring-up check: GPU path vs the verbatim-reference checker on soups (batch queries + renders)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
capi = pkg.capi
from oracle import bindings as ob

print("devices", capi.device_count())
R = ob.RefLib()
for (ntri, nm, scale) in [(2000, 1, 0.05), (3000, 3, 0.08), (20000, 5, 0.03)]:
    flat = ob.random_soup(ntri, seed=ntri, scale=scale, n_meshes=nm)
    lights = np.array([[0.0, 0.9, 0.0, 1, 1, 1], [-1, 1, -1, 0.5, 0.5, 0.5]], np.float32)
    sc = capi.Scene(flat, lights=lights)
    cs = R.scene(flat, lights); cb = cs.bvh(mode=1)
    m0, a0 = sc.nodes(); m1, a1 = cb.nodes()
    print(ntri, nm, "nodes", sc.num_nodes(), cb.num_nodes(), "meta", np.array_equal(m0, m1), "aabb", np.array_equal(a0.view(np.uint32), a1.view(np.uint32)))
    rays = ob.random_rays(200000, seed=5)
    t0 = time.time(); h, c = sc.intersect(rays, counts=True); t1 = time.time()
    oh, oc = cb.intersect(rays, counts=True); t2 = time.time()
    canon = flat.canonical_ids()
    gid = np.where(h["tri"] >= 0, canon[np.maximum(h["tri"], 0)], h["tri"])
    print("  closest: gpu %.3fs cpu %.3fs hits %d id_match %.6f t_bits %s counts %s bary %s normal %s" % (
        t1 - t0, t2 - t1, (oh["tri"] >= 0).sum(), (gid == oh["tri"]).mean(),
        np.array_equal(h["t"].view(np.uint32), oh["t"].view(np.uint32)), np.array_equal(c, oc),
        np.array_equal(h["alpha"].view(np.uint32), oh["alpha"].view(np.uint32)) and np.array_equal(h["gamma"].view(np.uint32), oh["gamma"].view(np.uint32)),
        np.array_equal(h["n"].view(np.uint32), oh["n"].view(np.uint32))))
    hb = sc.intersect_brute(rays[:20000])
    print("  brute==bvh t", np.array_equal(hb["t"].view(np.uint32), h["t"][:20000].view(np.uint32)))
    for (W, H, L) in [(128, 96, 1), (160, 120, 2), (96, 96, 5)]:
        cam = capi.make_camera(W, H); ocam = ob.default_camera(W, H)
        rgb, st = sc.render(cam, W, H, trace_limit=L)
        orgb, ocnt = cb.render(ocam, W, H, trace_limit=L)
        print("  render %dx%d L%d" % (W, H, L), st, {k: ocnt[k] for k in ("primary", "primary_hit", "shadow", "bounce")},
              "max|d|", np.abs(rgb - orgb).max(), "bit-equal px %.6f" % (rgb.view(np.uint32) == orgb.view(np.uint32)).all(axis=2).mean())
    # multi-rank tiling: union of per-rank renders == single render
    W, H = 200, 120
    cam = capi.make_camera(W, H)
    full, _ = sc.render(cam, W, H, trace_limit=2)
    for world in (2, 3, 8):
        acc = np.full((H, W, 3), np.nan, np.float32)
        for r in range(world):
            sc.render(cam, W, H, trace_limit=2, rank=r, world=world, out=acc)
        print("  world", world, "tiles union == full", np.array_equal(acc.view(np.uint32), full.view(np.uint32)))
    sc.close()
print("OK")
