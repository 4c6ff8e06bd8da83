"""CPU: the standalone restatement (oracle/cgrt_oracle.cpp) must reproduce, bit for bit, what the reference's own code
produced for the committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from oracle/_ref)."""
import numpy as np

from conftest import GOLDEN, bits, hits_equal, same_bits, load_golden
from oracle import bindings as ob


def test_unit_vectors(oracle):
    z = np.load(f"{GOLDEN}/units.npz")
    ar = z["aabb_rays"].view(ob.RAY_DTYPE).reshape(-1)
    hit, t = oracle.ray_aabb(z["aabb_boxes"], ar)
    assert np.array_equal(hit, z["aabb_hit"]) and same_bits(t, z["aabb_t"])
    tr = z["tri_rays"].view(ob.RAY_DTYPE).reshape(-1)
    out = oracle.ray_triangle(z["tri_in"], tr)
    gold = z["tri_out"].view(ob.HIT_DTYPE).reshape(-1)
    assert hits_equal(out, gold)
    assert gold["tri"].sum() > 500  # the vector set really contains hits
    assert same_bits(oracle.triangle_plane(z["tri_in"][:, :9]), z["planes"])
    ph, pt = oracle.ray_plane(z["planes"], tr)
    assert np.array_equal(ph, z["plane_hit"]) and same_bits(pt, z["plane_t"])
    assert np.array_equal(oracle.point_in_triangle(z["pit_in"]), z["pit_out"])
    st, sh, sn = oracle.ray_sphere(z["sph_in"], ar)
    assert np.array_equal(sh, z["sph_hit"]) and same_bits(st, z["sph_t"]) and same_bits(sn, z["sph_n"])


def test_bvh_topology(oracle, golden):
    b = oracle.scene(golden.flat, golden.lights).bvh()
    meta, aabb = b.nodes()
    assert np.array_equal(meta, golden.z["node_meta"])
    assert same_bits(aabb, golden.z["node_aabb"])
    assert b.num_levels() == int(golden.z["num_levels"])
    leaves = np.nonzero(meta[:, 0])[0]
    order = np.concatenate([b.leaf_triangles(i, meta[i, 4]) for i in leaves])
    assert np.array_equal(order, golden.z["leaf_tris"])


def test_report_structure():
    """weak pins the reference publishes (report.pdf Table 2): Cornell 32 triangles / 8 levels, monkey 968 / 11 levels"""
    from conftest import load_golden
    c, m = load_golden("cornell"), load_golden("monkey")
    assert c.flat.n_triangles == 32 and int(c.z["num_levels"]) == 8 and len(c.flat.vcount) == 8
    assert m.flat.n_triangles == 968 and int(m.z["num_levels"]) == 11 and int(m.flat.vcount.sum()) == 1968
    d = load_golden("dodge")
    assert d.flat.n_triangles == 16311 and len(d.flat.vcount) == 11 and int(d.flat.vcount.sum()) == 48921


def test_closest_hits_and_counts(oracle, golden):
    b = oracle.scene(golden.flat, golden.lights).bvh()
    hits, counts = b.intersect(golden.rays, counts=True)
    assert hits_equal(hits, golden.hits)
    assert np.array_equal(counts, golden.counts)


def test_bvh_equals_brute_force(oracle, golden):
    """the reference's own cross-check: BVH closest hit == intersectRayWithShape(Mesh) over all meshes"""
    sc = oracle.scene(golden.flat, golden.lights)
    hb = sc.intersect_brute(golden.rays[:3000])
    g = golden.hits[:3000]
    finite = np.isfinite(golden.rays["d"][:3000]).all(1)
    # t agrees everywhere; ids may differ only on exact ties (first-visited wins, visiting orders differ)
    assert same_bits(hb["t"][finite], g["t"][finite])
    assert (hb["tri"][finite] == g["tri"][finite]).mean() > 0.995


def test_rendered_frames(oracle, golden):
    b = oracle.scene(golden.flat, golden.lights).bvh()
    cam = ob.default_camera(golden.W, golden.H)
    for L in (1, 2, 5):
        img, cnt = b.render(cam, golden.W, golden.H, trace_limit=L, duplicate_shading=True)
        assert same_bits(img, golden.z[f"img_L{L}"]), f"trace limit {L}"
        got = [cnt[k] for k in ("primary", "primary_hit", "shadow", "bounce", "box_tests", "tri_tests")]
        assert got == golden.z[f"cnt_L{L}"].tolist()


def test_primary_rays(oracle, golden):
    cam = ob.default_camera(golden.W, golden.H)
    rays = oracle.generate_rays(cam, golden.W, golden.H)
    n = golden.W * golden.H
    assert np.array_equal(bits(rays.view(np.float32)), bits(golden.rays[:n].view(np.float32)))


def test_bloom_restatement_against_numpy(oracle):
    """orc_bloom (main.cpp:586-628 restated in C++) against an independent numpy transcription of the same in-place recurrence."""
    rng = np.random.default_rng(0)
    f = (rng.random((26, 31, 3)) * 0.8).astype(np.float32)
    got = oracle.bloom(f)
    H, W = f.shape[:2]
    m = np.where((f[..., 0] + f[..., 1] + f[..., 2] > 1)[..., None], f, 0).astype(np.float32)[::-1].copy()  # rows in ray space
    for y in range(H):
        for x in range(W):
            acc, cnt = m[y, x].copy(), 1
            for i in range(-10, 11):
                if y + i < 0 or y + i > H - 1:
                    continue
                for j in range(-10, 11):
                    if (i == 0 and j == 0) or x + j < 0 or x + j > W - 1:
                        continue
                    acc = (acc + m[y + i, x + j]).astype(np.float32)
                    cnt += 1
            m[y, x] = (acc / np.float32(cnt)).astype(np.float32)
    want = (m[::-1] + f).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_spherical_light_restatement_properties(oracle):
    """Soft shadows in the oracle (main.cpp:168-218): 200 sample rays per hit, a zero-radius light gives all-or-nothing factors,
    a finite one a penumbra whose mean lies between them; fixed seed -> same frame."""
    g = load_golden("cornell")
    none = np.zeros((0, 6), np.float32)
    sc = oracle.scene(g.flat, none)
    W = H = 48
    cam = ob.default_camera(W, H)
    sc.set_spherical_lights(np.array([[0, 0.45, 0, 0.0, 1, 1, 1]], np.float32), seed=1)
    hard, cnt = sc.bvh().render(cam, W, H, trace_limit=1, nthreads=1)
    assert cnt["shadow"] == 200 * cnt["primary_hit"]
    sc.set_spherical_lights(np.array([[0, 0.45, 0, 0.15, 1, 1, 1]], np.float32), seed=1)
    soft, _ = sc.bvh().render(cam, W, H, trace_limit=1, nthreads=1)
    again, _ = sc.bvh().render(cam, W, H, trace_limit=1, nthreads=1)
    assert np.array_equal(soft, again)
    changed = np.abs(soft - hard).sum(axis=2) > 1e-4
    assert 10 < changed.sum() < 0.6 * W * H  # a penumbra appears, most pixels keep their hard-shadow value
