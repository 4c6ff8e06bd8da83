"""CPU fuzz of the speculative traversal's certificates. tests/spec_harness/spec_harness.cpp compiles the PRODUCT's traversal
source (cgrt_device.cuh) for the host and runs, per ray, the literal reference-order traversal (traverseStrict) and the
speculative search + certificate without the exact replay. Contract checked here: whenever the certificate accepts a result,
it is the reference's result in every bit (triangle, distance / shadow flag); the harness's reference-order traversal is itself
tied to the oracle on a sample. Adversarial inputs: walls lying in the faces of their boxes, coplanar grids hit exactly on shared
edges and vertices, duplicated triangles (exact ties), slivers and degenerate triangles, axis-parallel rays with origins on
box faces, rays that start on surfaces, finite ray bounds."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import bindings as ob

HERE = os.path.join(ROOT, "tests", "spec_harness")
FLT_MAX = np.float32(np.finfo(np.float32).max)


@pytest.fixture(scope="module")
def harness():
    so = os.path.join(HERE, "libspec_harness.so")
    csrc = os.path.join(ROOT, "cg-raytracer_b200", "csrc")
    srcs = [os.path.join(HERE, "spec_harness.cpp"), os.path.join(csrc, "bvh_build.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in ("cgrt_device.cuh", "rt_math.cuh", "cgrt_kernels.h", "bvh_build.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cuda_inc = "/usr/local/cuda/include"
        if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
            pytest.skip("CUDA headers not found")
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
               "-Wno-attributes", "-I", cuda_inc, "-I", csrc] + srcs + ["-o", so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.spec_run.restype = C.c_int
    return lib


def run(lib, flat, rays, mode=0, max_dist=None, eps=0.001, sah=True, depth=12):
    d = flat.desc()
    n = rays.shape[0]
    r = np.ascontiguousarray(rays).view(np.float32).reshape(n, 8)
    ex = np.zeros((n, 2), np.int32)
    fa = np.zeros((n, 2), np.int32)
    cert = np.zeros(n, np.uint8)
    st = np.zeros(8, np.int64)
    md = np.ascontiguousarray(max_dist if max_dist is not None else np.zeros(n, np.float32), np.float32)

    def fp(a):
        return a.ctypes.data_as(C.c_void_p)

    rc = lib.spec_run(C.byref(d), C.c_int(depth), C.c_int(1 if sah else 0), fp(r), C.c_int64(n), C.c_int(mode), fp(md),
                      C.c_float(eps), fp(ex), fp(fa), fp(cert), fp(st))
    assert rc == 0
    keys = ("rays", "certified", "deferred", "mismatch", "first", "tree", "always", "wide")
    return ex, fa, cert.astype(bool), dict(zip(keys, st.tolist()))


# ---- scenes ------------------------------------------------------------------------------------------------------------------
def flat_from(meshes):
    """meshes: list of (positions[n,3], triangles[m,3]); flat +z normals (normals do not enter the traversal)"""
    V, T, vc, tc = [], [], [], []
    for P, F in meshes:
        P = np.asarray(P, np.float32)
        nrm = np.zeros_like(P)
        nrm[:, 2] = 1.0
        V.append(np.concatenate([P, nrm], axis=1))
        T.append(np.asarray(F, np.uint32))
        vc.append(len(P))
        tc.append(len(F))
    mats = np.tile(np.array([0.7, 0.7, 0.7, 0.5, 0.5, 0.5, 8.0, 1.0], np.float32), (len(meshes), 1))
    return ob.FlatScene(np.array(vc, np.int32), np.array(tc, np.int32), np.concatenate(V), np.concatenate(T), mats,
                        np.zeros((0, 12), np.float32))


def box_walls(h=0.7):
    q = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * h
    faces = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4)]
    return flat_from([(q[list(f)], [[0, 1, 2], [0, 2, 3]]) for f in faces])


def grid_planes(n=24, duplicate=False):
    """two parallel coplanar triangle grids (every ray through a grid vertex / edge meets shared edges); optionally every
    triangle twice (exact ties)"""
    meshes = []
    xs = np.linspace(-0.8, 0.8, n + 1, dtype=np.float32)
    for z in (0.0, -0.35):
        P = np.array([[x, y, z] for y in xs for x in xs], np.float32)
        F = []
        for j in range(n):
            for i in range(n):
                a = j * (n + 1) + i
                F += [[a, a + 1, a + n + 2], [a, a + n + 2, a + n + 1]]
        if duplicate:
            F = F + F
        meshes.append((P, F))
    return flat_from(meshes), xs


def slivers():
    flat = ob.random_soup(4000, seed=99, scale=0.05, n_meshes=3)
    v = flat.vertices.copy()
    t = flat.triangles
    rng = np.random.default_rng(11)
    for k in range(40):
        a, b, c = t[k]
        v[c, :3] = v[a, :3] + (v[b, :3] - v[a, :3]) * np.float32(0.5) + rng.normal(0, 10.0 ** -float(rng.integers(3, 10)), 3).astype(np.float32)
    a, b, c = t[41]
    v[b, :3] = v[a, :3]
    return ob.FlatScene(flat.vcount, flat.tcount, v, t, flat.materials, flat.spheres)


# ---- rays --------------------------------------------------------------------------------------------------------------------
def mk_rays(o, d, t):
    r = np.zeros(len(o), ob.RAY_DTYPE)
    r["o"] = o
    r["d"] = d
    r["t"] = t
    return r


def ray_mix(flat, seed, n=150000, grid=None):
    rng = np.random.default_rng(seed)
    sets = []
    a = ob.random_rays(n, seed=seed)
    a["t"][::5] = np.float32(0.9)
    sets.append(a)
    # axis-parallel rays, origins snapped to a grid (zero direction components, origins on box faces)
    k = n // 3
    d = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice(np.array([-1, 1], np.float32), k)[:, None]
    o = (np.round(rng.uniform(-1, 1, (k, 3)) * 10) / 10 * 0.7).astype(np.float32)
    sets.append(mk_rays(o, d, np.full(k, FLT_MAX)))
    # rays leaving surfaces (bounce / shadow origins: P + 0.001 d), bounded like a reflection ray or unbounded
    V, T = flat.vertices[:, :3], flat.triangles.astype(np.int64)
    off = np.concatenate([[0], np.cumsum(flat.vcount)[:-1]]).astype(np.int64)
    moff = np.repeat(off, flat.tcount)
    tri = rng.integers(0, len(T), k)
    w = rng.dirichlet([1, 1, 1], k).astype(np.float32)
    P = (V[T[tri, 0] + moff[tri]] * w[:, :1] + V[T[tri, 1] + moff[tri]] * w[:, 1:2] + V[T[tri, 2] + moff[tri]] * w[:, 2:3]).astype(np.float32)
    d = rng.normal(size=(k, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    tb = np.where(rng.random(k) < 0.5, np.float32(1.0), FLT_MAX).astype(np.float32)
    sets.append(mk_rays((P + np.float32(0.001) * d).astype(np.float32), d, tb))
    if grid is not None:  # straight at grid vertices and edge midpoints, from a few origins and from straight above
        xs = np.concatenate([grid, (grid[:-1] + grid[1:]) / 2]).astype(np.float32)
        tx, ty = np.meshgrid(xs, xs)
        tgt = np.stack([tx.ravel(), ty.ravel(), np.zeros(tx.size, np.float32)], axis=1).astype(np.float32)
        for org in ([0.0, 0.0, 2.0], [0.3, -0.2, 1.5], [0.8, 0.8, 0.5]):
            o = np.tile(np.array(org, np.float32), (len(tgt), 1))
            sets.append(mk_rays(o, (tgt - o).astype(np.float32), np.full(len(tgt), FLT_MAX)))
        o = (tgt + np.array([0, 0, 1.0], np.float32)).astype(np.float32)
        sets.append(mk_rays(o, np.tile(np.array([0, 0, -1.0], np.float32), (len(tgt), 1)), np.full(len(tgt), FLT_MAX)))
    return np.concatenate(sets)


SCENES = ["soup", "soup_meshes", "boxes", "grid", "grid_dup", "slivers", "cornell", "monkey", "dodge"]


def scene(kind):
    grid = None
    if kind == "soup":
        flat = ob.random_soup(20000, seed=77, scale=0.04)
    elif kind == "soup_meshes":
        flat = ob.random_soup(4000, seed=78, scale=0.15, n_meshes=40)
    elif kind == "boxes":
        flat = box_walls()
    elif kind == "grid":
        flat, grid = grid_planes()
    elif kind == "grid_dup":
        flat, grid = grid_planes(duplicate=True)
    elif kind == "slivers":
        flat = slivers()
    else:
        flat = load_golden(kind).flat
    return flat, grid


@pytest.mark.parametrize("sah", [True, False], ids=["sah", "ref"])
@pytest.mark.parametrize("kind", SCENES)
def test_certified_results_are_the_references(harness, kind, sah):
    flat, grid = scene(kind)
    rays = ray_mix(flat, seed=len(kind) + 7 * sah, grid=grid)
    ex, fa, cert, st = run(harness, flat, rays, mode=0, sah=sah)
    assert st["tree"] == 1
    assert st["mismatch"] == 0, (st, rays[st["first"]], ex[st["first"]], fa[st["first"]])
    assert np.array_equal(ex[cert], fa[cert])
    hits = (ex[:, 0] >= 0).mean()
    share = st["deferred"] / st["rays"]
    print(f"[{kind}/{'sah' if sah else 'ref'}] rays {st['rays']} hit {hits:.3f} deferred {share:.5f} always {st['always']}")
    assert hits > 0.02
    if kind == "grid_dup":
        assert share > 0.01  # duplicated triangles are exact ties: the certificate must refuse them
    # any hit: unbounded range, random ranges, and ranges around the hit distances (the predicate's boundary)
    rng = np.random.default_rng(5)
    far = rays.copy()
    far["t"] = FLT_MAX
    exf, _, _, _ = run(harness, flat, far, mode=0, sah=sah)
    t_hit = np.where(exf[:, 0] >= 0, exf[:, 1].view(np.float32), np.float32(1.0)).astype(np.float32)
    near = (t_hit * rng.choice(np.array([0.999, 1.0, 1.001], np.float32), len(rays))).astype(np.float32) + np.float32(0.001)
    for md in (np.full(len(rays), np.inf, np.float32), rng.uniform(0, 2, len(rays)).astype(np.float32), near):
        _, _, _, sa = run(harness, flat, far, mode=1, max_dist=md, sah=sah)
        assert sa["mismatch"] == 0, sa


@pytest.mark.parametrize("kind", SCENES)
def test_exact_replay_traversal_is_the_literal_one(harness, kind):
    """the traversal that handles deferred rays (filtered reference box decisions, culling sub-trees in the reference leaves)
    gives the literal reference-order traversal's result for EVERY ray, closest and any hit"""
    flat, grid = scene(kind)
    rays = ray_mix(flat, seed=40 + len(kind), n=90000, grid=grid)
    ex, fa, _, st = run(harness, flat, rays, mode=2)
    assert st["mismatch"] == 0, (st, rays[st["first"]], ex[st["first"]], fa[st["first"]])
    far = rays.copy()
    far["t"] = FLT_MAX
    md = np.random.default_rng(9).uniform(0, 2, len(rays)).astype(np.float32)
    _, _, _, sa = run(harness, flat, far, mode=3, max_dist=md)
    assert sa["mismatch"] == 0, sa


def test_harness_reference_order_traversal_is_the_oracles(harness):
    """ties the harness's ground truth (the product's traverseStrict compiled for the host) to the CPU oracle"""
    for kind in ("soup_meshes", "boxes", "cornell"):
        flat, grid = scene(kind)
        rays = ray_mix(flat, seed=3, n=30000, grid=grid)
        ex, _, _, _ = run(harness, flat, rays)
        g = ob.OracleLib().scene(flat).bvh().intersect(rays)
        canon = flat.canonical_ids()
        ids = np.where(ex[:, 0] >= 0, canon[np.maximum(ex[:, 0], 0)], -1)
        assert np.array_equal(ids, g["tri"])
        hit = g["tri"] >= 0
        assert np.array_equal(ex[hit, 1], g["t"][hit].view(np.int32))


@pytest.mark.parametrize("kind", ["boxes", "grid", "soup_meshes", "cornell"])
def test_special_float_values(harness, kind):
    """garbage in, reference behaviour out: NaN / inf / denormal / zero components in origins and directions, negative, NaN
    and zero ray bounds. The reference has defined (if odd) behaviour for all of it - e.g. its in-plane shortcut accepts t = 0
    whatever ray.t is - and both the certified speculative results and the exact replay traversal must reproduce it."""
    rng = np.random.default_rng(7)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-38, -1e-38, 1e38, -1e38, 1e-45, 3.4e38, 1.0, -1.0, 0.5], np.float32)
    flat, _ = scene(kind)
    n = 80000
    rays = ob.random_rays(n, seed=5)
    for field in ("o", "d"):
        m = rng.random((n, 3)) < 0.25
        rays[field] = np.where(m, special[rng.integers(0, len(special), (n, 3))], rays[field])
    rays["t"] = np.where(rng.random(n) < 0.4, special[rng.integers(0, len(special), n)], rays["t"]).astype(np.float32)
    md = special[rng.integers(0, len(special), n)].astype(np.float32)
    for sah in (True, False):
        for mode in (0, 2):
            ex, fa, _, st = run(harness, flat, rays, mode=mode, sah=sah)
            assert st["mismatch"] == 0, (mode, st, rays[st["first"]], ex[st["first"]], fa[st["first"]])
        for mode in (1, 3):
            _, _, _, st = run(harness, flat, rays, mode=mode, max_dist=md, sah=sah)
            assert st["mismatch"] == 0, (mode, st)


def test_first_stage_slab_test_is_sound(harness):
    """The certificates' first stage (reciprocal-direction slab test with margins, cgrt_device.cuh slabFastHit) may only claim
    what the reference's own slabTest confirms: a hit of the unbounded ray at a distance <= the claimed bound. 40 M generated
    (box, ray) pairs weighted towards the hard cases; the stage must also decide most ordinary cases (otherwise it is useless)."""
    import ctypes as C
    harness.spec_slab_fast_check.restype = C.c_int
    out = (C.c_int64 * 4)()
    tot = [0, 0, 0, 0]
    for seed in (1, 2, 3, 4):
        assert harness.spec_slab_fast_check(C.c_int64(10_000_000), C.c_uint64(seed), out) == 0
        tot = [a + int(b) for a, b in zip(tot, out)]
    assert tot[2] == 0, f"{tot[2]} claims of the fast stage contradicted by the reference's slabTest (of {tot[1]} claims)"
    # (looseness: the bound is relative to the largest of the six slab distances, so far-away slabs widen it - by design)
    assert tot[1] > tot[0] // 50 and tot[3] < tot[1] // 5, tot
