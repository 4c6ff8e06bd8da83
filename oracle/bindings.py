"""TEST INFRASTRUCTURE: ctypes bindings for the two CPU checkers in oracle/.

  RefLib     -> oracle/_ref/libcgrt_ref.so : the reference's own TUs compiled verbatim (oracle/ref_harness.cpp)
  OracleLib  -> oracle/liboracle.so        : the standalone restatement (oracle/cgrt_oracle.cpp)

Both expose the same call surface (prefix `ref_` / `orc_`), so tests can run the same checks against either.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

RAY_DTYPE = np.dtype([("o", "f4", 3), ("t", "f4"), ("d", "f4", 3), ("pad", "f4")])
HIT_DTYPE = np.dtype([("t", "f4"), ("tri", "i4"), ("alpha", "f4"), ("beta", "f4"), ("gamma", "f4"), ("n", "f4", 3)])


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_meshes", C.c_int32),
        ("mesh_vertex_count", C.POINTER(C.c_int32)),
        ("mesh_triangle_count", C.POINTER(C.c_int32)),
        ("vertices", C.POINTER(C.c_float)),
        ("triangles", C.POINTER(C.c_uint32)),
        ("materials", C.POINTER(C.c_float)),
        ("n_spheres", C.c_int32),
        ("spheres", C.POINTER(C.c_float)),
    ]


class CameraDesc(C.Structure):
    _fields_ = [("fovy", C.c_float), ("aspect", C.c_float), ("dist", C.c_float),
                ("lookAt", C.c_float * 3), ("euler", C.c_float * 3)]


def default_camera(W, H):
    """Reference preset: fovy 50 deg, distance 3, lookAt 0, Euler (20,20,0) deg (main.cpp:730-731); aspect W/H (window.cpp:334-337)."""
    c = CameraDesc()
    c.fovy = np.float32(50.0) * np.float32(0.01745329251994329576923690768489)  # glm::radians
    c.aspect = np.float32(W) / np.float32(H)
    c.dist = 3.0
    c.lookAt[:] = [0.0, 0.0, 0.0]
    rad = np.float32(20.0) * np.float32(0.01745329251994329576923690768489)
    c.euler[:] = [rad, rad, 0.0]
    return c


class FlatScene:
    """Flat scene arrays (the layout of cgrt_scene_desc in include/cgrt_b200.h)."""

    def __init__(self, vcount, tcount, vertices, triangles, materials, spheres=None):
        self.vcount = np.ascontiguousarray(vcount, dtype=np.int32)
        self.tcount = np.ascontiguousarray(tcount, dtype=np.int32)
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 6)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
        self.materials = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 8)
        self.spheres = np.ascontiguousarray(spheres if spheres is not None else np.zeros((0, 12)), dtype=np.float32).reshape(-1, 12)
        assert self.vertices.shape[0] == int(self.vcount.sum())
        assert self.triangles.shape[0] == int(self.tcount.sum())
        assert self.materials.shape[0] == len(self.vcount) == len(self.tcount)

    @property
    def n_triangles(self):
        return int(self.tcount.sum())

    def desc(self):
        d = SceneDesc()
        d.n_meshes = len(self.vcount)
        d.mesh_vertex_count = self.vcount.ctypes.data_as(C.POINTER(C.c_int32))
        d.mesh_triangle_count = self.tcount.ctypes.data_as(C.POINTER(C.c_int32))
        d.vertices = self.vertices.ctypes.data_as(C.POINTER(C.c_float))
        d.triangles = self.triangles.ctypes.data_as(C.POINTER(C.c_uint32))
        d.materials = self.materials.ctypes.data_as(C.POINTER(C.c_float))
        d.n_spheres = self.spheres.shape[0]
        d.spheres = self.spheres.ctypes.data_as(C.POINTER(C.c_float))
        return d

    def global_positions(self):
        """[T][3][3] vertex positions per global triangle id (mesh order, then triangle order)."""
        out = np.empty((self.n_triangles, 3, 3), np.float32)
        vo = to = 0
        for nv, nt in zip(self.vcount, self.tcount):
            tri = self.triangles[to:to + nt].astype(np.int64) + vo
            out[to:to + nt] = self.vertices[tri, :3]
            vo += nv
            to += nt
        return out

    def canonical_ids(self):
        """id -> smallest id with bit-identical vertex positions (duplicate triangles are indistinguishable to a ray)."""
        pos = self.global_positions().reshape(self.n_triangles, 9).view(np.uint32)
        _, first, inv = np.unique(pos, axis=0, return_index=True, return_inverse=True)
        # np.unique returns first occurrence index for each unique row
        return first[inv.reshape(-1)].astype(np.int32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class _CpuChecker:
    """Common wrapper; `p` is the symbol prefix."""

    def __init__(self, path, p):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.p = p
        L = self.lib
        f = lambda name: getattr(L, p + name)
        f("scene_create").restype = C.c_void_p
        f("scene_create").argtypes = [C.POINTER(SceneDesc)]
        f("scene_destroy").argtypes = [C.c_void_p]
        f("scene_set_lights").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        f("bvh_create").restype = C.c_void_p
        f("bvh_create").argtypes = [C.c_void_p, C.c_int, C.c_int]
        f("bvh_destroy").argtypes = [C.c_void_p]
        f("bvh_num_levels").argtypes = [C.c_void_p]
        f("bvh_num_nodes").argtypes = [C.c_void_p]
        f("bvh_export_nodes").argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        f("bvh_leaf_triangles").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int]
        f("intersect").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int]
        f("intersect_brute").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float), C.c_int]
        f("ray_aabb").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
        f("ray_triangle").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float)]
        f("ray_plane").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
        f("triangle_plane").argtypes = [C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float)]
        f("point_in_triangle").argtypes = [C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_uint8)]
        f("ray_sphere").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float)]
        f("generate_rays").argtypes = [C.POINTER(CameraDesc), C.c_int, C.c_int, C.POINTER(C.c_float)]
        f("render").argtypes = [C.c_void_p, C.POINTER(CameraDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int]
        f("max_threads").restype = C.c_int

    def fn(self, name):
        return getattr(self.lib, self.p + name)

    # -- scene / bvh ---------------------------------------------------------------------------------------------------
    def scene(self, flat, lights=None):
        return CpuScene(self, flat, lights)

    # -- unit functions ------------------------------------------------------------------------------------------------
    def ray_aabb(self, boxes, rays):
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        n = boxes.shape[0]
        hit = np.zeros(n, np.uint8)
        t = np.zeros(n, np.float32)
        self.fn("ray_aabb")(_fp(boxes), _fp(rays.view(np.float32)), n, hit.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(t))
        return hit.astype(bool), t

    def ray_triangle(self, tris, rays):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 18)
        n = tris.shape[0]
        out = np.zeros(n, HIT_DTYPE)
        self.fn("ray_triangle")(_fp(tris), _fp(rays.view(np.float32)), n, _fp(out.view(np.float32)))
        return out

    def ray_plane(self, planes, rays):
        planes = np.ascontiguousarray(planes, np.float32).reshape(-1, 4)
        n = planes.shape[0]
        hit = np.zeros(n, np.uint8)
        t = np.zeros(n, np.float32)
        self.fn("ray_plane")(_fp(planes), _fp(rays.view(np.float32)), n, hit.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(t))
        return hit.astype(bool), t

    def triangle_plane(self, tris):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        out = np.zeros((tris.shape[0], 4), np.float32)
        self.fn("triangle_plane")(_fp(tris), tris.shape[0], _fp(out))
        return out

    def point_in_triangle(self, v0v1v2np):
        a = np.ascontiguousarray(v0v1v2np, np.float32).reshape(-1, 15)
        out = np.zeros(a.shape[0], np.uint8)
        self.fn("point_in_triangle")(_fp(a), a.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def ray_sphere(self, spheres, rays):
        s = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4)
        out = np.zeros((s.shape[0], 5), np.float32)
        self.fn("ray_sphere")(_fp(s), _fp(rays.view(np.float32)), s.shape[0], _fp(out))
        return out[:, 0].copy(), out[:, 1].copy().view(np.int32).astype(bool), out[:, 2:5].copy()

    def generate_rays(self, cam, W, H):
        rays = np.zeros(W * H, RAY_DTYPE)
        self.fn("generate_rays")(C.byref(cam), W, H, _fp(rays.view(np.float32)))
        return rays

    def max_threads(self):
        return int(self.fn("max_threads")())

    def bloom(self, rgb):
        """bloomEffect (main.cpp:586-628) on a frame in Screen layout; restated in the oracle library only (main.cpp cannot be built)."""
        lib = OracleLib().lib
        lib.orc_bloom.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float)]
        rgb = np.ascontiguousarray(rgb, np.float32)
        H, W = rgb.shape[:2]
        out = np.zeros_like(rgb)
        lib.orc_bloom(_fp(rgb), W, H, _fp(out))
        return out


class CpuScene:
    def __init__(self, lib, flat, lights=None):
        self.lib = lib
        self.flat = flat
        d = flat.desc()
        self.h = C.c_void_p(lib.fn("scene_create")(C.byref(d)))
        self._bvhs = []
        if lights is not None:
            self.set_lights(lights)

    def set_lights(self, lights):
        l = np.ascontiguousarray(lights, np.float32).reshape(-1, 6)
        self.lib.fn("scene_set_lights")(self.h, l.shape[0], _fp(l))

    def set_spherical_lights(self, lights, seed=1):
        """[n][7] = position, radius, colour (src/scene.h:47-51); restated in the oracle library only (main.cpp:168-218)."""
        l = np.ascontiguousarray(lights, np.float32).reshape(-1, 7)
        f = self.lib.lib.orc_scene_set_spherical_lights
        f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        f(self.h, l.shape[0], _fp(l))
        self.lib.lib.orc_set_soft_shadow_seed.argtypes = [C.c_uint]
        self.lib.lib.orc_set_soft_shadow_seed(seed)

    def bvh(self, mode=1, max_depth=12):
        return CpuBVH(self, mode, max_depth)

    def intersect_brute(self, rays, nthreads=0):
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        self.lib.fn("intersect_brute")(self.h, _fp(rays.view(np.float32)), rays.shape[0], _fp(hits.view(np.float32)), nthreads)
        return hits

    def close(self):
        if self.h:
            self.lib.fn("scene_destroy")(self.h)
            self.h = None


class CpuBVH:
    def __init__(self, scene, mode, max_depth):
        self.scene = scene
        self.lib = scene.lib
        self.h = C.c_void_p(self.lib.fn("bvh_create")(scene.h, mode, max_depth))

    def num_levels(self):
        return int(self.lib.fn("bvh_num_levels")(self.h))

    def num_nodes(self):
        return int(self.lib.fn("bvh_num_nodes")(self.h))

    def nodes(self):
        n = self.num_nodes()
        meta = np.zeros((n, 5), np.int32)
        aabb = np.zeros((n, 6), np.float32)
        if n:
            self.lib.fn("bvh_export_nodes")(self.h, meta.ctypes.data_as(C.POINTER(C.c_int32)), _fp(aabb))
        return meta, aabb

    def leaf_triangles(self, node, count):
        out = np.zeros(max(count, 1), np.int32)
        k = self.lib.fn("bvh_leaf_triangles")(self.h, node, out.ctypes.data_as(C.POINTER(C.c_int32)), count)
        assert k == count, (k, count)
        return out[:count]

    def intersect(self, rays, counts=False, nthreads=0):
        n = rays.shape[0]
        hits = np.zeros(n, HIT_DTYPE)
        cnt = np.zeros((n, 2), np.uint32) if counts else None
        self.lib.fn("intersect")(self.h, _fp(rays.view(np.float32)), n, _fp(hits.view(np.float32)),
                                 cnt.ctypes.data_as(C.POINTER(C.c_uint32)) if counts else None, nthreads)
        return (hits, cnt) if counts else hits

    def render(self, cam, W, H, trace_limit=2, duplicate_shading=False, y0=0, y1=None, nthreads=0):
        """Returns (rgb[H,W,3] in Screen layout, counters dict). Rows outside [y0,y1) stay 0."""
        y1 = H if y1 is None else y1
        rgb = np.zeros((H, W, 3), np.float32)
        cnt = np.zeros(6, np.uint64)
        self.lib.fn("render")(self.h, C.byref(cam), W, H, trace_limit, int(duplicate_shading), _fp(rgb),
                              cnt.ctypes.data_as(C.POINTER(C.c_uint64)), y0, y1, nthreads)
        keys = ["primary", "primary_hit", "shadow", "bounce", "box_tests", "tri_tests"]
        return rgb, dict(zip(keys, [int(x) for x in cnt]))

    def close(self):
        if self.h:
            self.lib.fn("bvh_destroy")(self.h)
            self.h = None


def ref_lib_path():
    return os.path.join(HERE, "_ref", "libcgrt_ref.so")


def oracle_lib_path():
    return os.path.join(HERE, "liboracle.so")


_cache = {}


def RefLib():
    if "ref" not in _cache:
        _cache["ref"] = _CpuChecker(ref_lib_path(), "ref_")
    return _cache["ref"]


def OracleLib():
    if "orc" not in _cache:
        _cache["orc"] = _CpuChecker(oracle_lib_path(), "orc_")
    return _cache["orc"]


def make_rays(origins, dirs, t=None):
    n = origins.shape[0]
    r = np.zeros(n, RAY_DTYPE)
    r["o"] = origins
    r["d"] = dirs
    r["t"] = np.float32(np.finfo(np.float32).max) if t is None else t
    return r


def random_soup(n_tris, seed=1234, scale=0.01, n_meshes=1, smooth_normals=True):
    """Random-triangle soup in the spirit of SURVEY §8(d) C4: centres U(-1,1)^3, vertices centre + scale*U(-1,1)^3."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n_tris, 1, 3)).astype(np.float32)
    v = (c + np.float32(scale) * rng.uniform(-1, 1, (n_tris, 3, 3)).astype(np.float32)).astype(np.float32)
    if smooth_normals:
        nrm = rng.normal(size=(n_tris, 3, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=2, keepdims=True).astype(np.float32)
    else:
        e1 = v[:, 1] - v[:, 0]
        e2 = v[:, 2] - v[:, 0]
        fn = np.cross(e1, e2)
        fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-30)
        nrm = np.repeat(fn[:, None, :], 3, axis=1)
    verts = np.concatenate([v, nrm.astype(np.float32)], axis=2).reshape(-1, 6)
    per = np.full(n_meshes, n_tris // n_meshes, np.int32)
    per[: n_tris % n_meshes] += 1
    tris = []
    for nt in per:
        tris.append(np.arange(nt * 3, dtype=np.uint32).reshape(-1, 3))
    mats = np.zeros((n_meshes, 8), np.float32)
    mats[:, 0:3] = rng.uniform(0.2, 0.9, (n_meshes, 3))
    mats[:, 3:6] = rng.uniform(0.0, 0.6, (n_meshes, 3))
    mats[:, 6] = rng.uniform(1, 30, n_meshes)
    mats[:, 7] = 1.0
    return FlatScene(per * 3, per, verts, np.concatenate(tris), mats)


def random_rays(n, seed=5678, tmax=None):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return make_rays(o, d.astype(np.float32), tmax)


def dragon_standin_fixture():
    """(FlatScene, lights) of the C3 stand-in scene from tests/golden/dragon_standin.npz (written by
    tests/golden/make_dragon_fixture.py from the product's host generator): lets the CPU arms run without the product library."""
    z = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "dragon_standin.npz"))
    return FlatScene(z["vcount"], z["tcount"], z["vertices"], z["triangles"], z["materials"], z["spheres"]), z["lights"]
