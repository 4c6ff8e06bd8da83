"""CPU: restatement vs the reference's own TUs (oracle/_ref) on fresh seeded inputs, including the reference CONSTRUCTOR
(mode 0) and edge cases (single triangle, two triangles, many meshes per leaf, empty scene, spheres)."""
import numpy as np
import pytest

from conftest import hits_equal, same_bits
from oracle import bindings as ob


@pytest.mark.parametrize("ntri,nmesh,scale,seed", [(1, 1, 0.5, 4), (2, 1, 0.5, 5), (3, 1, 0.4, 6), (7, 3, 0.5, 7),
                                                   (500, 1, 0.1, 8), (900, 40, 0.2, 9), (2500, 2, 0.05, 10)])
def test_tree_hits_images(reflib, oracle, ntri, nmesh, scale, seed):
    flat = ob.random_soup(ntri, seed=seed, scale=scale, n_meshes=nmesh)
    lights = np.array([[0.0, 0.9, 0.0, 1, 1, 1], [-1, 1, -1, 0.5, 0.5, 0.5]], np.float32)
    rb = reflib.scene(flat, lights).bvh(mode=0)  # the reference constructor itself
    rf = reflib.scene(flat, lights).bvh(mode=1)  # range-based fill of reference Node structs
    ob_ = oracle.scene(flat, lights).bvh()
    m0, a0 = rb.nodes()
    for other in (rf, ob_):
        m1, a1 = other.nodes()
        assert np.array_equal(m0, m1) and same_bits(a0, a1)
        for i in np.nonzero(m0[:, 0])[0]:
            assert np.array_equal(rb.leaf_triangles(i, m0[i, 4]), other.leaf_triangles(i, m1[i, 4]))
    rays = ob.random_rays(5000, seed=seed + 100)
    rays["t"][::7] = np.float32(0.7)
    h0, c0 = rb.intersect(rays, counts=True)
    h1, c1 = ob_.intersect(rays, counts=True)
    assert hits_equal(h1, h0) and np.array_equal(c0, c1)
    cam = ob.default_camera(64, 48)
    i0, k0 = rb.render(cam, 64, 48, trace_limit=3, duplicate_shading=True)
    i1, k1 = ob_.render(cam, 64, 48, trace_limit=3, duplicate_shading=False)
    assert same_bits(i0, i1)
    assert all(k0[k] == k1[k] for k in ("primary", "primary_hit", "shadow", "bounce"))


def test_empty_scene_and_spheres(reflib, oracle):
    empty = ob.FlatScene(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 6)), np.zeros((0, 3)), np.zeros((0, 8)),
                         spheres=[[3, -2, 10.2, 1, .8, .2, .2, 0, 0, 0, 1, 1], [-2, 2, 4, 2, .6, .8, .2, 0, 0, 0, 1, 1],
                                  [0, 0, 6, .75, .2, .2, .8, 0, 0, 0, 1, 1]])  # the Spheres preset, src/scene.cpp:51-56
    rays = ob.random_rays(4000, seed=1)
    rays["o"][:, 2] -= 3
    for lib in (reflib, oracle):
        b = lib.scene(empty).bvh()
        assert b.num_nodes() == 0
    h0 = reflib.scene(empty).bvh(mode=0).intersect(rays)
    h1 = oracle.scene(empty).bvh().intersect(rays)
    assert same_bits(h0["t"], h1["t"]) and same_bits(h0["n"], h1["n"])
    assert (h0["t"] < 1e30).sum() > 50
