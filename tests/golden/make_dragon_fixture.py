"""Writes tests/golden/dragon_standin.npz: the flat arrays of the C3 stand-in scene (procedural torus knot, 87 040 triangles,
harness-assigned mirror material) as produced by the product's host generator (cgrt_host_scene_dragon_standin).
The CPU arms of bench.py (`--impl reference`, cpu_baseline) and the full-size parity tests read the scene from this file, so
they never load libcgrt_b200.so; tests/test_host_logic.py checks that the generator still reproduces the file bit for bit.
    python tests/golden/make_dragon_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as ge  # noqa: E402

ge.build()
d = ge.load_package().capi.dragon_standin()
np.savez_compressed(os.path.join(HERE, "dragon_standin.npz"), vcount=d.vcount, tcount=d.tcount, vertices=d.vertices,
                    triangles=d.triangles, materials=d.materials, spheres=d.spheres, lights=d.lights)
print("triangles", int(d.tcount.sum()), "vertices", int(d.vcount.sum()))
