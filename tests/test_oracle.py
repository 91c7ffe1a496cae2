"""The oracle (oracle/spica_oracle.cc) against (1) the reference's own live known-answer tests,
(2) golden vectors produced by the unmodified reference compiled here (tests/golden/make_golden.py)
and (3) - when oracle/_ref and /root/reference are present - the reference run live."""
import math
import os

import numpy as np
import pytest

from oracle import binding as ob
from spica_b200 import scenes

REF_DATA = "/root/reference/tests/data"


def test_kat_triangle_intersection():
    # reference tests/test_geometry.cc:50-77
    tri = [1, 0, 0, 0, 0, 0, 0, 1, 0]
    ok, d, _ = ob.ray_init([0, 0, -1], [1, 1, 2])          # (1,1,1) - (0,0,-1)
    assert ok
    hit, t, u, v = ob.triangle_intersect(tri, [0, 0, -1], d)
    assert hit and t == math.sqrt(6.0) / 2.0                 # EXPECT_EQ: exact
    for o in ([-0.1, -0.1, 1.0], [0.6, 0.6, 1.0], [-0.1, 1.1, 1.0], [1.1, -0.1, 1.0]):
        ok, d, _ = ob.ray_init(o, [0, 0, -1])
        assert not ob.triangle_intersect(tri, o, d)[0]


def test_kat_bounds_intersection():
    # reference tests/test_geometry.cc:119-125
    ok, d, inv = ob.ray_init([0.5, 0.5, -1.0], [0, 0, 1])
    hit, tn, tf = ob.bounds_intersect([0, 0, 0], [1, 1, 1], [0.5, 0.5, -1.0], inv)
    assert hit and tn == pytest.approx(1.0, abs=0, rel=4e-16) and tf == pytest.approx(2.0, abs=0, rel=4e-16)


def test_kat_ray():
    # reference tests/test_ray.cc:13-23
    ok, d, inv = ob.ray_init([1, 2, 3], [1, 0, 0])
    assert ok and d == [1.0, 0.0, 0.0] and inv == [1.0, 1.0e32, 1.0e32]
    assert not ob.ray_init([1, 2, 3], [0, 0, 0])[0]          # the reference aborts (ASSERT_DEATH)


def test_ray_normalisation_is_reciprocal_multiply():
    # core/vector3d_detail.h:134-137: v *= (1.0 / |v|), not v / |v|
    d = np.array([0.3, -1.7, 2.9])
    ok, dn, _ = ob.ray_init([0, 0, 0], d)
    s = 1.0 / math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    assert dn == [d[0] * s, d[1] * s, d[2] * s]


@pytest.mark.parametrize("name", ["golden_torus", "golden_cube"])
def test_golden_bvh_topology_and_hits(name, request):
    g = request.getfixturevalue(name)
    nodes = ob.bvh_build(g["tris"])
    ref = g["bvh_nodes"]
    assert len(nodes) == len(ref) == 2 * len(g["tris"]) - 1
    for k in ("lo", "hi", "left", "right", "prim", "axis"):
        assert np.array_equal(nodes[k], ref[k]), k
    prim, t, _, _ = ob.trace_closest(nodes, g["tris"], g["rays"])
    assert np.array_equal(prim, g["prim"]) and np.array_equal(t, g["t"])     # bit exact
    occ = ob.trace_any(nodes, g["tris"], g["any_rays"])
    assert np.array_equal(occ, g["occluded"])
    prim64, t64, _, _ = ob.trace_closest(nodes, g["tris"], g["rays64"])
    assert np.array_equal(prim64, g["prim64"]) and np.array_equal(t64, g["t64"])


def test_golden_f64_vertices(golden_f64verts):
    g = golden_f64verts
    nodes = ob.bvh_build(g["tris"])
    assert np.array_equal(nodes["prim"], g["bvh_nodes"]["prim"])
    prim, t, _, _ = ob.trace_closest(nodes, g["tris"], g["rays"])
    assert np.array_equal(prim, g["prim"]) and np.array_equal(t, g["t"])


def test_bvh_matches_bruteforce(golden_torus):
    # the method of the reference's (stale) tests/test_trimesh.cc:153-205
    g = golden_torus
    nodes = ob.bvh_build(g["tris"])
    rays = g["rays"][:2000]
    p0, t0, _, _ = ob.trace_closest(nodes, g["tris"], rays)
    p1, t1 = ob.trace_bruteforce(g["tris"], rays)
    assert np.array_equal(t0, t1)
    assert np.array_equal(p0, p1)


def test_empty_and_single():
    assert len(ob.bvh_build(np.zeros((0, 9)))) == 0
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float64)
    nodes = ob.bvh_build(tri)
    assert len(nodes) == 1 and nodes["prim"][0] == 0
    rays = np.array([[0.2, 0.2, 1, 0, 0, -1, 0, 1e32], [2, 2, 1, 0, 0, -1, 0, 1e32],
                     [0.2, 0.2, 1, 0, 0, 0, 0, 1e32]], dtype=np.float32)
    prim, t, _, _ = ob.trace_closest(nodes, tri, rays)
    assert prim.tolist() == [0, -1, -1] and t[0] == 1.0


@pytest.mark.skipif(not (ob.have_ref() and os.path.isdir(REF_DATA)), reason="needs oracle/_ref and /root/reference")
@pytest.mark.parametrize("mesh", ["box", "kitten", "bunny"])
def test_live_reference_meshes(mesh):
    # BVH == reference's on the reference's own fixtures; hits bit-exact; QBVH path agrees
    # (the intent of the stale tests/test_scene.cc:96-126)
    ply = os.path.join(REF_DATA, mesh + ".ply")
    v, f = scenes.read_ply(ply)
    tris = scenes.mesh_triangles(v, f)
    rays = scenes.incoherent_rays(20000, v.min(0), v.max(0), seed=5)
    info, prim, t, ref_nodes = ob.ref_raycast(rays, ply=ply, dump_bvh=True)
    nodes = ob.bvh_build(tris)
    for k in ("lo", "hi", "left", "right", "prim", "axis"):
        assert np.array_equal(nodes[k], ref_nodes[k]), k
    p, tt, _, _ = ob.trace_closest(nodes, tris, rays)
    assert np.array_equal(p, prim) and np.array_equal(tt, t)
    _, ps, ts = ob.ref_raycast(rays, ply=ply, simd=1)
    assert np.array_equal(ps, prim)
