"""Parity tests proper: the CUDA ray-cast path, called through the C ABI (include/spica_b200.h),
against the oracle and the committed golden vectors of the real reference.

Bar: primitive ids bit-exact; t bit-exact in double (and <= 1e-5 relative through the float32 hit
record, the tolerance BASELINE.json states); any-hit flags equal."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from spica_b200 import capi, scenes

pytestmark = pytest.mark.gpu

T_RTOL = 1e-5   # BASELINE.json north_star: "hit distance t must match within 1e-5 relative"


def _as64(rays32):
    return np.ascontiguousarray(rays32.astype(np.float64))


def _check_closest(ctx, tris, rays, prim_ref, t_ref, max_ties=0):
    hits = ctx.trace_closest(rays)
    ties = np.nonzero(hits["prim"] != prim_ref)[0]
    for i in ties:      # only genuine exact-t ties may differ (own-built tree)
        r = rays[i].astype(np.float64)
        ok, d, _ = ob.ray_init(r[:3], r[3:6])
        hit, t_other, _, _ = ob.triangle_intersect(tris[hits["prim"][i]], r[:3], d, r[7])
        assert hit and t_other == t_ref[i], "ray %d: wrong primitive" % i
    assert len(ties) <= max_ties
    h = prim_ref >= 0
    assert np.allclose(hits["t"][h], t_ref[h], rtol=T_RTOL, atol=0)
    assert (hits["t"][~h] == 0).all()
    # float64 entry point on the same rays: t bit-exact
    h64 = ctx.trace_closest(_as64(rays))
    same = h64["prim"] == prim_ref
    assert same.sum() >= len(rays) - max_ties
    assert np.array_equal(h64["t"][same & h], t_ref[same & h])
    return hits


DEFAULT_VARIANT = 5     # spica_b200/csrc/context.h: opt_variant


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("max_leaf", [1, 3])
def test_golden_torus(gpu_ctx, golden_torus, variant, max_leaf):
    """Every kernel variant against the compiled reference's answers: 0 one thread per ray, 1 persistent,
    2 pooled pre-test, 3 early select, 4 shared-memory stack, 5 three visits per pooled phase (default)."""
    g = golden_torus
    gpu_ctx.set_option("trace_variant", variant)
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build(max_leaf_tris=max_leaf)
    _check_closest(gpu_ctx, g["tris"], g["rays"], g["prim"], g["t"])
    assert np.array_equal(gpu_ctx.trace_any(g["any_rays"]), g["occluded"])
    h64 = gpu_ctx.trace_closest(g["rays64"])
    assert np.array_equal(h64["prim"], g["prim64"]) and np.array_equal(h64["t"], g["t64"])
    assert np.array_equal(gpu_ctx.trace_any(_as64(g["any_rays"])), g["occluded"])
    gpu_ctx.set_option("trace_variant", DEFAULT_VARIANT)


def test_golden_cube_own_and_imported_tree(gpu_ctx, golden_cube):
    g = golden_cube
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build()
    _check_closest(gpu_ctx, g["tris"], g["rays"], g["prim"], g["t"])
    gpu_ctx.import_binary(g["bvh_nodes"])
    _check_closest(gpu_ctx, g["tris"], g["rays"], g["prim"], g["t"])
    assert np.array_equal(gpu_ctx.trace_any(g["any_rays"]), g["occluded"])


def test_golden_f64_vertices(gpu_ctx, golden_f64verts):
    g = golden_f64verts
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build()
    assert gpu_ctx.stats()["tri_format"] == 1
    _check_closest(gpu_ctx, g["tris"], g["rays"], g["prim"], g["t"])


@pytest.mark.parametrize("variant", [2, 3, 4, 5])
@pytest.mark.parametrize("offset", [0.0, 250.0])
def test_transformed_mesh_double_vertices_on_the_pooled_kernels(gpu_ctx, variant, offset):
    """A shape with a toWorld transform has vertices that are not float32 numbers (the hosts transform in double, as
    core/triangle.cc does): 80-byte double records for the exact test, and a float32-rounded copy for the pooled
    pre-test whose error bounds carry the rounding (trace_core.h: triPretestMayHit, rnd).  Rays aimed at mesh vertices
    and edge midpoints sit inside that band; the far offset makes the rounding step large against the triangles."""
    from tests.conftest import transformed_torus_case
    tris, rays, lo, hi = transformed_torus_case(offset)
    nodes = ob.bvh_build(tris)
    p0, t0, _, _ = ob.trace_closest(nodes, tris, rays)
    assert (p0 >= 0).sum() > 20000
    gpu_ctx.set_option("trace_variant", variant)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build()
    assert gpu_ctx.stats()["tri_format"] == 1
    _check_closest(gpu_ctx, tris, rays, p0, t0, max_ties=64)        # aimed at shared vertices / edges: genuine ties
    gpu_ctx.import_binary(nodes.view(capi.IMPORT_NODE))
    _check_closest(gpu_ctx, tris, rays, p0, t0, max_ties=0)
    anyr = scenes.incoherent_rays(30000, lo, hi, seed=13, anyhit=True)
    assert np.array_equal(gpu_ctx.trace_any(anyr), ob.trace_any(nodes, tris, anyr))
    gpu_ctx.set_option("trace_variant", DEFAULT_VARIANT)


@pytest.mark.parametrize("seed", [2, 4, 6])          # (2 and 4 hold rays on which the reference's traversal loses a hit; all six run in the CPU emulation test)
def test_random_triangle_soups_against_the_brute_force(gpu_ctx, seed):
    """tests/test_traversal_emul.py::test_random_triangle_soups_float32_and_double_vertices through the C ABI: slivers, needles,
    touching triangles, float32 and double vertices, rays aimed at vertices and edges.  The bar is the reference's triangle
    test run over ALL triangles (its own traversal loses a hit to its box test now and then on such rays)."""
    from tests.conftest import random_soup_cases
    for tris, rays in random_soup_cases(seed):
        pb, tb = ob.trace_bruteforce(tris, rays)
        gpu_ctx.set_triangles(tris)
        for builder in (0, 2):
            gpu_ctx.build(builder=builder, max_leaf_tris=1 + (seed + builder) % 3)
            _check_closest(gpu_ctx, tris, rays, pb, tb, max_ties=len(rays))      # ties: verified one by one in _check_closest
        occ = gpu_ctx.trace_any(rays)
        assert np.array_equal(occ, (pb >= 0).astype(np.uint8))


def test_medium_torus_vs_oracle_with_import(gpu_ctx):
    v, f = scenes.torus_mesh(200, 100)
    tris = scenes.mesh_triangles(v, f)
    rays = np.concatenate([scenes.incoherent_rays(60000, v.min(0), v.max(0), seed=3), scenes.primary_rays(128, 128)], 0)
    nodes = ob.bvh_build(tris)
    p0, t0, _, _ = ob.trace_closest(nodes, tris, rays)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build()
    _check_closest(gpu_ctx, tris, rays, p0, t0, max_ties=32)     # diagonal primary rays: genuine ties
    gpu_ctx.import_binary(nodes.view(capi.IMPORT_NODE))
    _check_closest(gpu_ctx, tris, rays, p0, t0, max_ties=0)      # imported tree: reference tie order
    anyr = scenes.incoherent_rays(40000, v.min(0), v.max(0), seed=4, anyhit=True)
    assert np.array_equal(gpu_ctx.trace_any(anyr), ob.trace_any(nodes, tris, anyr))


def test_edge_cases(gpu_ctx, golden_torus):
    g = golden_torus
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build()
    # empty batch and ragged sizes (not a multiple of the warp / block size)
    assert len(gpu_ctx.trace_closest(np.zeros((0, 8), np.float32))) == 0
    for n in (1, 31, 33, 1000):
        h = gpu_ctx.trace_closest(g["rays"][:n])
        assert np.array_equal(h["prim"], g["prim"][:n])
    # zero-direction rays miss (the reference aborts there); tmax exactly at the hit distance
    rays = g["rays"][:2000].copy()
    hit = g["prim"][:2000] >= 0
    rays[hit, 7] = g["t"][:2000][hit].astype(np.float32)
    rays[::5, 3:6] = 0
    nodes = ob.bvh_build(g["tris"])
    p0, t0, _, _ = ob.trace_closest(nodes, g["tris"], rays)
    h = gpu_ctx.trace_closest(rays)
    assert np.array_equal(h["prim"], p0)
    assert (h["prim"][::5] == -1).all()
    # far-away origins exercise the re-origined culling ray
    rng = np.random.default_rng(1)
    far = np.zeros((4000, 8), np.float32)
    far[:, :3] = rng.uniform(-1, 1, (4000, 3)) * 1e4
    far[:, 3:6] = -far[:, :3] + rng.normal(size=(4000, 3)) * 0.5
    far[:, 7] = 1e32
    p0, t0, _, _ = ob.trace_closest(nodes, g["tris"], far)
    h = gpu_ctx.trace_closest(far)
    assert np.array_equal(h["prim"], p0) and (p0 >= 0).sum() > 100


def test_empty_scene_and_errors(gpu_ctx):
    gpu_ctx.set_triangles(np.zeros((0, 9)))
    gpu_ctx.build()
    rays = scenes.incoherent_rays(100, [-1, -1, -1], [1, 1, 1])
    assert (gpu_ctx.trace_closest(rays)["prim"] == -1).all()
    assert (gpu_ctx.trace_any(rays) == 0).all()
    gpu_ctx.set_triangles(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float64))
    with pytest.raises(capi.SpbError):      # trace before build
        gpu_ctx.trace_closest(rays)
    with pytest.raises(capi.SpbError):
        gpu_ctx.set_option("no_such_option", 1)


def test_device_resident_path_and_counters(gpu_ctx, golden_torus):
    g = golden_torus
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build()
    n = len(g["rays"])
    d_rays = gpu_ctx.dev_alloc(n * 32)
    d_hits = gpu_ctx.dev_alloc(n * 16)
    gpu_ctx.dev_upload(d_rays, g["rays"])
    gpu_ctx.set_option("counters", 1)
    before = gpu_ctx.counters()
    gpu_ctx.trace_closest_dev(d_rays, n, d_hits)
    after = gpu_ctx.counters()
    gpu_ctx.set_option("counters", 0)
    hits = np.empty(n, dtype=capi.HIT)
    gpu_ctx.dev_download(hits, d_hits)
    assert np.array_equal(hits["prim"], g["prim"])
    assert after["kernel_launches"] == before["kernel_launches"] + 1
    assert after["last_kernel_ms"] > 0 and after["node_visits"] > before["node_visits"]
    gpu_ctx.dev_free(d_rays); gpu_ctx.dev_free(d_hits)


def test_full_size_properties_1m_triangles(gpu_ctx):
    """BASELINE config 2 geometry (1M-triangle torus) with size-independent properties:
    a strided subset against the oracle, hit points lie on the mesh, closest t never exceeds the
    any-hit witness, and re-tracing with tmax just below t finds nothing closer."""
    v, f = scenes.torus_mesh(1000, 500)
    tris = scenes.mesh_triangles(v, f)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build()
    n = 1 << 20
    rays = scenes.incoherent_rays(n, v.min(0), v.max(0), seed=2)
    hits = gpu_ctx.trace_closest(rays)
    sub = slice(0, n, 64)
    nodes = ob.bvh_build(tris)
    p0, t0, _, _ = ob.trace_closest(nodes, tris, rays[sub])
    assert np.array_equal(hits["prim"][sub], p0)
    h = p0 >= 0
    assert np.allclose(hits["t"][sub][h], t0[h], rtol=T_RTOL, atol=0)
    # idempotence / minimality: shrinking tmax below the closest hit leaves no hit
    hit = hits["prim"] >= 0
    r2 = rays[hit][:200000].copy()
    r2[:, 7] = hits["t"][hit][:200000] * np.float32(1 - 1e-4)
    assert (gpu_ctx.trace_closest(r2)["prim"] == -1).all()
    assert (gpu_ctx.trace_any(r2) == 0).all()
    # any-hit agrees with closest-hit existence on the same rays
    assert np.array_equal(gpu_ctx.trace_any(rays[:200000]), (hits["prim"][:200000] >= 0).astype(np.uint8))


def _wide_bytes(ctx):
    st, nodes, tris = ctx.export_bvh()
    return st, bytes(nodes), bytes(tris)


@pytest.mark.parametrize("max_leaf", [1, 3])
def test_device_sah_builder_is_the_host_builder_byte_for_byte(gpu_ctx, golden_torus, golden_f64verts, max_leaf):
    """builder 0 (default): binned SAH, cost-optimal 8-wide collapse and quantisation on the device, in the host builder's
    double arithmetic.  The 8-wide BVH in HBM must be the host builder's (builder 2), byte for byte -- nodes and
    leaf-ordered triangle records -- on meshes without coincident centroids; and every answer is the reference's."""
    meshes = [golden_torus["tris"], golden_f64verts["tris"]]      # (not the cube: the two triangles of a face have the same centroid)
    v, f = scenes.torus_mesh(300, 150)                      # 90,000 triangles: large-node levels, chunked passes, small subtrees
    meshes.append(scenes.mesh_triangles(v, f))
    for tris in meshes:
        gpu_ctx.set_triangles(tris)
        gpu_ctx.build(max_leaf_tris=max_leaf, builder=2)
        st_h, n_h, t_h = _wide_bytes(gpu_ctx)
        gpu_ctx.build(max_leaf_tris=max_leaf, builder=0)
        st_d, n_d, t_d = _wide_bytes(gpu_ctx)
        assert st_d.builder == 0 and st_h.builder == 2
        assert (st_d.n_wide_nodes, st_d.n_binary_nodes, st_d.tri_format, st_d.max_depth) == (st_h.n_wide_nodes, st_h.n_binary_nodes, st_h.tri_format, st_h.max_depth)
        assert t_d == t_h, "triangle records differ"
        assert n_d == n_h, "wide nodes differ"
        assert abs(st_d.sah_cost / st_h.sah_cost - 1.0) < 1e-9 and st_d.inflate == st_h.inflate
    g = golden_torus
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build(max_leaf_tris=max_leaf)
    _check_closest(gpu_ctx, g["tris"], g["rays"], g["prim"], g["t"])
    assert np.array_equal(gpu_ctx.trace_any(g["any_rays"]), g["occluded"])


def test_device_builder_edge_cases(gpu_ctx):
    r = np.array([[0.2, 0.2, 1, 0, 0, -1, 0, 1e32]], dtype=np.float32)
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float64)
    gpu_ctx.set_triangles(one)
    gpu_ctx.build()
    assert gpu_ctx.trace_closest(r)["prim"][0] == 0 and gpu_ctx.stats()["n_binary_nodes"] == 1
    # coincident centroids (duplicated triangles): the index split; exact ties -> lowest index
    for reps in (2, 9, 700, 3000):
        gpu_ctx.set_triangles(np.repeat(one, reps, axis=0))
        gpu_ctx.build()
        assert gpu_ctx.trace_closest(r)["prim"][0] == 0
        assert gpu_ctx.stats()["n_binary_nodes"] == 2 * reps - 1
    # two far-apart clusters of duplicates, and a degenerate (flat) mesh: all centroids in one plane
    two = np.concatenate([np.repeat(one, 600, axis=0), np.repeat(one + 10.0, 600, axis=0)])
    gpu_ctx.set_triangles(two)
    gpu_ctx.build()
    assert gpu_ctx.trace_closest(r)["prim"][0] == 0
    r2 = r.copy(); r2[0, :3] += 10.0
    assert gpu_ctx.trace_closest(r2)["prim"][0] == 600
    n = 40
    xs, ys = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64))
    flat = np.stack([xs, ys, 0 * xs, xs + 1, ys, 0 * xs, xs, ys + 1, 0 * xs], -1).reshape(-1, 9)
    gpu_ctx.set_triangles(flat)
    gpu_ctx.build()
    rays = np.zeros((n * n, 8), np.float32)
    rays[:, 0] = xs.ravel() + 0.25; rays[:, 1] = ys.ravel() + 0.25; rays[:, 2] = 1; rays[:, 5] = -1; rays[:, 7] = 1e32
    assert np.array_equal(gpu_ctx.trace_closest(rays)["prim"], np.arange(n * n))


def test_export_import_and_clone_share_one_tree(gpu_ctx, golden_torus):
    """spb_bvh_export -> spb_bvh_import_wide (another context, any process) and spb_ctx_clone_scene (same process, device to
    device): the adopting context answers exactly like the builder's, without building."""
    g = golden_torus
    gpu_ctx.set_triangles(g["tris"])
    gpu_ctx.build()
    st, nodes, tris = gpu_ctx.export_bvh()
    other = capi.Context(0)
    try:
        other.set_triangles(g["tris"])
        with pytest.raises(capi.SpbError):
            other.trace_closest(g["rays"])                 # nothing built or adopted yet
        other.import_wide(st, nodes, tris)
        assert other.stats()["builder"] == -1
        _check_closest(other, g["tris"], g["rays"], g["prim"], g["t"])
        bad = capi.BvhStats.from_buffer_copy(st); bad.node_bytes += 80
        with pytest.raises(capi.SpbError):
            other.import_wide(bad, nodes, tris)
    finally:
        other.close()
    clone = capi.Context(0)
    try:
        clone.clone_scene_from(gpu_ctx)
        _check_closest(clone, g["tris"], g["rays"], g["prim"], g["t"])
        assert np.array_equal(clone.trace_any(g["any_rays"]), g["occluded"])
        assert _wide_bytes(clone)[1:] == (bytes(nodes), bytes(tris))
    finally:
        clone.close()
    # a cloned scene renders the same image (materials, lights and attributes travel with it)
    img = capi.cornell_render(gpu_ctx, 48, 48, 4, seed=3, variant="glossy")
    clone = capi.Context(0)
    try:
        clone.clone_scene_from(gpu_ctx)
        cam = scenes.CORNELL_CAMERA
        c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], 48, 48)
        clone.render_begin(48, 48, c2w, r2c, max_depth=8, seed=3)
        clone.render_samples(0, 4, 1)
        assert np.allclose(clone.film_resolve(), img, rtol=1e-4, atol=1e-5)
    finally:
        clone.close()


def test_full_size_c2_against_the_compiled_reference(gpu_ctx):
    """BASELINE configs[1] at full size: the 1,000,000-triangle torus, 131,072 rays of each C2 ray set, answered by the
    unmodified reference's BVHAccel::intersect (tests/golden/make_golden_c2.py -> raycast_c2_ref.npz).  Ids bit-exact (rays whose
    two candidates are hit at exactly the same double t may name the other triangle: own-built tree), t within 1e-5."""
    from tests.golden.make_golden_c2 import c2_rays
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raycast_c2_ref.npz"))
    v, f = scenes.torus_mesh(1000, 500)
    tris = scenes.mesh_triangles(v, f)
    inc, pri, anyr = c2_rays(v)
    assert len(inc) == int(g["n"])
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build()
    assert gpu_ctx.stats()["builder"] == 0 and gpu_ctx.stats()["max_depth"] <= 14
    _check_closest(gpu_ctx, tris, inc, g["prim_incoherent"], g["t_incoherent"], max_ties=0)
    _check_closest(gpu_ctx, tris, pri, g["prim_primary"], g["t_primary"], max_ties=8)       # the pinhole grid's diagonal: genuine ties
    assert np.array_equal(gpu_ctx.trace_any(anyr), g["occluded"])


def test_gpu_lbvh_builder_gives_identical_hits(gpu_ctx, golden_torus, golden_cube):
    """builder = 1: Morton sort + Karras hierarchy + refit on the device (replaces the reference's
    top-down BVHAccel::constructRec). Any valid tree must return the reference's (prim, t)."""
    for g in (golden_torus, golden_cube):
        gpu_ctx.set_triangles(g["tris"])
        gpu_ctx.build(builder=1)
        _check_closest(gpu_ctx, g["tris"], g["rays"], g["prim"], g["t"])
        assert np.array_equal(gpu_ctx.trace_any(g["any_rays"]), g["occluded"])
    # single triangle and duplicated triangles (equal Morton codes)
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float64)
    gpu_ctx.set_triangles(one)
    gpu_ctx.build(builder=1)
    r = np.array([[0.2, 0.2, 1, 0, 0, -1, 0, 1e32]], dtype=np.float32)
    assert gpu_ctx.trace_closest(r)["prim"][0] == 0
    dup = np.repeat(one, 9, axis=0)
    gpu_ctx.set_triangles(dup)
    gpu_ctx.build(builder=1)
    assert gpu_ctx.trace_closest(r)["prim"][0] == 0          # exact tie -> lowest index
    # medium mesh against the oracle, plus the build statistics
    v, f = scenes.torus_mesh(300, 150)
    tris = scenes.mesh_triangles(v, f)
    rays = scenes.incoherent_rays(50000, v.min(0), v.max(0), seed=6)
    nodes = ob.bvh_build(tris)
    p0, t0, _, _ = ob.trace_closest(nodes, tris, rays)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.build(builder=1)
    _check_closest(gpu_ctx, tris, rays, p0, t0, max_ties=4)
    st = gpu_ctx.stats()
    assert st["n_binary_nodes"] == 2 * len(tris) - 1 and st["n_tris"] == len(tris)


def test_deep_imported_tree_continues_the_stack_in_local_memory(gpu_ctx):
    """A chain-shaped imported tree is deeper than the shared-memory stack of variants 4 and 5: variant 5 runs with the
    stack continued in local memory (variant 4 falls back to variant 3) and the answers are still the reference's."""
    from tests.test_traversal_emul import chain_tree
    v, f = scenes.torus_mesh(10, 7)
    tris = scenes.mesh_triangles(v, f)
    nodes = chain_tree(tris)
    rays = scenes.incoherent_rays(20000, v.min(0), v.max(0), seed=3)
    p0, t0, _, _ = ob.trace_closest(nodes, tris, rays)
    gpu_ctx.set_triangles(tris)
    gpu_ctx.import_binary(nodes.view(capi.IMPORT_NODE))
    assert gpu_ctx.stats()["max_depth"] > 14
    for variant in (2, 3, 4, 5):
        gpu_ctx.set_option("trace_variant", variant)
        _check_closest(gpu_ctx, tris, rays, p0, t0, max_ties=0)
        assert np.array_equal(gpu_ctx.trace_any(rays), ob.trace_any(nodes, tris, rays))
    gpu_ctx.set_option("trace_variant", DEFAULT_VARIANT)


def test_greedy_and_cost_optimal_collapse_agree(gpu_ctx, monkeypatch):
    v, f = scenes.torus_mesh(160, 80)
    tris = scenes.mesh_triangles(v, f)
    rays = scenes.incoherent_rays(100000, v.min(0), v.max(0), seed=9)
    gpu_ctx.set_triangles(tris)
    res = {}
    for col in ("1", "0"):
        monkeypatch.setenv("SPICA_BVH_COLLAPSE", col)
        gpu_ctx.build(builder=2)
        res[col] = (gpu_ctx.trace_closest(rays), gpu_ctx.stats()["n_wide_nodes"])
    assert np.array_equal(res["1"][0], res["0"][0])
    assert res["1"][1] < res["0"][1]
