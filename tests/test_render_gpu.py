"""Image parity of the wavefront path tracer against the UNMODIFIED reference renderer
(BASELINE.json: "relMSE against the reference path tracer ... threshold set from the reference's
own seed-to-seed variance").

Fixtures (tests/golden/cornell_*_ref.npz) hold K reference renders of the same scene at 64 spp,
made by tests/golden/make_cornell_golden.py; tau = 1.5 x the largest pairwise relMSE between
reference runs. Everything goes through the C ABI (spb_render_*)."""
import os

import numpy as np
import pytest

from spica_b200 import capi, scenes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(variant):
    g = np.load(os.path.join(GOLDEN, "cornell_%s_ref.npz" % variant))
    runs = g["runs"].astype(np.float64)
    return runs, int(g["spp"]), int(g["width"]), int(g["height"]), int(g["max_depth"]), float(g["pair_relmse"][0])


@pytest.mark.parametrize("variant", ["diffuse", "glossy", "zoo", "plastic", "textured"])
def test_cornell_relmse_against_reference(gpu_ctx, variant):
    if not os.path.exists(os.path.join(GOLDEN, "cornell_%s_ref.npz" % variant)):
        pytest.skip("no fixture for " + variant)
    runs, spp, w, h, depth, pair_max = _golden(variant)
    K = len(runs)
    mean = runs.mean(0)
    tau = 1.5 * pair_max
    # (1) equal sample count: the GPU image is as close to a reference run as reference runs are to each other
    img = capi.cornell_render(gpu_ctx, w, h, spp, max_depth=depth, variant=variant, seed=11)
    r_equal = max(scenes.rel_mse(img, runs[k], mean) for k in range(K))
    assert r_equal <= tau, (r_equal, tau)
    # (2) bias check: a converged GPU image must agree with the reference mean to within the
    # reference mean's own noise (pairwise / (2K)), with the same 1.5 margin + the GPU's residual
    hi = capi.cornell_render(gpu_ctx, w, h, 64 * spp, max_depth=depth, variant=variant, seed=12)
    r_hi = scenes.rel_mse(hi, mean, mean)
    assert r_hi <= 1.5 * pair_max / (2 * K) * (1.0 + K / 64.0), (r_hi, pair_max / (2 * K))
    # energy: the mean radiance agrees to 1 %
    assert abs(hi.mean() / mean.mean() - 1.0) < 0.01, (hi.mean(), mean.mean())
    out = os.environ.get("SPB_TEST_OUT")
    if out:
        np.save(os.path.join(out, "cornell_%s_gpu.npy" % variant), hi.astype(np.float32))


@pytest.mark.parametrize("name,variant", [("baseline_c1_ref.npz", "diffuse"), ("baseline_c4_16spp_ref.npz", "glossy")])
def test_baseline_configurations_against_reference(gpu_ctx, name, variant):
    """Image parity AT the BASELINE configurations: C1 (512x512, 64 spp, depth 8) in full, C4 (glossy Cornell box, 1920x1080,
    depth 16) at a reduced 16 spp.  The fixtures hold K = 4 runs of the unmodified reference, box-averaged over bin x bin
    pixels (tests/golden/make_baseline_golden.py); the GPU image is binned the same way; tau = 1.5 x the largest pairwise
    relMSE of the binned reference runs."""
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("no fixture " + name)
    g = np.load(path)
    runs = g["runs"].astype(np.float64)
    K, b = len(runs), int(g["bin"])
    w, h, spp, depth = int(g["width"]), int(g["height"]), int(g["spp"]), int(g["max_depth"])
    mean = runs.mean(0)
    pair_max = float(g["pair_relmse"][0])
    img = scenes.bin_image(capi.cornell_render(gpu_ctx, w, h, spp, max_depth=depth, variant=variant, seed=21).astype(np.float64), b)
    r_equal = max(scenes.rel_mse(img, runs[k], mean) for k in range(K))
    assert r_equal <= 1.5 * pair_max, (r_equal, pair_max)
    # bias: 16x the samples against the mean of the reference runs (whose own noise is pairwise / (2K))
    hi = scenes.bin_image(capi.cornell_render(gpu_ctx, w, h, 16 * spp, max_depth=depth, variant=variant, seed=22).astype(np.float64), b)
    r_hi = scenes.rel_mse(hi, mean, mean)
    assert r_hi <= 1.5 * pair_max / (2 * K) * (1.0 + K / 16.0), (r_hi, pair_max / (2 * K))
    assert abs(hi.mean() / mean.mean() - 1.0) < 0.01, (hi.mean(), mean.mean())


def test_baseline_c5_scene_in_miniature(gpu_ctx):
    """BASELINE configs[4]'s scene (torus + ground under an environment map, rough dielectric, tent filter, depth 16) with 100 k
    instead of 10 M triangles at 480x270: K = 4 runs of the unmodified reference (tests/golden/make_baseline_golden.py c5)
    against the arrays bench.py feeds the C ABI for the full-size configuration (scenes.envscene_arrays)."""
    path = os.path.join(GOLDEN, "baseline_c5_small_ref.npz")
    if not os.path.exists(path):
        pytest.skip("no fixture baseline_c5_small_ref.npz")
    g = np.load(path)
    runs = g["runs"].astype(np.float64)
    K, b = len(runs), int(g["bin"])
    w, h, spp, depth, nu, nv = (int(g[k]) for k in ("width", "height", "spp", "max_depth", "nu", "nv"))
    mean = runs.mean(0)
    pair_max = float(g["pair_relmse"][0])
    img = scenes.bin_image(capi.envscene_render(gpu_ctx, w, h, spp, max_depth=depth, nu=nu, nv=nv, seed=41).astype(np.float64), b)
    r_equal = max(scenes.rel_mse(img, runs[k], mean) for k in range(K))
    assert r_equal <= 1.5 * pair_max, (r_equal, pair_max)
    hi = scenes.bin_image(capi.envscene_render(gpu_ctx, w, h, 16 * spp, max_depth=depth, nu=nu, nv=nv, seed=42).astype(np.float64), b)
    r_hi = scenes.rel_mse(hi, mean, mean)
    assert r_hi <= 1.5 * pair_max / (2 * K) * (1.0 + K / 16.0), (r_hi, pair_max / (2 * K))
    assert abs(hi.mean() / mean.mean() - 1.0) < 0.01, (hi.mean(), mean.mean())


def test_directlighting_integrator_against_reference(gpu_ctx):
    """integrators/directlighting on the "zoo" scene (it has a mirror): emitted light at depth 0
    only, one light sample per vertex, continuation through specular reflection only.  The reference's own
    directlighting renders are the bar; their seed-to-seed variance is ~15x lower than the path tracer's."""
    g = np.load(os.path.join(GOLDEN, "cornell_direct_ref.npz"))
    runs = g["runs"].astype(np.float64)
    K = len(runs)
    mean = runs.mean(0)
    pair_max = float(g["pair_relmse"][0])
    spp, w, h, depth = int(g["spp"]), int(g["width"]), int(g["height"]), int(g["max_depth"])
    img = capi.cornell_render(gpu_ctx, w, h, spp, max_depth=depth, variant="zoo", seed=31, integrator="directlighting")
    r_equal = max(scenes.rel_mse(img, runs[k], mean) for k in range(K))
    assert r_equal <= 1.5 * pair_max, (r_equal, pair_max)
    hi = capi.cornell_render(gpu_ctx, w, h, 64 * spp, max_depth=depth, variant="zoo", seed=32, integrator="directlighting")
    r_hi = scenes.rel_mse(hi, mean, mean)
    assert r_hi <= 1.5 * pair_max / (2 * K) * (1.0 + K / 64.0), (r_hi, pair_max / (2 * K))
    assert abs(hi.mean() / mean.mean() - 1.0) < 0.01, (hi.mean(), mean.mean())
    # and it is not the path tracer: no indirect light
    pt = capi.cornell_render(gpu_ctx, w, h, spp, max_depth=depth, variant="zoo", seed=31)
    assert pt.mean() > 1.3 * img.mean()


def test_film_output_stage_on_the_device(gpu_ctx, tmp_path):
    """spb_film_resolve_rgbe / _ldr == the reference's pixel encodings (core/image.cc:60-88; core/tmo.cc:53-71 +
    core/image.cc:484-487) applied on the host to the floats spb_film_resolve returns for the same film."""
    capi.cornell_render(gpu_ctx, 96, 64, 8, max_depth=6, variant="glossy", seed=4)
    rgb = gpu_ctx.film_resolve().astype(np.float64)
    d = rgb.max(-1)
    m, e = np.frexp(d)
    lit = d > 1e-32
    scale = np.where(lit, m * 256.0 / np.where(lit, d, 1.0), 0.0)
    want = np.zeros(rgb.shape[:2] + (4,), dtype=np.uint8)
    want[..., :3] = (rgb * scale[..., None]).astype(np.uint8)
    want[..., 3] = np.where(lit, e + 128, 0).astype(np.uint8)
    got = gpu_ctx.film_resolve_rgbe()
    assert np.array_equal(got, want)
    for gamma in (2.2, 1.0):
        ldr = gpu_ctx.film_resolve_ldr(gamma)
        ref = (255.0 * np.clip(np.power(rgb, 1.0 / gamma), 0.0, 1.0)).astype(np.uint8)
        diff = np.abs(ldr.astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-3        # CUDA's pow is not correctly rounded: a value on a step boundary may move
    # the films of the C++ host write their files from these bytes: same file as encoding on the host
    import subprocess
    from spica_b200 import host
    xml = os.path.join(os.path.dirname(GOLDEN), "golden", "scenes", "cornell_diffuse.xml")
    a = host.run_cli(xml, str(tmp_path / "dev"), seed=9)
    env = dict(os.environ, SPICA_HOST_ENCODE="1")
    b = subprocess.run([host.CLI_PATH, "-i", xml, "-o", str(tmp_path / "hostenc"), "--seed", "9"], capture_output=True, text=True,
                       cwd=os.path.dirname(host.CLI_PATH), env=env)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    ia, ib = scenes.read_hdr(str(tmp_path / "dev.hdr")), scenes.read_hdr(str(tmp_path / "hostenc.hdr"))
    assert (np.abs(ia - ib) <= ib.max(-1, keepdims=True) / 128 + 1e-6).all()      # same seed; float atomics may reorder


def test_sample_partition_is_deterministic(gpu_ctx):
    """Counter-based sampler: rendering samples {0..7} in one call == two interleaved halves
    (what two GPUs do) up to float32 summation order."""
    a = capi.cornell_render(gpu_ctx, 64, 64, 8, seed=5)
    fa = gpu_ctx.film_read()
    capi.cornell_render(gpu_ctx, 64, 64, 4, seed=5, first=0, stride=2)
    gpu_ctx.render_samples(1, 4, 2)
    fb = gpu_ctx.film_read()
    assert np.allclose(fa, fb, rtol=1e-4, atol=1e-5)
    assert np.array_equal(fa[..., 3], np.full((64, 64), 8.0, dtype=np.float32))
    # small waves (many waves per pass) give the same film
    gpu_ctx.set_option("wave_slots", 4096)
    capi.cornell_render(gpu_ctx, 64, 64, 8, seed=5)
    fc = gpu_ctx.film_read()
    gpu_ctx.set_option("wave_slots", 1 << 22)
    assert np.allclose(fa, fc, rtol=1e-4, atol=1e-5)
    assert np.isfinite(a).all()


def test_film_add_and_stats(gpu_ctx):
    capi.cornell_render(gpu_ctx, 32, 32, 2, seed=3)
    f = gpu_ctx.film_read()
    st = gpu_ctx.render_stats()
    assert st["paths"] == 32 * 32 * 2 and st["rays_closest"] >= st["paths"] and st["rays_shadow"] > 0
    gpu_ctx.film_add(f)
    assert np.allclose(gpu_ctx.film_read(), 2 * f)


def test_async_render_graph_and_single_rank_reduce(gpu_ctx):
    """spb_render_samples_async + spb_render_wait, the plain-launch loop and a (single-rank) NCCL film reduce all give the
    film of the synchronous call (same sample set; float atomics only reorder the sums)."""
    capi.cornell_render(gpu_ctx, 96, 80, 6, seed=8, variant="glossy")
    ref = gpu_ctx.film_read()
    # asynchronous, in two queued halves
    capi.cornell_render(gpu_ctx, 96, 80, 0, seed=8, variant="glossy")
    gpu_ctx.render_samples_async(0, 3, 1)
    gpu_ctx.render_samples_async(3, 3, 1)
    gpu_ctx.render_wait()
    assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
    st = gpu_ctx.render_stats()
    assert st["paths"] == 96 * 80 * 6 and st["iterations"] > 0 and st["rays_closest"] >= st["paths"]
    # plain launches instead of the CUDA graph
    gpu_ctx.set_option("render_graph", 0)
    capi.cornell_render(gpu_ctx, 96, 80, 6, seed=8, variant="glossy")
    gpu_ctx.set_option("render_graph", 1)
    assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
    # a one-rank communicator: reduce to root 0 and all-reduce are the identity
    try:
        cid = capi.comm_unique_id()
    except capi.SpbError:
        pytest.skip("libnccl not loadable")
    gpu_ctx.comm_init(cid, 1, 0)
    gpu_ctx.film_reduce(0)
    gpu_ctx.film_reduce_async(-1)
    gpu_ctx.render_wait()
    assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
    assert gpu_ctx.render_stats()["reduce_ms"] > 0
    gpu_ctx.L.spb_comm_destroy(gpu_ctx.h)


def test_film_reduce_over_peer_memory(gpu_ctx):
    """spb_film_reduce_peers: three contexts of one process render interleaved sample indices; one kernel on the root's GPU
    sums the other films into the root's (peer loads over NVLink when the contexts sit on different GPUs, plain loads on the
    same GPU -- a one-GPU box runs the latter).  The sum is the one-context film; the other films stay as they were."""
    capi.cornell_render(gpu_ctx, 96, 80, 6, seed=8, variant="glossy")
    ref = gpu_ctx.film_read()
    devices = [0, 0]
    try:
        capi.Context(1).close()
        devices = [1, 0]
    except capi.SpbError:
        pass
    others = [capi.Context(d) for d in devices]
    try:
        capi.cornell_render(gpu_ctx, 96, 80, 0, seed=8, variant="glossy")
        gpu_ctx.render_samples_async(0, 2, 3)
        for k, o in enumerate(others):
            capi.cornell_render(o, 96, 80, 0, seed=8, variant="glossy")
            o.render_samples_async(k + 1, 2, 3)          # still queued when the reduce is called: it waits for them
        gpu_ctx.film_reduce_peers(others)
        assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
        assert gpu_ctx.render_stats()["reduce_ms"] > 0
        for k, o in enumerate(others):
            w = o.film_read()[..., 3].mean()                          # two of the six samples per pixel, untouched by the reduce
            assert abs(w - ref[..., 3].mean() / 3) < 1e-3 * ref[..., 3].mean()
        gpu_ctx.film_reduce_peers([])                                 # nothing to add
        assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
        with pytest.raises(capi.SpbError):
            gpu_ctx.film_reduce_peers([gpu_ctx])
        with pytest.raises(capi.SpbError):
            gpu_ctx.film_reduce_peers([others[0], others[0]])
        capi.cornell_render(others[1], 64, 64, 1, seed=8)             # a film of another size
        with pytest.raises(capi.SpbError):
            gpu_ctx.film_reduce_peers(others)
    finally:
        for o in others:
            o.close()


_IPC_CHILD = """
import sys
sys.path.insert(0, %r)
from spica_b200 import capi
try:
    ctx = capi.Context(1)
except capi.SpbError:
    ctx = capi.Context(0)
capi.cornell_render(ctx, 96, 80, 0, seed=8, variant="glossy")
ctx.render_samples(1, 3, 2)                      # the odd sample indices
print(ctx.film_export_handle().hex(), flush=True)
sys.stdin.readline()                             # the parent has summed: now the film may go
ctx.close()
"""


def test_film_reduce_across_processes_over_cuda_ipc(gpu_ctx):
    """spb_film_export_handle / spb_film_import_handles / spb_film_reduce_imported: another PROCESS (on the second GPU when
    there is one, else on the same GPU) renders the odd sample indices and ships 96 bytes; this process maps that film and adds
    it to its own with the peer-sum kernel.  The pipe is the job's barrier."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    capi.cornell_render(gpu_ctx, 96, 80, 6, seed=8, variant="glossy")
    ref = gpu_ctx.film_read()
    child = subprocess.Popen([sys.executable, "-c", _IPC_CHILD % root], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    try:
        line = ""
        while len(line.strip()) != 192:                               # (skips log lines)
            line = child.stdout.readline()
            assert line, "the child process died"
        handle = bytes.fromhex(line.strip())
        capi.cornell_render(gpu_ctx, 96, 80, 0, seed=8, variant="glossy")
        gpu_ctx.render_samples(0, 3, 2)                               # the even sample indices
        gpu_ctx.film_import_handles([handle])
        gpu_ctx.film_reduce_imported()
        assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
        assert gpu_ctx.render_stats()["reduce_ms"] > 0
        with pytest.raises(capi.SpbError):
            gpu_ctx.film_import_handles([b"\0" * 96])                 # not a handle
        with pytest.raises(capi.SpbError):
            gpu_ctx.film_import_handles([gpu_ctx.film_export_handle()])   # a film of this very process cannot be mapped
        gpu_ctx.film_import_handles([])
        gpu_ctx.film_reduce_imported()                                # nothing mapped: the film stays
        assert np.allclose(gpu_ctx.film_read(), ref, rtol=1e-4, atol=1e-5)
    finally:
        try:
            child.stdin.write("done\n"); child.stdin.flush()
        except OSError:
            pass
        child.wait(timeout=60)


def test_scene_change_invalidates_a_begun_render(gpu_ctx):
    """spb_scene_set_triangles after spb_render_begin used to leave the kernels' parameter block pointing at the
    freed tree.  Now every scene change requires a new spb_render_begin."""
    capi.cornell_render(gpu_ctx, 32, 32, 1, seed=2)
    tris, mid, lid, mats, lights = scenes.cornell_arrays("diffuse")
    gpu_ctx.set_triangles(tris, material_id=mid, light_id=lid)
    with pytest.raises(capi.SpbError):
        gpu_ctx.render_samples(1, 1, 1)            # no tree, stale shading records
    gpu_ctx.build()
    with pytest.raises(capi.SpbError):
        gpu_ctx.render_samples(1, 1, 1)            # still no new spb_render_begin
    capi.cornell_render(gpu_ctx, 32, 32, 1, seed=2)
    gpu_ctx.set_materials(mats)
    with pytest.raises(capi.SpbError):
        gpu_ctx.render_samples(1, 1, 1)
    # the bounce counter of a path has 12 bits
    cam = scenes.CORNELL_CAMERA
    c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], 32, 32)
    with pytest.raises(capi.SpbError):
        gpu_ctx.render_begin(32, 32, c2w, r2c, max_depth=5000)
    with pytest.raises(capi.SpbError):
        gpu_ctx.set_option("trace_variant", 99)
    img = capi.cornell_render(gpu_ctx, 32, 32, 2, seed=2)
    assert np.isfinite(img).all() and img.mean() > 0


def test_unusual_scenes_render(gpu_ctx):
    """Paths the fixtures do not reach: vertices that are not float32-exact (80-byte triangle records: the traversal falls back to
    the float64-triangle kernels inside the render loop), and a scene with no triangles at all under an environment map."""
    ref = capi.cornell_render(gpu_ctx, 64, 64, 32, seed=6)
    tris, mid, lid, mats, lights = scenes.cornell_arrays("diffuse")
    gpu_ctx.set_triangles(tris * (1.0 + 1e-10), material_id=mid, light_id=lid)
    gpu_ctx.set_materials(mats)
    gpu_ctx.set_lights(lights)
    gpu_ctx.build()
    assert gpu_ctx.stats()["tri_format"] == 1
    cam = scenes.CORNELL_CAMERA
    c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], 64, 64)
    gpu_ctx.render_begin(64, 64, c2w, r2c, max_depth=8, seed=6)
    gpu_ctx.render_samples(0, 32, 1)
    img = gpu_ctx.film_resolve()
    assert np.isfinite(img).all() and abs(img.mean() / ref.mean() - 1.0) < 0.02, (img.mean(), ref.mean())
    # nothing but the sky: every camera ray escapes and is answered by Envmap::Le (lights/envmap.cc:130-134)
    sc = scenes.envscene_arrays(8, 4)
    gpu_ctx.set_triangles(np.zeros((0, 9)))
    gpu_ctx.set_materials([])
    gpu_ctx.set_envmap(sc["env"], to_world=sc["env_to_world"], scale=sc["env_scale"], radius=sc["env_radius"])
    gpu_ctx.set_lights([], envmap_at=0)
    gpu_ctx.build()
    cam = sc["camera"]
    c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], 48, 32)
    gpu_ctx.render_begin(48, 32, c2w, r2c, max_depth=4, seed=1)
    gpu_ctx.render_samples(0, 4, 1)
    sky = gpu_ctx.film_resolve()
    st = gpu_ctx.render_stats()
    assert np.isfinite(sky).all() and sky.min() > 0 and st["paths"] == 48 * 32 * 4 and st["rays_shadow"] == 0
    assert np.array_equal(gpu_ctx.film_read()[..., 3], np.full((32, 48), 4.0, dtype=np.float32))
    gpu_ctx.set_envmap(None)
