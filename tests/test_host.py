"""The C++17 host (spica's plugin surface + Mitsuba-style scene loading + `spica -i scene.xml`).
CPU tests cover the host logic up to the C-ABI boundary (no device needed); GPU tests render the
committed scene files through the host library and the CLI and compare with the reference images."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from spica_b200 import host, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_host_library_exports():
    L = host.load()
    for s in host.SYMBOLS:
        assert hasattr(L, s)
    assert os.access(host.CLI_PATH, os.X_OK)


@pytest.mark.parametrize("variant", ["diffuse", "glossy", "zoo", "plastic"])
def test_parse_cornell_matches_the_flat_scene(variant):
    r = host.parse_scene(os.path.join(SCENES, "cornell_%s.xml" % variant))
    tris, mid, lid, mats, lights = scenes.cornell_arrays(variant)
    assert r["n_triangles"] == len(tris) == 32
    assert r["n_lights"] == len(lights) == 2 and r["n_emitter_triangles"] == 2      # one AreaLight per emitter triangle
    assert (r["width"], r["height"], r["sample_count"], r["max_depth"]) == (128, 128, 64, 8)
    assert np.array_equal(r["verts"], tris)                                          # float32 positions widened to double
    cam = scenes.CORNELL_CAMERA
    c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], 128, 128)
    assert np.allclose(r["camera_to_world"], c2w, atol=1e-15)
    assert np.allclose(r["raster_to_camera"], r2c, atol=1e-15)


def test_ply_shape_transform_and_defaults(tmp_path):
    v, f = scenes.torus_mesh(16, 8)
    scenes.write_ply(str(tmp_path / "t.ply"), v, f)
    xml = """<?xml version="1.0"?>
<!-- comment -->
<scene version="0.5.0">
  <integrator type="path"/>
  <sensor type="perspective">
    <float name="fov" value="45"/>
    <transform name="toWorld"><lookAt origin="0 0 5" target="0 0 0" up="0 1 0"/></transform>
    <sampler type="ldsampler"><integer name="sampleCount" value="4"/></sampler>
    <film type="ldrfilm"><integer name="width" value="40"/><integer name="height" value="20"/><rfilter type="gaussian"/></film>
  </sensor>
  <bsdf type="diffuse" id="m"><rgb name="reflectance" value="0.5"/></bsdf>
  <shape type="ply">
    <string name="filename" value="t.ply"/>
    <transform name="toWorld"><scale x="2" y="2" z="2"/><translate x="1" y="0" z="0"/></transform>
    <ref id="m"/>
  </shape>
</scene>"""
    p = tmp_path / "s.xml"
    p.write_text(xml)
    r = host.parse_scene(str(p))
    assert r["n_triangles"] == 256 and r["n_lights"] == 0
    assert r["max_depth"] == 16 and r["sample_count"] == 4 and r["filter"] == 2      # parser default maxDepth (sceneparser.cc:63)
    want = scenes.mesh_triangles(v, f).reshape(-1, 3, 3) * 2.0 + np.array([1.0, 0.0, 0.0])
    assert np.allclose(r["verts"].reshape(-1, 3, 3), want, atol=1e-12)
    assert (r["width"], r["height"]) == (40, 20)


def test_ply_loader_large_mesh_polygons_and_broken_files(tmp_path):
    """The PLY loader reads the face block in one piece and builds the triangles (and the primitives of a mesh of >= 4096
    triangles, in one allocation) on all cores: same triangles as the arrays scene; polygons keep their first three indices
    (core/meshio.cc:143-160); a truncated file and an index out of range abort with the reference's kind of message."""
    import struct
    import sys
    d = str(tmp_path)
    xml = scenes.write_envscene(d, 64, 64, 4, 4, name="e", nu=100, nv=60)            # 12,000 + 2 triangles: the block path
    r = host.parse_scene(xml)
    assert r["n_triangles"] == 12002 and np.array_equal(r["verts"], scenes.envscene_arrays(100, 60)["tris"])
    ply = os.path.join(d, "e_torus.ply")
    verts = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1]], np.float32)

    def write(faces, truncate=0):
        hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
               "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(verts), len(faces))).encode()
        body = verts.tobytes() + b"".join(struct.pack("<B%di" % len(f), len(f), *f) for f in faces)
        open(ply, "wb").write(hdr + (body[:-truncate] if truncate else body))

    write([(0, 1, 2, 3), (0, 1, 4)])                                                 # a quad, then a triangle
    r = host.parse_scene(xml)
    assert r["n_triangles"] == 4
    assert r["verts"][0].tolist() == [0, 0, 0, 1, 0, 0, 1, 1, 0] and r["verts"][1].tolist() == [0, 0, 0, 1, 0, 0, 0, 0, 1]
    code = "from spica_b200 import host; host.parse_scene(%r)" % xml
    for faces, truncate, message in (([(0, 1, 2), (0, 1, 4)], 5, "truncated"), ([(0, 1, 9)], 0, "index out of range")):
        write(faces, truncate)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
        assert p.returncode != 0 and message in (p.stdout + p.stderr)


def test_cli_fails_loudly_without_device_or_scene(tmp_path):
    from tests.conftest import has_cuda_device
    r = subprocess.run([host.CLI_PATH, "-i", str(tmp_path / "missing.xml")], capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to open" in r.stderr
    if not has_cuda_device():
        r = host.run_cli(os.path.join(SCENES, "cornell_diffuse.xml"), str(tmp_path / "o"))
        assert r.returncode != 0 and "no CUDA device" in r.stderr             # no CPU fallback


@pytest.mark.gpu
def test_host_render_matches_reference(tmp_path):
    g = np.load(os.path.join(GOLDEN, "cornell_diffuse_ref.npz"))
    runs = g["runs"].astype(np.float64)
    mean = runs.mean(0)
    tau = 1.5 * float(g["pair_relmse"][0])
    img = host.render_scene(os.path.join(SCENES, "cornell_diffuse.xml"), str(tmp_path / "lib_out"), seed=3)
    assert img.shape == (128, 128, 3)
    assert max(scenes.rel_mse(img, r, mean) for r in runs) <= tau
    # the film wrote <prefix>.hdr like the reference's hdrfilm; RGBE keeps 8 mantissa bits
    hdr = scenes.read_hdr(str(tmp_path / "lib_out.hdr"))
    assert (np.abs(hdr - img) <= img.max(-1, keepdims=True) / 128 + 1e-6).all()      # shared exponent
    # the CLI, called as a user of the reference calls it
    r = host.run_cli(os.path.join(SCENES, "cornell_diffuse.xml"), str(tmp_path / "cli_out"), seed=3)
    assert r.returncode == 0, r.stderr
    cli = scenes.read_hdr(str(tmp_path / "cli_out.hdr"))
    # same seed -> same samples; the film's float atomics may add in a different order, which can move an
    # RGBE mantissa by one step
    assert (np.abs(cli - hdr) <= hdr.max(-1, keepdims=True) / 128 + 1e-6).all()
    assert "Finish!!" in r.stdout


@pytest.mark.gpu
def test_textured_scene_through_the_host_matches_the_direct_call(tmp_path):
    """bitmap / checkerboard textures + OBJ vn/vt through the XML loader == the same scene set up through the C ABI."""
    from spica_b200 import capi
    xml = os.path.join(SCENES, "cornell_textured.xml")
    img = host.render_scene(xml, str(tmp_path / "tex"), seed=5)
    ctx = capi.Context(0)
    direct = capi.cornell_render(ctx, 128, 128, 64, max_depth=8, variant="textured", seed=5)
    ctx.close()
    assert np.allclose(img, direct, rtol=1e-4, atol=1e-5)
    g = np.load(os.path.join(GOLDEN, "cornell_textured_ref.npz"))
    runs = g["runs"].astype(np.float64)
    assert max(scenes.rel_mse(img, r, runs.mean(0)) for r in runs) <= 1.5 * float(g["pair_relmse"][0])


@pytest.mark.gpu
def test_envmap_roughdielectric_ply_scene_matches_reference(tmp_path):
    """Environment lighting (lights/envmap.cc), a PLY mesh with a GGX rough dielectric
    (bsdfs/roughdielectric.cc), a tent filter and a rotated envmap, loaded from the committed XML
    through the host: the reference's own image statistics are the bar."""
    g = np.load(os.path.join(GOLDEN, "envtorus_ref.npz"))
    runs = g["runs"].astype(np.float64)
    K = len(runs)
    mean = runs.mean(0)
    pair_max = float(g["pair_relmse"][0])
    xml = os.path.join(SCENES, "envtorus.xml")
    r = host.parse_scene(xml, with_verts=False)
    assert r["n_lights"] == 1 and r["n_emitter_triangles"] == 0 and r["filter"] == 1
    img = host.render_scene(xml, str(tmp_path / "env"), seed=21)
    assert max(scenes.rel_mse(img, run, mean) for run in runs) <= 1.5 * pair_max
    hi = host.render_scene(xml, str(tmp_path / "env_hi"), seed=22, spp=64 * int(g["spp"]))
    r_hi = scenes.rel_mse(hi, mean, mean)
    assert r_hi <= 1.5 * pair_max / (2 * K) * (1.0 + K / 64.0), (r_hi, pair_max / (2 * K))
    assert abs(hi.mean() / mean.mean() - 1.0) < 0.01, (hi.mean(), mean.mean())
    out = os.environ.get("SPB_TEST_OUT")
    if out:
        np.save(os.path.join(out, "envtorus_gpu.npy"), hi.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("film_reduce", ["peers", "nccl"])
def test_two_gpus_render_the_single_gpu_image(tmp_path, monkeypatch, film_reduce):
    """`--gpus 2`: the replica's scene and BVH are cloned device to device (spb_ctx_clone_scene), the sample indices are
    interleaved, the films are summed into GPU 0's -- by a kernel reading the other film over peer memory (default) or by one
    NCCL reduce (SPICA_FILM_REDUCE=nccl) -- and the image is the one-GPU image up to float summation order.
    (Skipped on a one-GPU box.)"""
    from spica_b200 import capi
    monkeypatch.setenv("SPICA_FILM_REDUCE", film_reduce)
    try:
        capi.Context(1).close()          # (not torch.cuda.device_count(): importing torch after the CUDA libraries of this repo fails)
    except capi.SpbError:
        pytest.skip("needs two GPUs")
    xml = os.path.join(SCENES, "cornell_glossy.xml")
    one = host.render_scene(xml, str(tmp_path / "g1"), gpus=1, seed=13)
    two = host.render_scene(xml, str(tmp_path / "g2"), gpus=2, seed=13)
    assert np.allclose(one, two, rtol=1e-4, atol=1e-5)
    xml = os.path.join(SCENES, "envtorus.xml")
    one = host.render_scene(xml, str(tmp_path / "e1"), gpus=1, seed=14)
    two = host.render_scene(xml, str(tmp_path / "e2"), gpus=2, seed=14)
    assert np.allclose(one, two, rtol=1e-4, atol=1e-5)
