"""The plugin shim for the UNMODIFIED reference host: the reference's own
`spica` binary (oracle/_ref/bin/spica, prebuilt by oracle/Makefile) loads this repo's plugins/path.so
and plugins/bvh.so through its own loader (core/cobject.cc:32-50) and renders on the GPU.
CPU tests check that the reference's loader accepts the plugins, that the scene it parsed is
resolved into the right PODs, and that there is no CPU fallback; GPU tests compare the image the
reference's film wrote with the reference's own renders."""
import os

import numpy as np
import pytest

from spica_b200 import capi, host, refhost, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(not refhost.available(REF), reason="needs the compiled reference host (oracle/_ref) and the shim plugins")


def _dump(tmp_path, xml, **kw):
    d = str(tmp_path / "dump.bin")
    r = refhost.run(xml, str(tmp_path / "out"), str(tmp_path / "run"), REF, env={"SPICA_B200_DUMP_SCENE": d, "SPICA_SEED": 5}, **kw)
    assert r.returncode == 0, r.stderr
    assert "not rendering" in r.stdout
    return refhost.read_dump(d)


@pytest.mark.parametrize("variant", ["diffuse", "glossy", "zoo", "plastic"])
@pytest.mark.parametrize("gpu_plugins", [("path", "bvh"), ("path",)])
def test_reference_host_scene_resolves_to_the_flat_scene(tmp_path, variant, gpu_plugins):
    """The scene as the reference's parser built it == the flat arrays the C ABI tests use; also with the
    reference's own `bvh` accelerator left in place (the integrator then builds its tree itself)."""
    s = _dump(tmp_path, os.path.join(SCENES, "cornell_%s.xml" % variant), gpu_plugins=gpu_plugins)
    tris, mid, lid, mats, lights = scenes.cornell_arrays(variant)
    assert np.array_equal(s["verts"], tris)
    assert np.array_equal(s["material_id"], mid) and np.array_equal(s["light_id"], lid)
    assert not s["any_normals"] and not s["any_uv"]             # OBJ quads without vn / vt: face normals, zero uv
    assert len(s["materials"]) == len(mats)
    for got, want in zip(s["materials"], mats):
        assert got["type"] == capi.MAT_TYPES[want["type"]]
        for key, field in (("reflectance", "kr"), ("specularReflectance", "kr"), ("specularTransmittance", "kt"), ("eta", "eta"), ("k", "k")):
            if key in want:
                assert np.allclose(got[field], want[key]), (key, got, want)
        if "alpha" in want:
            assert got["alpha_u"] == got["alpha_v"] == pytest.approx(want["alpha"])
        if "intIOR" in want:
            assert got["eta"][0] == pytest.approx(want["intIOR"])
        assert got["distribution"] == (1 if want.get("distribution") == "ggx" else 0)
    assert [(l["prim"], l["type"]) for l in s["lights"]] == [(p, 0) for p, _ in lights]
    for l, (_, rad) in zip(s["lights"], lights):
        assert np.allclose(l["radiance"], rad)
    d = s["desc"]
    assert (d["width"], d["height"], d["max_depth"], s["sample_count"], d["seed"], d["rr_start_bounce"]) == (128, 128, 8, 64, 5, 3)
    cam = scenes.CORNELL_CAMERA
    c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], 128, 128)
    assert np.allclose(d["camera_to_world"].reshape(4, 4), c2w, atol=1e-15)
    assert np.allclose(d["raster_to_camera"].reshape(4, 4), r2c, atol=1e-15)


def test_reference_host_textured_scene_resolves(tmp_path):
    """bitmap + checkerboard textures on reflectance-type parameters, OBJ files with vn + vt: texture PODs, bindings,
    texels and per-triangle texcoords as the reference's parser built them."""
    s = _dump(tmp_path, os.path.join(SCENES, "cornell_textured.xml"))
    tris, mid, lid, mats, lights = scenes.cornell_arrays("textured")
    assert np.array_equal(s["verts"], tris) and np.array_equal(s["material_id"], mid)
    assert s["any_uv"] and np.array_equal(s["uvs"], scenes.cornell_uvs("textured"))
    assert np.allclose(s["normals"], scenes.cornell_normals("textured"), atol=1e-7)
    assert len(s["materials"]) == len(mats)
    kinds = {0: "bitmap", 1: "checkerboard"}
    for got, want, bind in zip(s["materials"], mats, s["material_textures"]):
        assert got["type"] == capi.MAT_TYPES[want["type"]]
        for slot, keys in ((0, ("reflectance", "specularReflectance")), (1, ("specularTransmittance", "diffuseReflectance"))):
            wanted = [want[k] for k in keys if k in want]
            if wanted and isinstance(wanted[0], dict):
                t = s["textures"][bind[slot]]
                assert kinds[t["type"]] == wanted[0]["texture"]
                if t["type"] == 1:
                    assert np.allclose(t["color0"], wanted[0]["color0"]) and np.allclose(t["color1"], wanted[0]["color1"])
                    assert (t["uoffset"], t["voffset"], t["uscale"], t["vscale"]) == (0.0, 0.25, 3.0, 2.0)
                else:
                    img = scenes.texture_image()
                    assert (t["height"], t["width"]) == img.shape[:2]
                    tex = s["texels"][t["texel_offset"]:t["texel_offset"] + img.shape[0] * img.shape[1]].reshape(img.shape)
                    assert np.array_equal(tex, img)                  # the RGBE file holds these texels exactly
            else:
                assert bind[slot] == -1
    assert len(s["textures"]) == 3                                   # one object per <texture> element: checker, poster, decal


def test_reference_host_envmap_scene_resolves(tmp_path):
    """PLY mesh with vertex normals, rough dielectric, tent filter, rotated environment map: the shim's
    PODs equal what this repo's own C++ host resolves from the same file."""
    xml = os.path.join(SCENES, "envtorus.xml")
    s = _dump(tmp_path, xml)
    r = host.parse_scene(xml)
    assert s["n_triangles"] == r["n_triangles"] and np.array_equal(s["verts"], r["verts"])
    assert np.allclose(s["desc"]["camera_to_world"].reshape(4, 4), r["camera_to_world"], atol=1e-15)
    assert np.allclose(s["desc"]["raster_to_camera"].reshape(4, 4), r["raster_to_camera"], atol=1e-15)
    assert s["desc"]["filter"] == r["filter"] == 1 and list(s["desc"]["filter_radius"]) == [1.0, 1.0]
    assert [l["type"] for l in s["lights"]] == [1] and (s["light_id"] == -1).all()
    assert s["env_radius"] == 6.0 and s["env_rgb"].shape[2] == 3
    env = scenes.read_hdr(os.path.join(SCENES, "envtorus_env.hdr"))
    # level 0 of the reference's pyramid = its own RGBE decode (m / 256, core/image.cc:380-382; read_hdr adds the half step) x scale 1.5
    assert s["env_rgb"].shape == env.shape and (np.abs(s["env_rgb"] - 1.5 * env) <= 1.5 * env.max(-1, keepdims=True) / 256 + 1e-6).all()
    types = sorted(m["type"] for m in s["materials"])
    assert types == [0, 3]                                       # diffuse ground, roughdielectric torus
    rd = [m for m in s["materials"] if m["type"] == 3][0]
    assert rd["alpha_u"] > 0 and rd["eta"][0] > 1


def test_reference_host_has_no_cpu_fallback(tmp_path):
    from tests.conftest import has_cuda_device
    if has_cuda_device():
        pytest.skip("a device is present")
    r = refhost.run(os.path.join(SCENES, "cornell_diffuse.xml"), str(tmp_path / "o"), str(tmp_path / "run"), REF)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not os.path.exists(str(tmp_path / "o.hdr"))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["diffuse", "zoo", "plastic", "textured"])
def test_reference_host_renders_on_the_gpu(tmp_path, variant):
    """`spica -i scene.xml` of the unmodified reference, GPU plugins swapped in: the .hdr its hdrfilm
    wrote is within the reference's own seed-to-seed variance of the reference's renders."""
    g = np.load(os.path.join(GOLDEN, "cornell_%s_ref.npz" % variant))
    runs = g["runs"].astype(np.float64)
    mean = runs.mean(0)
    tau = 1.5 * float(g["pair_relmse"][0])
    out = str(tmp_path / "out")
    r = refhost.run(os.path.join(SCENES, "cornell_%s.xml" % variant), out, str(tmp_path / "run"), REF, env={"SPICA_SEED": 3})
    assert r.returncode == 0, r.stderr
    assert "Finish!!" in r.stdout and "GPU BVH" in r.stdout
    img = scenes.read_hdr(out + ".hdr")
    assert img.shape == (128, 128, 3)
    assert max(scenes.rel_mse(img, run, mean) for run in runs) <= tau
    # same seed through this repo's own host: the same image up to the RGBE quantisation of the file
    own = host.render_scene(os.path.join(SCENES, "cornell_%s.xml" % variant), str(tmp_path / "own"), seed=3)
    assert (np.abs(img - own) <= own.max(-1, keepdims=True) / 128 + 1e-6).all()


@pytest.mark.gpu
def test_reference_host_on_two_gpus(tmp_path):
    """SPICA_GPUS=2 under the unmodified reference host: the replica is cloned from the accelerator's context, the films meet on
    GPU 0 (over peer memory; with SPICA_FILM_REDUCE=nccl through a communicator that comes up on its own threads) -- and the
    file is the one-GPU file.  (Skipped on a one-GPU box.)"""
    from spica_b200 import capi
    try:
        capi.Context(1).close()          # (not torch.cuda.device_count(): importing torch after the CUDA libraries of this repo fails)
    except capi.SpbError:
        pytest.skip("needs two GPUs")
    xml = os.path.join(SCENES, "cornell_zoo.xml")
    imgs = []
    for k, (g, how) in enumerate(((1, "peers"), (2, "peers"), (2, "nccl"))):
        out = str(tmp_path / ("out%d" % k))
        r = refhost.run(xml, out, str(tmp_path / ("run%d" % k)), REF, env={"SPICA_SEED": 5, "SPICA_GPUS": g, "SPICA_FILM_REDUCE": how})
        assert r.returncode == 0, r.stderr
        imgs.append(scenes.read_hdr(out + ".hdr"))
    for other in imgs[1:]:
        assert (np.abs(imgs[0] - other) <= imgs[0].max(-1, keepdims=True) / 64 + 1e-6).all()      # RGBE quantisation of two float summation orders


@pytest.mark.gpu
def test_reference_host_directlighting_plugin(tmp_path):
    """plugins/directlighting.so: `<integrator type="directlighting">` of the unmodified reference host on the GPU."""
    g = np.load(os.path.join(GOLDEN, "cornell_direct_ref.npz"))
    runs = g["runs"].astype(np.float64)
    mean = runs.mean(0)
    out = str(tmp_path / "out")
    r = refhost.run(os.path.join(SCENES, "cornell_direct.xml"), out, str(tmp_path / "run"), REF, env={"SPICA_SEED": 7})
    assert r.returncode == 0, r.stderr
    img = scenes.read_hdr(out + ".hdr")
    assert max(scenes.rel_mse(img, run, mean) for run in runs) <= 1.5 * float(g["pair_relmse"][0])


@pytest.mark.gpu
def test_reference_host_envmap_scene_and_reference_accelerator(tmp_path):
    """Environment lighting through the shim, once with the GPU `bvh` plugin and once with the
    reference's own accelerator left in place: same seed, same image."""
    g = np.load(os.path.join(GOLDEN, "envtorus_ref.npz"))
    runs = g["runs"].astype(np.float64)
    mean = runs.mean(0)
    xml = os.path.join(SCENES, "envtorus.xml")
    imgs = []
    for plugins in [("path", "bvh"), ("path",)]:
        out = str(tmp_path / ("out_" + "_".join(plugins)))
        r = refhost.run(xml, out, str(tmp_path / "run"), REF, env={"SPICA_SEED": 21}, gpu_plugins=plugins)
        assert r.returncode == 0, r.stderr
        imgs.append(scenes.read_hdr(out + ".hdr"))
    # same seed, same samples; the film's float atomics may add in a different order, so allow one RGBE step
    assert (np.abs(imgs[0] - imgs[1]) <= imgs[0].max(-1, keepdims=True) / 128 + 1e-6).all()
    assert max(scenes.rel_mse(imgs[0], run, mean) for run in runs) <= 1.5 * float(g["pair_relmse"][0])


@pytest.mark.gpu
def test_reference_integrator_on_the_gpu_accelerator(tmp_path):
    """The other half of the drop-in: the reference's OWN path integrator (CPU) calling the scalar
    Accelerator::intersect of the GPU `bvh` plugin.  One launch per ray, so a tiny image."""
    xml = scenes.write_cornell(str(tmp_path / "scene"), 24, 24, 16, 4, variant="diffuse", name="tiny")
    out_gpu, out_ref = str(tmp_path / "gpu_accel"), str(tmp_path / "ref_accel")
    r = refhost.run(xml, out_gpu, str(tmp_path / "run1"), REF, threads=2, gpu_plugins=("bvh",))
    assert r.returncode == 0, r.stderr
    assert "GPU BVH" in r.stdout
    r2 = refhost.run(xml, out_ref, str(tmp_path / "run2"), REF, threads=2, gpu_plugins=())
    assert r2.returncode == 0, r2.stderr
    a, b = scenes.read_hdr(out_gpu + ".hdr"), scenes.read_hdr(out_ref + ".hdr")
    assert a.shape == b.shape == (24, 24, 3) and np.isfinite(a).all()
    # different time(0)/thread seeds: compare the means of two 16-spp renders loosely (9216 paths each: the ratio of the means
    # scatters by ~4 %; at the 2 spp this test used first it scattered by ~12 % and the 0.25 bar failed once in ~20 runs)
    assert abs(a.mean() / b.mean() - 1.0) < 0.25, (a.mean(), b.mean())
