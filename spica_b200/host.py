"""ctypes binding of the C++17 host (spica_b200/host -> libspica_host.so): load a Mitsuba-style
scene file through the reference-compatible plugin surface and render it on the GPU(s).
Plumbing only; errors abort the process like the reference does (core/common.h:109-115)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libspica_host.so")
CLI_PATH = os.path.join(HERE, "bin", "spica")
SYMBOLS = ["sph_render_scene", "sph_parse_scene"]
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libspica_host.so not built: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.sph_render_scene.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_ulonglong, C.c_int, C.c_void_p, C.c_longlong,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sph_parse_scene.argtypes = [C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.c_void_p, C.c_longlong]
        _lib = L
    return _lib


def render_scene(xml_path, output_prefix="/tmp/spica_host_out", gpus=1, seed=1, spp=0, max_pixels=1 << 24):
    """Renders the scene file; returns the normalised float32 image [h, w, 3] (also saved by the film)."""
    L = load()
    buf = np.empty(max_pixels * 3, dtype=np.float32)
    w, h = C.c_int(), C.c_int()
    rc = L.sph_render_scene(xml_path.encode(), output_prefix.encode(), gpus, seed, spp, buf.ctypes.data, buf.size,
                            C.byref(w), C.byref(h))
    if rc != 0:
        raise RuntimeError("image larger than max_pixels")
    return buf[: w.value * h.value * 3].reshape(h.value, w.value, 3).copy()


def parse_scene(xml_path, with_verts=True):
    L = load()
    info = (C.c_longlong * 8)()
    cam = (C.c_double * 32)()
    L.sph_parse_scene(xml_path.encode(), info, cam, None, 0)
    verts = None
    if with_verts:
        verts = np.empty((info[0], 9), dtype=np.float64)
        L.sph_parse_scene(xml_path.encode(), info, cam, verts.ctypes.data, verts.size)
    keys = ["n_triangles", "n_lights", "width", "height", "sample_count", "max_depth", "n_emitter_triangles", "filter"]
    out = {k: int(info[i]) for i, k in enumerate(keys)}
    c = np.array(cam[:], dtype=np.float64)
    out["camera_to_world"] = c[:16].reshape(4, 4)
    out["raster_to_camera"] = c[16:].reshape(4, 4)
    out["verts"] = verts
    return out


def run_cli(xml_path, output_prefix, gpus=1, seed=None, extra=()):
    """`spica -i scene.xml -o out` exactly as a user of the reference would call it."""
    cmd = [CLI_PATH, "-i", xml_path, "-o", output_prefix, "--gpus", str(gpus)] + list(extra)
    if seed is not None:
        cmd += ["--seed", str(seed)]
    return subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(CLI_PATH))
