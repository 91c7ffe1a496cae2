"""ctypes binding of the C ABI (include/spica_b200.h) - what tests/ and bench.py call.

This is plumbing: no algorithm lives here, and there is no fallback.  If libspica_b200.so is
missing, or there is no CUDA device, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPICA_B200_LIB") or os.path.join(HERE, "lib", "libspica_b200.so")     # the override is for A/B measurement builds (tools/)

RAY_F32 = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("tmin", "<f4"), ("tmax", "<f4")])
RAY_F64 = np.dtype([("o", "<f8", 3), ("d", "<f8", 3), ("tmin", "<f8"), ("tmax", "<f8")])
HIT = np.dtype([("t", "<f4"), ("prim", "<i4"), ("u", "<f4"), ("v", "<f4")])
HIT_F64 = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("prim", "<i4"), ("pad", "<i4")])
IMPORT_NODE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("left", "<i4"), ("right", "<i4"),
                        ("prim", "<i4"), ("axis", "<i4")])
assert RAY_F32.itemsize == 32 and RAY_F64.itemsize == 64 and HIT.itemsize == 16
assert HIT_F64.itemsize == 32 and IMPORT_NODE.itemsize == 64


class BuildOpts(C.Structure):
    _fields_ = [("builder", C.c_int32), ("max_leaf_tris", C.c_int32), ("sah_bins", C.c_int32),
                ("reserved_", C.c_int32)]


class BvhStats(C.Structure):
    _fields_ = [("n_tris", C.c_int64), ("n_wide_nodes", C.c_int64), ("n_binary_nodes", C.c_int64),
                ("node_bytes", C.c_int64), ("tri_bytes", C.c_int64), ("sah_cost", C.c_double),
                ("build_seconds", C.c_double), ("tri_format", C.c_int32), ("max_depth", C.c_int32),
                ("world_lo", C.c_double * 3), ("world_hi", C.c_double * 3), ("inflate", C.c_double),
                ("builder", C.c_int32), ("reserved_", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("last_kernel_ms", C.c_double), ("kernel_launches", C.c_int64), ("rays", C.c_int64),
                ("node_visits", C.c_int64), ("tri_tests", C.c_int64)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("distribution", C.c_int32), ("kr", C.c_float * 3), ("kt", C.c_float * 3),
                ("eta", C.c_float * 3), ("k", C.c_float * 3), ("alpha_u", C.c_float), ("alpha_v", C.c_float)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("prim", C.c_int32), ("radiance", C.c_float * 3), ("pad_", C.c_float)]


class Texture(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("reserved_", C.c_int32),
                ("texel_offset", C.c_int64), ("color0", C.c_float * 3), ("color1", C.c_float * 3),
                ("uoffset", C.c_float), ("voffset", C.c_float), ("uscale", C.c_float), ("vscale", C.c_float)]


class RenderDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32), ("filter", C.c_int32),
                ("filter_radius", C.c_double * 2), ("filter_sigma", C.c_double),
                ("camera_to_world", C.c_double * 16), ("raster_to_camera", C.c_double * 16),
                ("lens_radius", C.c_double), ("focal_distance", C.c_double), ("seed", C.c_uint64),
                ("rr_start_bounce", C.c_int32), ("integrator", C.c_int32)]


class RenderStats(C.Structure):
    _fields_ = [("paths", C.c_int64), ("rays_closest", C.c_int64), ("rays_shadow", C.c_int64),
                ("rays_mis", C.c_int64), ("kernel_launches", C.c_int64), ("render_ms", C.c_double),
                ("reduce_ms", C.c_double), ("iterations", C.c_int64)]


MAT_TYPES = {"none": -1, "diffuse": 0, "dielectric": 1, "roughconductor": 2, "roughdielectric": 3, "conductor": 4,
             "plastic": 5, "roughplastic": 6}
FILTERS = {"box": 0, "tent": 1, "gaussian": 2}
INTEGRATORS = {"path": 0, "directlighting": 1}
assert C.sizeof(Material) == 64 and C.sizeof(Light) == 24 and C.sizeof(Texture) == 64


class SpbError(RuntimeError):
    pass


_lib = None

# every symbol include/spica_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "spb_version", "spb_last_error", "spb_ctx_create", "spb_ctx_destroy",
    "spb_scene_set_triangles", "spb_bvh_build", "spb_bvh_import_binary", "spb_bvh_get_stats",
    "spb_bvh_export", "spb_bvh_import_wide", "spb_ctx_clone_scene",
    "spb_trace_closest", "spb_trace_closest_f64", "spb_trace_any", "spb_trace_any_f64",
    "spb_trace_closest_dev", "spb_trace_any_dev", "spb_get_counters", "spb_set_option",
    "spb_dev_alloc", "spb_dev_free", "spb_dev_upload", "spb_dev_download", "spb_dev_sync",
    "spb_ctx_stream",
    "spb_scene_set_triangle_attributes", "spb_scene_set_materials", "spb_scene_set_lights", "spb_scene_set_envmap",
    "spb_scene_set_textures", "spb_scene_set_material_textures",
    "spb_render_begin", "spb_render_samples", "spb_render_samples_async", "spb_render_wait", "spb_film_reduce", "spb_film_reduce_async", "spb_film_reduce_peers", "spb_film_export_handle", "spb_film_import_handles", "spb_film_reduce_imported", "spb_film_read", "spb_film_resolve", "spb_film_resolve_rgbe", "spb_film_resolve_ldr",
    "spb_film_add",
    "spb_render_get_stats", "spb_comm_get_unique_id", "spb_comm_init", "spb_film_allreduce", "spb_comm_destroy",
]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpbError("libspica_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.spb_version.restype = C.c_int
    L.spb_last_error.restype = C.c_char_p; L.spb_last_error.argtypes = [vp]
    L.spb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.spb_ctx_destroy.argtypes = [vp]; L.spb_ctx_destroy.restype = None
    L.spb_scene_set_triangles.argtypes = [vp, vp, vp, vp, vp, vp, i64]
    L.spb_bvh_build.argtypes = [vp, C.POINTER(BuildOpts)]
    L.spb_bvh_import_binary.argtypes = [vp, vp, i64, i32]
    L.spb_bvh_get_stats.argtypes = [vp, C.POINTER(BvhStats)]
    L.spb_bvh_export.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
    L.spb_bvh_import_wide.argtypes = [vp, C.POINTER(BvhStats), vp, vp]
    L.spb_ctx_clone_scene.argtypes = [vp, vp]
    for name in ("spb_trace_closest", "spb_trace_closest_f64", "spb_trace_any", "spb_trace_any_f64",
                 "spb_trace_closest_dev", "spb_trace_any_dev"):
        getattr(L, name).argtypes = [vp, vp, i64, vp]
    L.spb_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.spb_set_option.argtypes = [vp, C.c_char_p, i64]
    L.spb_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.spb_dev_free.argtypes = [vp, vp]
    L.spb_dev_upload.argtypes = [vp, vp, vp, C.c_size_t]
    L.spb_dev_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.spb_dev_sync.argtypes = [vp]
    L.spb_ctx_stream.argtypes = [vp]; L.spb_ctx_stream.restype = vp
    L.spb_scene_set_triangle_attributes.argtypes = [vp, vp, vp, i64]
    L.spb_scene_set_materials.argtypes = [vp, C.POINTER(Material), i32]
    L.spb_scene_set_lights.argtypes = [vp, C.POINTER(Light), i32]
    L.spb_scene_set_textures.argtypes = [vp, C.POINTER(Texture), i32, vp, i64]
    L.spb_scene_set_material_textures.argtypes = [vp, vp, i32]
    L.spb_scene_set_envmap.argtypes = [vp, vp, i32, i32, vp, C.c_double, vp, C.c_double]
    L.spb_render_begin.argtypes = [vp, C.POINTER(RenderDesc)]
    L.spb_render_samples.argtypes = [vp, i32, i32, i32]
    L.spb_render_samples_async.argtypes = [vp, i32, i32, i32]
    L.spb_render_wait.argtypes = [vp]
    L.spb_film_reduce.argtypes = [vp, i32]
    L.spb_film_reduce_async.argtypes = [vp, i32]
    L.spb_film_reduce_peers.argtypes = [vp, C.POINTER(vp), i32]
    L.spb_film_export_handle.argtypes = [vp, C.c_char_p]
    L.spb_film_import_handles.argtypes = [vp, C.c_char_p, i32]
    L.spb_film_reduce_imported.argtypes = [vp]
    L.spb_film_read.argtypes = [vp, vp]
    L.spb_film_resolve.argtypes = [vp, vp]
    L.spb_film_add.argtypes = [vp, vp]
    L.spb_film_resolve_rgbe.argtypes = [vp, vp]
    L.spb_film_resolve_ldr.argtypes = [vp, C.c_double, vp]
    L.spb_render_get_stats.argtypes = [vp, C.POINTER(RenderStats)]
    L.spb_comm_get_unique_id.argtypes = [C.c_char_p]
    L.spb_comm_init.argtypes = [vp, C.c_char_p, i32, i32]
    L.spb_film_allreduce.argtypes = [vp]
    L.spb_comm_destroy.argtypes = [vp]
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):          # torch tensor (pinned host or device memory)
        return a.data_ptr()
    return a.ctypes.data


class Context:
    """One per GPU (spb_ctx)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.spb_ctx_create(device, C.byref(h))
        if rc != 0:
            raise SpbError("spb_ctx_create failed (%d): %s" % (rc, self.L.spb_last_error(None).decode()))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.spb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SpbError("spica_b200 error %d: %s" % (rc, self.L.spb_last_error(self.h).decode()))

    # ---- scene / BVH
    def set_triangles(self, tris, normals=None, material_id=None, light_id=None, uvs=None):
        tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
        n = tris.shape[0]
        nm = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32).reshape(n, 9)
        uv = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(n, 6)
        mi = None if material_id is None else np.ascontiguousarray(material_id, dtype=np.int32)
        li = None if light_id is None else np.ascontiguousarray(light_id, dtype=np.int32)
        self._check(self.L.spb_scene_set_triangles(self.h, _ptr(tris), _ptr(nm), _ptr(uv), _ptr(mi), _ptr(li), n))

    def build(self, max_leaf_tris=0, sah_bins=32, builder=0):
        o = BuildOpts(builder, max_leaf_tris, sah_bins, 0)
        self._check(self.L.spb_bvh_build(self.h, C.byref(o)))

    def import_binary(self, nodes, root=0):
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 64
        self._check(self.L.spb_bvh_import_binary(self.h, _ptr(nodes), len(nodes), root))

    def stats(self):
        s = BvhStats()
        self._check(self.L.spb_bvh_get_stats(self.h, C.byref(s)))
        return {k: (list(getattr(s, k)) if k.startswith("world") else getattr(s, k)) for k, _ in s._fields_}

    def export_bvh(self):
        """(stats struct, node bytes, triangle bytes) of the 8-wide BVH in HBM."""
        s = BvhStats()
        self._check(self.L.spb_bvh_get_stats(self.h, C.byref(s)))
        nodes = np.empty(max(s.node_bytes, 1), dtype=np.uint8)
        tris = np.empty(max(s.tri_bytes, 1), dtype=np.uint8)
        self._check(self.L.spb_bvh_export(self.h, _ptr(nodes), nodes.nbytes, _ptr(tris), tris.nbytes))
        return s, nodes[:s.node_bytes], tris[:s.tri_bytes]

    def import_wide(self, stats, nodes, tris):
        self._check(self.L.spb_bvh_import_wide(self.h, C.byref(stats), _ptr(np.ascontiguousarray(nodes)), _ptr(np.ascontiguousarray(tris))))

    def clone_scene_from(self, src):
        self._check(self.L.spb_ctx_clone_scene(self.h, src.h))

    def set_option(self, name, value):
        self._check(self.L.spb_set_option(self.h, name.encode(), int(value)))

    def counters(self):
        c = Counters()
        self._check(self.L.spb_get_counters(self.h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in c._fields_}

    # ---- host-buffer entry points (numpy arrays or pinned torch tensors)
    @staticmethod
    def _rays(rays):
        if hasattr(rays, "data_ptr"):
            n = rays.shape[0]
            f64 = rays.element_size() == 8
            return rays, n, f64
        rays = np.ascontiguousarray(rays)
        if rays.dtype in (RAY_F32, RAY_F64):
            return rays, rays.shape[0], rays.dtype == RAY_F64
        assert rays.ndim == 2 and rays.shape[1] == 8 and rays.dtype in (np.float32, np.float64)
        return rays, rays.shape[0], rays.dtype == np.float64

    def trace_closest(self, rays, out=None):
        rays, n, f64 = self._rays(rays)
        if out is None:
            out = np.empty(n, dtype=HIT_F64 if f64 else HIT)
        fn = self.L.spb_trace_closest_f64 if f64 else self.L.spb_trace_closest
        self._check(fn(self.h, _ptr(rays), n, _ptr(out)))
        return out

    def trace_any(self, rays, out=None):
        rays, n, f64 = self._rays(rays)
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        fn = self.L.spb_trace_any_f64 if f64 else self.L.spb_trace_any
        self._check(fn(self.h, _ptr(rays), n, _ptr(out)))
        return out

    # ---- device-resident entry points (raw device pointers or torch cuda tensors)
    def trace_closest_dev(self, d_rays, n, d_hits):
        self._check(self.L.spb_trace_closest_dev(self.h, _ptr(d_rays), n, _ptr(d_hits)))

    def trace_any_dev(self, d_rays, n, d_occ):
        self._check(self.L.spb_trace_any_dev(self.h, _ptr(d_rays), n, _ptr(d_occ)))

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.spb_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p):
        self._check(self.L.spb_dev_free(self.h, p))

    def dev_upload(self, d_dst, host):
        host = np.ascontiguousarray(host)
        self._check(self.L.spb_dev_upload(self.h, d_dst, _ptr(host), host.nbytes))

    def dev_download(self, host, d_src):
        self._check(self.L.spb_dev_download(self.h, _ptr(host), d_src, host.nbytes))

    def sync(self):
        self._check(self.L.spb_dev_sync(self.h))

    # ---- the path tracer
    def set_materials(self, mats):
        """mats: list of dicts {type, reflectance | specularReflectance/specularTransmittance/intIOR |
        eta/k/alpha/distribution} using the reference's XML parameter names."""
        arr = (Material * max(len(mats), 1))()
        texs, texels, bind = [], [], np.full((max(len(mats), 1), 2), -1, dtype=np.int32)
        for i, m in enumerate(mats):
            m = dict(m)
            # texture-valued reflectances ({"texture": "bitmap", "image": array} / {"texture": "checkerboard", ...})
            for slot, keys in ((0, ("reflectance", "specularReflectance")), (1, ("specularTransmittance", "diffuseReflectance"))):
                for k in keys:
                    if isinstance(m.get(k), dict):
                        d = m[k]
                        t = Texture()
                        if d["texture"] == "bitmap":
                            img = np.ascontiguousarray(d["image"], dtype=np.float32)
                            t.type, t.height, t.width = 0, img.shape[0], img.shape[1]
                            t.texel_offset = sum(len(x) for x in texels) // 3
                            texels.append(img.reshape(-1))
                        else:
                            t.type = 1
                            for j in range(3):
                                t.color0[j] = d["color0"][j]; t.color1[j] = d["color1"][j]
                            t.uoffset, t.voffset, t.uscale, t.vscale = d["uoffset"], d["voffset"], d["uscale"], d["vscale"]
                        bind[i, slot] = len(texs)
                        texs.append(t)
                        m[k] = (0.0, 0.0, 0.0)
            a = arr[i]
            a.type = MAT_TYPES[m["type"]]
            a.distribution = 1 if m.get("distribution", "beckmann") == "ggx" else 0
            kr = m.get("reflectance", m.get("specularReflectance", (1.0, 1.0, 1.0)))
            kt = m.get("specularTransmittance", m.get("diffuseReflectance", (1.0, 1.0, 1.0)))
            default_ior = {"roughdielectric": 1.3333, "plastic": 1.5, "roughplastic": 1.5}.get(m["type"], 1.333)
            eta = m.get("eta", (m.get("intIOR", default_ior),) * 3)
            k = m.get("k", (0.0, 0.0, 0.0))
            for j in range(3):
                a.kr[j] = kr[j]; a.kt[j] = kt[j]; a.eta[j] = eta[j]; a.k[j] = k[j]
            a.alpha_u = a.alpha_v = m.get("alpha", 0.1)
        self._check(self.L.spb_scene_set_materials(self.h, arr, len(mats)))
        if texs:
            tarr = (Texture * len(texs))(*texs)
            flat = np.concatenate(texels) if texels else np.zeros(0, dtype=np.float32)
            self._check(self.L.spb_scene_set_textures(self.h, tarr, len(texs), _ptr(flat) if len(flat) else None, len(flat) // 3))
            self._check(self.L.spb_scene_set_material_textures(self.h, _ptr(bind), len(mats)))

    def set_lights(self, lights, envmap_at=None):
        """lights: list of (prim, (r, g, b)) area lights; envmap_at: list position of an envmap light."""
        n = len(lights) + (1 if envmap_at is not None else 0)
        arr = (Light * max(n, 1))()
        src = list(lights)
        j = 0
        for i in range(n):
            if envmap_at is not None and i == envmap_at:
                arr[i].type = 1; arr[i].prim = -1
                continue
            prim, rad = src[j]; j += 1
            arr[i].type = 0; arr[i].prim = int(prim)
            for c in range(3):
                arr[i].radiance[c] = rad[c]
        self._check(self.L.spb_scene_set_lights(self.h, arr, n))

    def set_envmap(self, rgb, to_world=None, scale=1.0, center=(0.0, 0.0, 0.0), radius=2.0):
        if rgb is None:                  # removes the environment map
            self._check(self.L.spb_scene_set_envmap(self.h, None, 0, 0, None, 1.0, None, 2.0))
            return
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        h, w = rgb.shape[:2]
        m = np.ascontiguousarray(np.eye(4) if to_world is None else to_world, dtype=np.float64)
        c = np.ascontiguousarray(center, dtype=np.float64)
        self._check(self.L.spb_scene_set_envmap(self.h, _ptr(rgb), w, h, _ptr(m), float(scale), _ptr(c), float(radius)))

    def render_begin(self, width, height, camera_to_world, raster_to_camera, max_depth=16, seed=0, filter="box",
                     filter_radius=(1.0, 1.0), filter_sigma=0.5, lens_radius=0.0, focal_distance=50.0, rr_start=3,
                     integrator="path"):
        d = RenderDesc()
        d.width, d.height, d.max_depth, d.filter = width, height, max_depth, FILTERS[filter]
        d.filter_radius[0], d.filter_radius[1], d.filter_sigma = filter_radius[0], filter_radius[1], filter_sigma
        c2w = np.asarray(camera_to_world, dtype=np.float64).reshape(16)
        r2c = np.asarray(raster_to_camera, dtype=np.float64).reshape(16)
        for i in range(16):
            d.camera_to_world[i] = c2w[i]; d.raster_to_camera[i] = r2c[i]
        d.lens_radius, d.focal_distance, d.seed, d.rr_start_bounce = lens_radius, focal_distance, seed, rr_start
        d.integrator = INTEGRATORS[integrator]
        self._film_shape = (height, width)
        self._check(self.L.spb_render_begin(self.h, C.byref(d)))

    def render_samples(self, first, count, stride=1):
        self._check(self.L.spb_render_samples(self.h, first, count, stride))

    def film_read(self):
        out = np.empty(self._film_shape + (4,), dtype=np.float32)
        self._check(self.L.spb_film_read(self.h, _ptr(out)))
        return out

    def film_resolve_rgbe(self):
        out = np.empty(self._film_shape + (4,), dtype=np.uint8)
        self._check(self.L.spb_film_resolve_rgbe(self.h, _ptr(out)))
        return out

    def film_resolve_ldr(self, gamma=2.2):
        out = np.empty(self._film_shape + (3,), dtype=np.uint8)
        self._check(self.L.spb_film_resolve_ldr(self.h, float(gamma), _ptr(out)))
        return out

    def film_resolve(self):
        out = np.empty(self._film_shape + (3,), dtype=np.float32)
        self._check(self.L.spb_film_resolve(self.h, _ptr(out)))
        return out

    def film_add(self, rgbw):
        rgbw = np.ascontiguousarray(rgbw, dtype=np.float32)
        assert rgbw.shape == self._film_shape + (4,)
        self._check(self.L.spb_film_add(self.h, _ptr(rgbw)))

    def render_stats(self):
        s = RenderStats()
        self._check(self.L.spb_render_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def comm_init(self, comm_id, n_ranks, rank):
        self._check(self.L.spb_comm_init(self.h, comm_id, n_ranks, rank))

    def film_allreduce(self):
        self._check(self.L.spb_film_allreduce(self.h))

    def film_reduce(self, root=0):
        self._check(self.L.spb_film_reduce(self.h, root))

    def film_reduce_peers(self, others):
        """One process, several GPUs: sums the films of the `others` contexts into this one over peer memory (no communicator)."""
        arr = (C.c_void_p * max(len(others), 1))(*[o.h for o in others])
        self._check(self.L.spb_film_reduce_peers(self.h, arr, len(others)))

    def film_export_handle(self):
        """96 bytes another process of this node can map this context's film with (CUDA IPC); see include/spica_b200.h."""
        buf = C.create_string_buffer(96)
        self._check(self.L.spb_film_export_handle(self.h, buf))
        return buf.raw

    def film_import_handles(self, handles):
        self._check(self.L.spb_film_import_handles(self.h, b"".join(handles), len(handles)))

    def film_reduce_imported(self):
        self._check(self.L.spb_film_reduce_imported(self.h))

    def render_samples_async(self, first, count, stride=1):
        self._check(self.L.spb_render_samples_async(self.h, first, count, stride))

    def film_reduce_async(self, root=0):
        self._check(self.L.spb_film_reduce_async(self.h, root))

    def render_wait(self):
        self._check(self.L.spb_render_wait(self.h))


def comm_unique_id():
    L = load()
    buf = C.create_string_buffer(128)
    rc = L.spb_comm_get_unique_id(buf)
    if rc != 0:
        raise SpbError("spb_comm_get_unique_id failed: %s" % L.spb_last_error(None).decode())
    return buf.raw


def cornell_render(ctx, width, height, spp, max_depth=8, variant="diffuse", seed=1, first=0, stride=1, begin=True,
                   integrator="path"):
    """Sets up the Cornell scene of spica_b200.scenes on `ctx` and renders `spp` samples per pixel."""
    from . import scenes
    if begin:
        tris, mid, lid, mats, lights = scenes.cornell_arrays(variant)
        ctx.set_triangles(tris, normals=scenes.cornell_normals(variant), material_id=mid, light_id=lid, uvs=scenes.cornell_uvs(variant))
        for m in mats:                                   # bitmap textures: the image the XML's .hdr file holds
            for v in m.values():
                if isinstance(v, dict) and v["texture"] == "bitmap":
                    v["image"] = scenes.texture_image()
        ctx.set_materials(mats)
        ctx.set_lights(lights)
        ctx.build()
        cam = scenes.CORNELL_CAMERA
        c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], width, height)
        kw = dict(filter="gaussian", lens_radius=scenes.ZOO_LENS[0], focal_distance=scenes.ZOO_LENS[1]) if variant == "zoo" else {}
        ctx.render_begin(width, height, c2w, r2c, max_depth=max_depth, seed=seed, integrator=integrator, **kw)
    ctx.render_samples(first, spp, stride)
    return ctx.film_resolve()


def envscene_render(ctx, width, height, spp, max_depth=8, nu=96, nv=48, seed=1, first=0, stride=1, begin=True, scene=None):
    """Sets up spica_b200.scenes.envscene_arrays (torus + ground under an environment map) on `ctx` and renders `spp` samples."""
    from . import scenes
    if begin:
        sc = scene or scenes.envscene_arrays(nu, nv)
        ctx.set_triangles(sc["tris"], material_id=sc["material_id"], light_id=sc["light_id"])
        ctx.set_materials(sc["materials"])
        ctx.set_envmap(sc["env"], to_world=sc["env_to_world"], scale=sc["env_scale"], radius=sc["env_radius"])
        ctx.set_lights([], envmap_at=0)
        ctx.build()
        cam = sc["camera"]
        c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], width, height)
        ctx.render_begin(width, height, c2w, r2c, max_depth=max_depth, seed=seed, filter=sc["filter"])
    ctx.render_samples(first, spp, stride)
    return ctx.film_resolve()
