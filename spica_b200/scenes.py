"""Synthetic scenes and ray sets named by BASELINE.json.

Everything here is deterministic (counter-based hashes, no global RNG state) so that the CPU
oracle, the compiled reference and the GPU see bit-identical float32 inputs.
"""
import os
import struct

import numpy as np

INFTY = np.float32(1.0e32)  # reference core/common.h:54, as stored in a float32 ray record


def _hash_u32(x):
    """lowbias32-style integer hash on uint32 arrays (wraps mod 2^32)."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def _u01(idx, dim, seed):
    """Counter-based uniform in [0,1): keyed by (ray index, dimension, seed), 24-bit mantissa."""
    idx = np.asarray(idx, dtype=np.uint64)
    k = _hash_u32((idx & 0xFFFFFFFF) ^ np.uint64((seed * 0x9E3779B9) & 0xFFFFFFFF))
    k = _hash_u32(k.astype(np.uint64) + np.uint64((dim * 0x85EBCA6B + (seed << 7)) & 0xFFFFFFFF)
                  + (idx >> 32))
    return (k >> 8).astype(np.float64) * (1.0 / 16777216.0)


def torus_mesh(nu=1000, nv=500, R=1.0, r=0.4, bump=0.05, seed=1):
    """Closed displaced torus with exactly 2*nu*nv triangles, float32 vertices.

    Returns (verts float32 [nu*nv,3], faces int32 [2*nu*nv,3])."""
    i = np.arange(nu, dtype=np.int64)[:, None]
    j = np.arange(nv, dtype=np.int64)[None, :]
    h = _hash_u32((i * 73856093 ^ j * 19349663 ^ (seed * 83492791)) & 0xFFFFFFFF)
    h = (h >> 8).astype(np.float64) * (2.0 / 16777216.0) - 1.0
    u = 2.0 * np.pi * i / nu
    v = 2.0 * np.pi * j / nv
    rr = r * (1.0 + bump * h)
    x = (R + rr * np.cos(v)) * np.cos(u)
    y = (R + rr * np.cos(v)) * np.sin(u)
    z = rr * np.sin(v) + 0.0 * u
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    vid = lambda a, b: ((a % nu) * nv + (b % nv))
    a = vid(i, j); b = vid(i + 1, j); c = vid(i + 1, j + 1); d = vid(i, j + 1)
    f0 = np.stack([a, b, c], axis=-1).reshape(-1, 3)
    f1 = np.stack([a, c, d], axis=-1).reshape(-1, 3)
    faces = np.empty((2 * nu * nv, 3), dtype=np.int32)
    faces[0::2] = f0
    faces[1::2] = f1
    return verts, faces


def mesh_triangles(verts, faces, dtype=np.float64):
    """[n,9] world-space triangle soup (p0,p1,p2) - exact widening of the float32 vertices."""
    return verts.astype(dtype)[faces].reshape(-1, 9)


def write_ply(path, verts, faces):
    """binary_little_endian PLY the reference's loader accepts (core/meshio.cc:76-166)."""
    verts = np.ascontiguousarray(verts, dtype="<f4")
    faces = np.ascontiguousarray(faces, dtype="<i4")
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\n"
           "property float x\nproperty float y\nproperty float z\n"
           "element face %d\nproperty list uchar int vertex_indices\nend_header\n"
           % (len(verts), len(faces)))
    rec = np.empty(len(faces), dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    rec["n"] = 3
    rec["i"] = faces
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(verts.tobytes())
        f.write(rec.tobytes())


def read_ply(path):
    """Minimal reader for the same PLY subset (float xyz [+ ignored extra float props], tri faces)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    nv = nf = 0
    nprops = 0
    section = None
    for line in data[:end].decode("ascii", "replace").splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "element":
            section = t[1]
            if t[1] == "vertex":
                nv = int(t[2])
            elif t[1] == "face":
                nf = int(t[2])
        elif t[0] == "property" and section == "vertex":
            nprops += 1
    verts = np.frombuffer(data, dtype="<f4", count=nv * nprops, offset=end).reshape(nv, nprops)[:, :3]
    off = end + nv * nprops * 4
    rec = np.frombuffer(data, dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]), count=nf, offset=off)
    assert (rec["n"] == 3).all(), "only triangle faces supported"
    return np.array(verts, dtype=np.float32), np.array(rec["i"], dtype=np.int32)


def primary_rays(width=4096, height=4096, eye=(0.0, 0.0, 4.0), fov_deg=40.0):
    """Pinhole at `eye` looking at the origin (-z), one ray per pixel centre; float32 [n,8]."""
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64),
                         indexing="ij")
    th = np.tan(np.deg2rad(fov_deg) * 0.5)
    aspect = width / height
    px = ((xs + 0.5) / width * 2.0 - 1.0) * th * aspect
    py = (1.0 - (ys + 0.5) / height * 2.0) * th
    rays = np.zeros((width * height, 8), dtype=np.float32)
    rays[:, 0:3] = np.asarray(eye, dtype=np.float32)
    rays[:, 3] = px.reshape(-1)
    rays[:, 4] = py.reshape(-1)
    rays[:, 5] = -1.0
    rays[:, 7] = INFTY
    return rays


def incoherent_rays(n, lo, hi, seed=2, start=0, anyhit=False):
    """Origins uniform in the AABB [lo,hi] inflated by 10 %, directions uniform on the sphere,
    from a counter-based generator keyed (ray index, seed). float32 [n,8].

    anyhit=True sets tmax to the distance to a second uniform point (shadow-ray style)."""
    lo = np.asarray(lo, dtype=np.float64); hi = np.asarray(hi, dtype=np.float64)
    c = 0.5 * (lo + hi); e = 0.5 * (hi - lo) * 1.1
    idx = np.arange(start, start + n, dtype=np.uint64)
    rays = np.zeros((n, 8), dtype=np.float32)
    for k in range(3):
        rays[:, k] = (c[k] + e[k] * (2.0 * _u01(idx, k, seed) - 1.0)).astype(np.float32)
    z = 2.0 * _u01(idx, 3, seed) - 1.0
    phi = 2.0 * np.pi * _u01(idx, 4, seed)
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    rays[:, 3] = (s * np.cos(phi)).astype(np.float32)
    rays[:, 4] = (s * np.sin(phi)).astype(np.float32)
    rays[:, 5] = z.astype(np.float32)
    if anyhit:
        q = np.stack([c[k] + e[k] * (2.0 * _u01(idx, 5 + k, seed) - 1.0) for k in range(3)], -1)
        d = q - rays[:, 0:3].astype(np.float64)
        rays[:, 3:6] = d.astype(np.float32)
        rays[:, 7] = np.sqrt((d * d).sum(-1)).astype(np.float32)
    else:
        rays[:, 7] = INFTY
    return rays


def cube_triangles():
    """Unit cube [0,1]^3 as 12 triangles (same shape as the reference fixture box.ply: 8 v / 12 f)."""
    v = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = []
    for a, b, c, d in quads:
        f += [(a, b, c), (a, c, d)]
    return v, np.array(f, dtype=np.int32)


# ---------------------------------------------------------------------------------------------
# Cornell-box scenes (BASELINE.json configs 0, 2, 3): geometry as quads -> OBJ files + a
# Mitsuba-0.5-style XML the reference's parser accepts (spica/sceneparser.cc:190-334), and the
# same scene as flat arrays for the C ABI.
# ---------------------------------------------------------------------------------------------

def _quad(a, b, c, d):
    return [np.asarray(p, dtype=np.float64) for p in (a, b, c, d)]


def _box_quads(cx, cz, sx, sy, sz, y0, angle):
    """Axis-aligned box of half sizes (sx, sz), height sy standing on y0, rotated about +y by `angle`;
    outward-facing quads (5: no bottom)."""
    ca, sa = np.cos(angle), np.sin(angle)

    def P(x, y, z):
        return (cx + ca * x + sa * z, y0 + y, cz - sa * x + ca * z)
    x0, x1, z0, z1, y1 = -sx, sx, -sz, sz, sy
    return [
        _quad(P(x0, y1, z0), P(x0, y1, z1), P(x1, y1, z1), P(x1, y1, z0)),   # top (+y)
        _quad(P(x0, 0, z1), P(x1, 0, z1), P(x1, y1, z1), P(x0, y1, z1)),     # +z
        _quad(P(x1, 0, z0), P(x0, 0, z0), P(x0, y1, z0), P(x1, y1, z0)),     # -z
        _quad(P(x1, 0, z1), P(x1, 0, z0), P(x1, y1, z0), P(x1, y1, z1)),     # +x
        _quad(P(x0, 0, z0), P(x0, 0, z1), P(x0, y1, z1), P(x0, y1, z0)),     # -x
    ]


def cornell_parts(variant="diffuse"):
    """List of parts: (name, bsdf id, quads, radiance or None).

    variant "diffuse": BASELINE config 0 / 2 (all Lambertian + one emitter quad);
    variant "glossy" : config 3 (tall box rough conductor, short box dielectric);
    variant "zoo"    : specular conductor, Beckmann rough dielectric, GGX rough conductor back wall; its
                       XML also uses a thin-lens camera and a Gaussian filter;
    variant "textured": checkerboard floor, bitmap back wall, plastic tall box with a bitmap diffuse layer
                       (textures/checkerboard.cc, textures/bitmap.cc); its OBJ files carry vn + vt, which the
                       reference needs both of to keep the texcoords (core/meshio.cc:193-233);
    variant "plastic": SURVEY 8f rank 2 -- smooth plastic tall box, GGX rough plastic short box, Beckmann rough
                       plastic floor."""
    parts = [
        ("floor", {"plastic": "lacquer", "textured": "checker"}.get(variant, "white"), [_quad((-1, -1, -1), (-1, -1, 1), (1, -1, 1), (1, -1, -1))], None),
        ("ceiling", "white", [_quad((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1))], None),
        ("back", {"zoo": "brushed", "textured": "poster"}.get(variant, "white"), [_quad((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1))], None),
        ("left", "red", [_quad((-1, -1, -1), (-1, 1, -1), (-1, 1, 1), (-1, -1, 1))], None),
        ("right", "green", [_quad((1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1))], None),
        # emitter: faces down (e1 x e2 = -y), just below the ceiling
        ("light", "black", [_quad((-0.25, 0.995, -0.25), (0.25, 0.995, -0.25), (0.25, 0.995, 0.25), (-0.25, 0.995, 0.25))],
         (17.0, 12.0, 4.0)),
        ("tallbox", {"glossy": "metal", "zoo": "mirror", "plastic": "plastic", "textured": "decal"}.get(variant, "white"), _box_quads(-0.33, -0.3, 0.3, 1.2, 0.3, -1.0, 0.3), None),
        ("shortbox", {"glossy": "glass", "zoo": "frosted", "plastic": "satin"}.get(variant, "white"), _box_quads(0.35, 0.3, 0.3, 0.6, 0.3, -1.0, -0.3), None),
    ]
    return parts


CORNELL_BSDFS = {
    # id: (plugin type, {param: value})
    "white": ("diffuse", {"reflectance": (0.75, 0.75, 0.75)}),
    "red": ("diffuse", {"reflectance": (0.75, 0.25, 0.25)}),
    "green": ("diffuse", {"reflectance": (0.25, 0.75, 0.25)}),
    "black": ("diffuse", {"reflectance": (0.0, 0.0, 0.0)}),
    "metal": ("roughconductor", {"eta": (0.2, 0.92, 1.1), "k": (3.9, 2.45, 2.14), "alpha": 0.1}),
    # variant "zoo": the lobes / options the other scenes do not touch
    "mirror": ("conductor", {"eta": (0.2, 0.92, 1.1), "k": (3.9, 2.45, 2.14)}),
    "frosted": ("roughdielectric", {"specularReflectance": (1.0, 1.0, 1.0), "specularTransmittance": (1.0, 1.0, 1.0),
                                    "alpha": 0.2, "intIOR": 1.5, "distribution": "beckmann"}),
    "brushed": ("roughconductor", {"eta": (0.2, 0.92, 1.1), "k": (3.9, 2.45, 2.14), "alpha": 0.3, "distribution": "ggx"}),
    # variant "plastic"
    "plastic": ("plastic", {"diffuseReflectance": (0.2, 0.3, 0.7), "specularReflectance": (1.0, 1.0, 1.0), "intIOR": 1.5}),
    "satin": ("roughplastic", {"diffuseReflectance": (0.7, 0.3, 0.2), "specularReflectance": (0.9, 0.9, 0.9), "intIOR": 1.6,
                               "alpha": 0.15, "distribution": "ggx"}),
    "lacquer": ("roughplastic", {"diffuseReflectance": (0.6, 0.6, 0.6), "specularReflectance": (1.0, 1.0, 1.0), "alpha": 0.08}),
    # variant "textured": a texture-valued parameter is a dict {"texture": plugin type, ...plugin parameters}
    "checker": ("diffuse", {"reflectance": {"texture": "checkerboard", "color0": (0.8, 0.8, 0.8), "color1": (0.15, 0.2, 0.55),
                                            "uoffset": 0.0, "voffset": 0.25, "uscale": 3.0, "vscale": 2.0}}),
    "poster": ("diffuse", {"reflectance": {"texture": "bitmap"}}),
    "decal": ("plastic", {"diffuseReflectance": {"texture": "bitmap"}, "specularReflectance": (1.0, 1.0, 1.0), "intIOR": 1.5}),
    "glass": ("dielectric", {"specularReflectance": (1.0, 1.0, 1.0), "specularTransmittance": (1.0, 1.0, 1.0), "intIOR": 1.5}),
}
ZOO_LENS = (0.04, 3.4)      # apertureRadius, focusDistance of the "zoo" variant
CORNELL_CAMERA = {"origin": (0.0, 0.0, 3.9), "target": (0.0, 0.0, 0.0), "up": (0.0, 1.0, 0.0), "fov": 39.3}


QUAD_UVS = ((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0))     # texcoords of a quad's corners (variant "textured")


def texture_image(w=16, h=8):
    """The bitmap of the "textured" variant. Every texel is a multiple of 1/256 with its largest channel in
    [0.5, 1), so the RGBE file holds it exactly and the reference's decode (m / 256 * 2^e, core/image.cc:380-382)
    returns the same floats this function does."""
    y, x = np.mgrid[0:h, 0:w]
    r = 128 + ((x * 37 + y * 11) % 120)
    g = 40 + ((x * 13 + y * 29) % 160)
    b = 30 + (((x // 2 + y // 2) % 2) * 150)
    img = np.stack([r, g, b], -1).astype(np.float32) / 256.0
    assert (img.max(-1) >= 0.5).all() and (img.max(-1) < 1.0).all()
    return img


def cornell_uvs(variant):
    """[n_triangles, 6] float32 texcoords in cornell_arrays order, or None when the variant has none."""
    if variant != "textured":
        return None
    out = []
    for _, _, quads, _ in cornell_parts(variant):
        for _ in quads:
            out.append([c for k in (0, 1, 2) for c in QUAD_UVS[k]])
            out.append([c for k in (0, 2, 3) for c in QUAD_UVS[k]])
    return np.asarray(out, dtype=np.float32)


def cornell_normals(variant):
    """[n_triangles, 9] float32 vertex normals (= the quad's face normal, as written to the OBJ) or None."""
    if variant != "textured":
        return None
    out = []
    for _, _, quads, _ in cornell_parts(variant):
        for a, b, c, d in quads:
            n = np.cross(b - a, c - a)
            n = (n / np.linalg.norm(n)).astype(np.float32)
            out.append(np.tile(n, 3)); out.append(np.tile(n, 3))
    return np.asarray(out, dtype=np.float32)


def _quads_to_tris(quads):
    out = []
    for a, b, c, d in quads:
        out.append(np.concatenate([a, b, c]))
        out.append(np.concatenate([a, c, d]))
    return out


def write_cornell(dirpath, width=512, height=512, spp=64, max_depth=8, variant="diffuse", name="cornell", integrator="path"):
    """Writes <dirpath>/<name>.xml + one OBJ per part; returns the XML path."""
    os.makedirs(dirpath, exist_ok=True)
    parts = cornell_parts(variant)
    used = []
    for pname, bsdf, quads, rad in parts:
        with open(os.path.join(dirpath, "%s_%s.obj" % (name, pname)), "w") as f:
            f.write("# %s\no %s\n" % (pname, pname))
            nv = 0
            for qi, q in enumerate(quads):
                for p in q:
                    f.write("v %.9g %.9g %.9g\n" % (np.float32(p[0]), np.float32(p[1]), np.float32(p[2])))
                if variant == "textured":
                    n = np.cross(q[1] - q[0], q[2] - q[0])
                    n = (n / np.linalg.norm(n)).astype(np.float32)
                    f.write("vn %.9g %.9g %.9g\n" % (n[0], n[1], n[2]))
                    for uv in QUAD_UVS:
                        f.write("vt %g %g\n" % uv)
                    ix = [(nv + k, nv + k, qi + 1) for k in (1, 2, 3, 4)]
                    for tri in ((0, 1, 2), (0, 2, 3)):
                        f.write("f " + " ".join("%d/%d/%d" % ix[k] for k in tri) + "\n")
                else:
                    f.write("f %d %d %d\nf %d %d %d\n" % (nv + 1, nv + 2, nv + 3, nv + 1, nv + 3, nv + 4))
                nv += 4
        if bsdf not in used:
            used.append(bsdf)
    cam = CORNELL_CAMERA
    x = ['<?xml version="1.0" encoding="utf-8"?>', '<scene version="0.5.0">',
         '  <integrator type="%s">' % integrator, '    <integer name="maxDepth" value="%d"/>' % max_depth, '  </integrator>',
         '  <sensor type="perspective">', '    <float name="fov" value="%g"/>' % cam["fov"],
         '    <transform name="toWorld">',
         '      <lookAt origin="%g, %g, %g" target="%g, %g, %g" up="%g, %g, %g"/>' % (cam["origin"] + cam["target"] + cam["up"]),
         '    </transform>',
         '    <sampler type="independent">', '      <integer name="sampleCount" value="%d"/>' % spp, '    </sampler>',
         '    <film type="hdrfilm">', '      <integer name="width" value="%d"/>' % width,
         '      <integer name="height" value="%d"/>' % height,
         '      <rfilter type="gaussian"/>' if variant == "zoo" else '      <rfilter type="box"/>', '    </film>', '  </sensor>']
    if variant == "zoo":
        i = x.index('    <float name="fov" value="%g"/>' % cam["fov"])
        x[i + 1:i + 1] = ['    <float name="apertureRadius" value="%g"/>' % ZOO_LENS[0], '    <float name="focusDistance" value="%g"/>' % ZOO_LENS[1]]
    for b in used:
        typ, prm = CORNELL_BSDFS[b]
        x.append('  <bsdf type="%s" id="%s">' % (typ, b))
        for k, v in prm.items():
            if isinstance(v, dict):                      # nested texture plugin (spica/sceneparser.cc:303-327)
                x.append('    <texture type="%s" name="%s">' % (v["texture"], k))
                if v["texture"] == "bitmap":
                    write_hdr(os.path.join(dirpath, name + "_tex.hdr"), texture_image())
                    x.append('      <string name="filename" value="%s_tex.hdr"/>' % name)
                for tk, tv in v.items():
                    if tk == "texture":
                        continue
                    if isinstance(tv, tuple):
                        x.append('      <rgb name="%s" value="%g, %g, %g"/>' % ((tk,) + tv))
                    else:
                        x.append('      <float name="%s" value="%g"/>' % (tk, tv))
                x.append('    </texture>')
            elif isinstance(v, tuple):
                x.append('    <rgb name="%s" value="%g, %g, %g"/>' % ((k,) + v))
            elif isinstance(v, str):
                x.append('    <string name="%s" value="%s"/>' % (k, v))
            else:
                x.append('    <float name="%s" value="%g"/>' % (k, v))
        x.append('  </bsdf>')
    for pname, bsdf, quads, rad in parts:
        x.append('  <shape type="obj">')
        x.append('    <string name="filename" value="%s_%s.obj"/>' % (name, pname))
        x.append('    <ref id="%s"/>' % bsdf)
        if rad is not None:
            x.append('    <emitter type="area">')
            x.append('      <rgb name="radiance" value="%g, %g, %g"/>' % rad)
            x.append('    </emitter>')
        x.append('  </shape>')
    x.append('</scene>')
    path = os.path.join(dirpath, name + ".xml")
    with open(path, "w") as f:
        f.write("\n".join(x) + "\n")
    return path


def cornell_arrays(variant="diffuse"):
    """The same scene flattened for the C ABI: (tris [n,9] float64 (float32-exact), material_id,
    light_id, materials [list of dict], lights [list of (prim, radiance)])."""
    parts = cornell_parts(variant)
    ids = {}
    mats = []
    tris, mid, lid, lights = [], [], [], []
    for pname, bsdf, quads, rad in parts:
        if bsdf not in ids:
            ids[bsdf] = len(mats)
            typ, prm = CORNELL_BSDFS[bsdf]
            mats.append(dict(type=typ, **prm))
        for t in _quads_to_tris(quads):
            t = t.astype(np.float32).astype(np.float64)     # what tinyobjloader hands the reference (meshio.cc:178-205)
            if rad is not None:
                lid.append(len(lights))
                lights.append((len(tris), rad))
            else:
                lid.append(-1)
            tris.append(t)
            mid.append(ids[bsdf])
    return (np.asarray(tris, dtype=np.float64), np.asarray(mid, dtype=np.int32), np.asarray(lid, dtype=np.int32),
            mats, lights)


# ---- camera matrices, restating core/transform.cc:165-212 and core/camera.cc:20-38 ----------------

def look_at(origin, target, up):
    eye = np.asarray(origin, dtype=np.float64)
    d = np.asarray(target, dtype=np.float64) - eye
    d /= np.linalg.norm(d)
    u = np.asarray(up, dtype=np.float64)
    left = np.cross(u / np.linalg.norm(u), d)
    left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0] = left; m[:3, 1] = new_up; m[:3, 2] = d; m[:3, 3] = eye
    return m


def perspective_camera(to_world, fov_deg, width, height):
    """(camera_to_world, raster_to_camera) as the reference's PerspectiveCamera builds them."""
    fov = fov_deg * np.pi / 180.0
    aspect = width / height
    near, far = 1.0e-2, 1000.0
    pers = np.array([[1.0 / aspect, 0, 0, 0], [0, 1, 0, 0], [0, 0, far / (far - near), -far * near / (far - near)],
                     [0, 0, 1, 0]], dtype=np.float64)
    it = 1.0 / np.tan(fov / 2.0)
    c2s = np.diag([it, it, 1.0, 1.0]) @ pers
    # screen = [-1,1]^2: screenToRaster = scale(w,h,1) * scale(1/2,-1/2,1) * translate(1,-1,0)
    tr = np.eye(4); tr[0, 3] = 1.0; tr[1, 3] = -1.0
    s2r = np.diag([float(width), float(height), 1.0, 1.0]) @ np.diag([0.5, -0.5, 1.0, 1.0]) @ tr
    r2c = np.linalg.inv(c2s) @ np.linalg.inv(s2r)
    return np.ascontiguousarray(to_world, dtype=np.float64), np.ascontiguousarray(r2c)


def read_hdr(path):
    """Radiance RGBE reader for the files the reference writes (core/image.cc:390-435: per-scanline,
    per-channel literal runs). Decodes mantissa m as (m + 0.5) / 256 to undo the truncation of
    core/image.cc:72-88 without bias. Returns float32 [h, w, 3]."""
    with open(path, "rb") as f:
        data = f.read()
    pos = data.index(b"\n\n") + 2
    end = data.index(b"\n", pos)
    t = data[pos:end].split()
    h, w = int(t[1]), int(t[3])
    pos = end + 1
    buf = np.frombuffer(data, dtype=np.uint8)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    for y in range(h):
        assert buf[pos] == 2 and buf[pos + 1] == 2
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                n = int(buf[pos]); pos += 1
                if n > 128:          # run
                    n -= 128
                    img[y, x:x + n, c] = buf[pos]; pos += 1
                else:
                    img[y, x:x + n, c] = buf[pos:pos + n]; pos += n
                x += n
    e = img[..., 3].astype(np.float64)
    scale = np.where(e > 0, np.exp2(e - 128.0 - 8.0), 0.0)
    rgb = (img[..., :3].astype(np.float64) + 0.5) * scale[..., None]
    rgb[e == 0] = 0.0
    return rgb.astype(np.float32)


def bin_image(img, b):
    """Box average over b x b pixel blocks (the size is cropped to a multiple of b)."""
    if b <= 1:
        return img
    h, w = (img.shape[0] // b) * b, (img.shape[1] // b) * b
    return img[:h, :w].reshape(h // b, b, w // b, b, -1).mean(axis=(1, 3))


def rel_mse(a, b, ref):
    """mean over pixels and channels of (a-b)^2 / (ref^2 + 1e-2)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.mean((a - b) ** 2 / (ref ** 2 + 1e-2)))


def write_hdr(path, img):
    """Radiance RGBE writer in the layout the reference reads and writes (core/image.cc:270-435):
    0x02 0x02 scanline marker, per-channel literal runs of <= 127."""
    img = np.asarray(img, dtype=np.float64)
    h, w = img.shape[:2]
    d = img.max(-1)
    m, e = np.frexp(d)
    scale = np.where(d > 1e-32, m * 256.0 / np.maximum(d, 1e-300), 0.0)
    rgbe = np.zeros((h, w, 4), dtype=np.uint8)
    rgbe[..., :3] = np.clip(img * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(d > 1e-32, e + 128, 0).astype(np.uint8)
    out = bytearray()
    out += b"#?RADIANCE\n# Made with 100% pure HDR Shop\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0000000000000\n\n"
    out += ("-Y %d +X %d\n" % (h, w)).encode()
    for y in range(h):
        out += bytes([2, 2, (w >> 8) & 0xFF, w & 0xFF])
        for c in range(4):
            x = 0
            while x < w:
                n = min(127, w - x)
                out.append(n)
                out += rgbe[y, x:x + n, c].tobytes()
                x += n
    with open(path, "wb") as f:
        f.write(bytes(out))


def synthetic_envmap(w=64, h=32):
    """Small lat-long sky: horizon gradient + a bright sun lobe + a coloured ground."""
    v = (np.arange(h) + 0.5) / h
    u = (np.arange(w) + 0.5) / w
    theta = np.pi * v[:, None]
    phi = 2 * np.pi * u[None, :]
    sky = np.stack([0.35 + 0.3 * np.cos(theta) ** 2 + 0 * phi, 0.45 + 0.3 * np.cos(theta) ** 2 + 0 * phi,
                    0.7 + 0.25 * np.cos(theta) + 0 * phi], -1)
    sun_dir = np.array([np.sin(0.7) * np.cos(1.0), np.sin(0.7) * np.sin(1.0), np.cos(0.7)])
    d = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta) + 0 * phi], -1)
    sun = np.exp(np.minimum(0.0, (d @ sun_dir - 1.0) * 60.0))[..., None] * np.array([30.0, 26.0, 20.0])
    img = np.where(theta[..., None] > np.pi * 0.62, np.array([0.25, 0.2, 0.15]) * (1 + 0 * sky), sky) + sun
    return np.ascontiguousarray(img, dtype=np.float32)


ENV_CAMERA = dict(origin=(0.3, -2.6, 1.5), target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), fov=40.0)


def envscene_envmap():
    """The texels the reference reads back from write_envscene's .hdr file: RGBE-quantised, decoded as core/image.cc:380-382
    does (mantissa / 256 x 2^(e - 128), no half step)."""
    img = np.asarray(synthetic_envmap(), dtype=np.float64)
    d = img.max(-1)
    m, e = np.frexp(d)
    scale = np.where(d > 1e-32, m * 256.0 / np.maximum(d, 1e-300), 0.0)
    mant = np.clip(img * scale[..., None], 0, 255).astype(np.uint8).astype(np.float64)
    out = mant * np.where(d > 1e-32, np.exp2(e.astype(np.float64) - 8.0), 0.0)[..., None]
    return np.ascontiguousarray(out, dtype=np.float32)


def envscene_arrays(nu=96, nv=48):
    """write_envscene's scene as the arrays the C ABI takes (what the hosts resolve that XML file into): the torus
    (flat shaded: a PLY carries no normals, core/meshio.cc:76-166) + the ground quad, the two materials, the
    environment map with its toWorld rotation (radians, spica/sceneparser.cc:147-148), the camera.
    BASELINE configs[4] is envscene_arrays(2500, 2000): 10,000,000 + 2 triangles."""
    v, f = torus_mesh(nu, nv, R=0.6, r=0.25, bump=0.0)
    torus = mesh_triangles(v, f)
    # the OBJ reader keeps float32 positions (core/meshio.cc: tinyobj's float), so the ground is at float32(-0.3); a double -0.3
    # here would also push the whole scene into the 96-byte double-precision triangle records (data layout)
    g = np.array([[-2, -2, -0.3], [2, -2, -0.3], [2, 2, -0.3], [-2, 2, -0.3]], dtype=np.float32).astype(np.float64)
    ground = np.stack([np.concatenate([g[0], g[1], g[2]]), np.concatenate([g[0], g[2], g[3]])])
    tris = np.concatenate([torus, ground])
    mid = np.concatenate([np.zeros(len(torus), np.int32), np.ones(2, np.int32)])
    mats = [dict(type="roughdielectric", specularReflectance=(1, 1, 1), specularTransmittance=(0.9, 0.95, 1.0), alpha=0.15, intIOR=1.5, distribution="ggx"),
            dict(type="diffuse", reflectance=(0.6, 0.55, 0.5))]
    c, s_ = np.cos(0.4), np.sin(0.4)
    to_world = np.array([[c, -s_, 0, 0], [s_, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)      # core/transform.cc:136-165, axis z
    return dict(tris=tris, material_id=mid, light_id=np.full(len(tris), -1, np.int32), materials=mats, env=envscene_envmap(), env_to_world=to_world,
                env_scale=1.5, env_radius=6.0, camera=ENV_CAMERA, filter="tent")


def write_envscene(dirpath, width=128, height=128, spp=64, max_depth=8, name="envtorus", nu=96, nv=48):
    """Torus (PLY, rough dielectric) over a diffuse ground quad (OBJ) lit ONLY by an environment map
    (BASELINE config 4 in miniature: env lighting + PLY mesh + roughdielectric). Returns the XML path."""
    os.makedirs(dirpath, exist_ok=True)
    v, f = torus_mesh(nu, nv, R=0.6, r=0.25, bump=0.0)
    write_ply(os.path.join(dirpath, name + "_torus.ply"), v, f)
    with open(os.path.join(dirpath, name + "_ground.obj"), "w") as fo:
        fo.write("o ground\nv -2 -2 -0.3\nv 2 -2 -0.3\nv 2 2 -0.3\nv -2 2 -0.3\nf 1 2 3\nf 1 3 4\n")
    write_hdr(os.path.join(dirpath, name + "_env.hdr"), synthetic_envmap())
    x = """<?xml version="1.0" encoding="utf-8"?>
<scene version="0.5.0">
  <integrator type="path">
    <integer name="maxDepth" value="%d"/>
  </integrator>
  <sensor type="perspective">
    <float name="fov" value="40"/>
    <transform name="toWorld">
      <lookAt origin="0.3, -2.6, 1.5" target="0, 0, 0" up="0, 0, 1"/>
    </transform>
    <sampler type="independent">
      <integer name="sampleCount" value="%d"/>
    </sampler>
    <film type="hdrfilm">
      <integer name="width" value="%d"/>
      <integer name="height" value="%d"/>
      <rfilter type="tent"/>
    </film>
  </sensor>
  <emitter type="envmap">
    <string name="filename" value="%s_env.hdr"/>
    <float name="worldRadius" value="6.0"/>
    <float name="scale" value="1.5"/>
    <transform name="toWorld">
      <rotate x="0" y="0" z="1" angle="0.4"/>
    </transform>
  </emitter>
  <bsdf type="roughdielectric" id="frosted">
    <rgb name="specularReflectance" value="1, 1, 1"/>
    <rgb name="specularTransmittance" value="0.9, 0.95, 1"/>
    <float name="alpha" value="0.15"/>
    <float name="intIOR" value="1.5"/>
    <string name="distribution" value="ggx"/>
  </bsdf>
  <bsdf type="diffuse" id="ground">
    <rgb name="reflectance" value="0.6, 0.55, 0.5"/>
  </bsdf>
  <shape type="ply">
    <string name="filename" value="%s_torus.ply"/>
    <ref id="frosted"/>
  </shape>
  <shape type="obj">
    <string name="filename" value="%s_ground.obj"/>
    <ref id="ground"/>
  </shape>
</scene>
""" % (max_depth, spp, width, height, name, name, name)
    path = os.path.join(dirpath, name + ".xml")
    with open(path, "w") as fo:
        fo.write(x)
    return path
