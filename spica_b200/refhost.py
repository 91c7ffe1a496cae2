"""Drives the UNMODIFIED reference host (`spica -i scene.xml`, any build of it) with this
repo's plugins/path.so and plugins/bvh.so swapped in (spica_b200/refplugin).  The reference loads
plugins from ./plugins relative to its working directory (core/cobject.cc:12-13), so a run
directory is assembled whose plugins/ holds the reference's own film / sampler / camera / bsdf /
emitter plugins plus the two GPU ones.  Plumbing for tests and INTEGRATION.md's recipe."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(HERE, "lib", "refplugins")
SHIM_PLUGINS = ("path", "bvh", "directlighting")


def available(ref_root):
    return all(os.path.exists(os.path.join(SHIM_DIR, p + ".so")) for p in SHIM_PLUGINS) and \
        os.access(os.path.join(ref_root, "bin", "spica"), os.X_OK)


def make_run_dir(run_dir, ref_root, gpu_plugins=SHIM_PLUGINS):
    """ref_root: an install of the reference (bin/spica, bin/plugins/*.so, libspica_core.so)."""
    pdir = os.path.join(run_dir, "plugins")
    os.makedirs(pdir, exist_ok=True)
    src = os.path.join(ref_root, "bin", "plugins")
    for f in sorted(os.listdir(src)):
        dst = os.path.join(pdir, f)
        if os.path.lexists(dst):
            os.remove(dst)
        name = f[:-3]
        os.symlink(os.path.join(SHIM_DIR, f) if name in gpu_plugins else os.path.join(src, f), dst)
    for name in gpu_plugins:            # GPU plugins the reference build does not have
        dst = os.path.join(pdir, name + ".so")
        if not os.path.lexists(dst):
            os.symlink(os.path.join(SHIM_DIR, name + ".so"), dst)
    return run_dir


def run(xml_path, output_prefix, run_dir, ref_root, threads=1, env=None, gpu_plugins=SHIM_PLUGINS):
    make_run_dir(run_dir, ref_root, gpu_plugins)
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(HERE, "lib"), ref_root, e.get("LD_LIBRARY_PATH", "")])
    e.update({k: str(v) for k, v in (env or {}).items()})
    return subprocess.run([os.path.join(ref_root, "bin", "spica"), "-i", xml_path, "-t", str(threads), "-o", output_prefix],
                          cwd=run_dir, env=e, capture_output=True, text=True)


def read_dump(path):
    """Parses the file written under SPICA_B200_DUMP_SCENE (refplugin/path_plugin.cc)."""
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype=np.int64, count=8)
    n, nm, nl, any_n, any_uv, spp, ew, eh = (int(x) for x in hdr)
    off = 64

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a
    out = dict(n_triangles=n, any_normals=bool(any_n), any_uv=bool(any_uv), sample_count=spp)
    out["verts"] = take(np.float64, n * 9).reshape(n, 9)
    out["normals"] = take(np.float32, n * 9).reshape(n, 9)
    out["uvs"] = take(np.float32, n * 6).reshape(n, 6)
    out["material_id"] = take(np.int32, n)
    out["light_id"] = take(np.int32, n)
    mats = (capi.Material * nm).from_buffer_copy(raw, off); off += C.sizeof(capi.Material) * nm
    lights = (capi.Light * nl).from_buffer_copy(raw, off); off += C.sizeof(capi.Light) * nl
    desc = capi.RenderDesc.from_buffer_copy(raw, off); off += C.sizeof(capi.RenderDesc)
    out["materials"] = [dict(type=m.type, distribution=m.distribution, kr=list(m.kr), kt=list(m.kt), eta=list(m.eta), k=list(m.k),
                             alpha_u=m.alpha_u, alpha_v=m.alpha_v) for m in mats]
    out["lights"] = [dict(type=l.type, prim=l.prim, radiance=list(l.radiance)) for l in lights]
    out["desc"] = {k: (np.array(getattr(desc, k)) if hasattr(getattr(desc, k), "__len__") else getattr(desc, k)) for k, _ in desc._fields_}
    if ew:
        out["env_light_to_world"] = take(np.float64, 16).reshape(4, 4)
        tail = take(np.float64, 4)
        out["env_center"], out["env_radius"] = tail[:3], float(tail[3])
        out["env_rgb"] = take(np.float32, ew * eh * 3).reshape(eh, ew, 3)
    nt, ntex = (int(x) for x in take(np.int64, 2))
    texs = (capi.Texture * max(nt, 1)).from_buffer_copy(raw[off:off + C.sizeof(capi.Texture) * max(nt, 1)].ljust(C.sizeof(capi.Texture) * max(nt, 1), b"\0"))
    off += C.sizeof(capi.Texture) * nt
    out["textures"] = [dict(type=t.type, width=t.width, height=t.height, texel_offset=t.texel_offset, color0=list(t.color0), color1=list(t.color1),
                            uoffset=t.uoffset, voffset=t.voffset, uscale=t.uscale, vscale=t.vscale) for t in texs[:nt]]
    out["material_textures"] = take(np.int32, 2 * nm).reshape(nm, 2)
    out["texels"] = take(np.float32, ntex * 3).reshape(ntex, 3)
    assert off == len(raw), (off, len(raw))
    return out
