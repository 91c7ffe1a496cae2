#!/usr/bin/env python
"""bench.py - the driver contract.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): raw ray-cast
benchmark on the synthetic 1,000,000-triangle torus; one "step" = one closest-hit pass over
16,777,216 incoherent rays (uniform origins in the inflated AABB, uniform directions).  The primary
ray set and the any-hit pass are timed once per run and reported beside it.

extra.render_c3 / render_c4 (every N): BASELINE configs[2] and [3] AS SPECIFIED -- Cornell box 1920x1080, depth 16,
          1024 spp (diffuse) / 256 spp (glossy + dielectric) IN TOTAL, sample indices partitioned over the N GPUs
          (strong scaling), one NCCL reduce of the film to rank 0 inside the timed frame; samples/s, the reduce's
          device time, a bytes-per-sample roofline and (N = 1) the reference renderer timed beside it.
value   : Mrays/s, rays resident in HBM, CUDA events on the launching stream, max over ranks.
e2e     : the same pass through the host-buffer C-ABI call (pinned host rays in, hit records out,
          both copies inside the timed region, pipelined in chunks by the library).
roofline: HBM bound; achieved = algorithmic bytes per ray (tests/golden/visits_c2.json)
          x rays / kernel time, against MEASURED_PEAKS.json.
cpu_baseline: the UNMODIFIED reference (oracle/_ref/raycast_ref, kind "reference") on a bounded
          strided sample of the same rays, all host threads; the same run also checks the GPU's
          (prim, t) for that sample against the reference's output.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 1 << 24
TORUS = (1000, 500)
METRIC = "Mrays/s (incoherent closest-hit, 1M-triangle mesh)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs: NVML (nvidia_ml_py)
    every 2 ms when it loads, else one nvidia-smi query per 100 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index = index
        self.rows = []          # (sm_mhz, sm_max_mhz, set(reasons))
        self.stop = False
        self.th = None
        self.source = "nvidia-smi"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.source = "nvml"
        except Exception:
            self.nv = None

    def _run(self):
        while not self.stop:
            try:
                if self.nv is not None:
                    sm = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    try:
                        bits = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        bits = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((sm, self.mx, {n for b, n in self.NVML_REASONS if bits & b}))
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in out.strip().split(",")]
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                self.rows.append((float(r[0]), float(r[1]), {nme for k, nme in enumerate(names) if r[3 + k].lower().startswith("active")}))
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = max([r[1] for r in self.rows], default=0.0)
        reasons = set()
        for r in self.rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def visits():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "visits_c2.json")))


def make_scene():
    from spica_b200 import scenes
    v, f = scenes.torus_mesh(*TORUS)
    return v, f, scenes.mesh_triangles(v, f)


def quiet_stdout():
    """The C++ host logs to stdout like the reference CLI; bench.py's stdout carries exactly one JSON line."""
    class Q:
        def __enter__(self):
            sys.stdout.flush()
            self.saved = os.dup(1)
            self.devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(self.devnull, 1)
        def __exit__(self, *a):
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved); os.close(self.devnull)
    return Q()


def reference_render(xml_small, pixels, spp_small, out_prefix):
    """`spica -i scene.xml -t <all host threads>` of the UNMODIFIED reference on a reduced sample count; samples/s."""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bin")
    if not os.path.exists(os.path.join(ref_bin, "spica")):
        return None, None
    from spica_b200 import scenes
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    subprocess.run(["./spica", "-i", xml_small, "-t", str(cores), "-o", out_prefix], cwd=ref_bin, check=True, stdout=subprocess.DEVNULL, timeout=900)
    dt = time.perf_counter() - t0
    return ({"value": pixels * spp_small / dt * 1e-6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
             "sample": "%d spp of the same scene file through the reference CLI (every pass is identical work, core/integrator.cc:64); "
                       "includes its scene load and per-pass .hdr save; %.1f s" % (spp_small, dt)},
            scenes.read_hdr(out_prefix + ".hdr"))


def cornell_c1_side_by_side():
    """BASELINE configs[0] through the reference-facing C++ host (scene file -> plugin surface -> C ABI -> .hdr file)."""
    import tempfile
    from spica_b200 import host, scenes
    d = tempfile.mkdtemp(prefix="spb_c1_")
    out = {"config": "Cornell box 512x512, 64 spp, max depth 8 (BASELINE configs[0])"}
    xml = scenes.write_cornell(d, 512, 512, 64, 8)
    secs = []
    with quiet_stdout():
        for w in range(2):                                                    # first use of the host library in this process, and the
            host.render_scene(xml, os.path.join(d, "warm"), seed=1)           # GPU back at its clocks after the CPU-only minutes before
        for rep in range(3):
            t0 = time.perf_counter()
            img = host.render_scene(xml, os.path.join(d, "gpu"), seed=1 + rep)
            secs.append(time.perf_counter() - t0)
    out["gpu_seconds_scene_file_to_image"] = min(secs)
    out["gpu_seconds_all_runs"] = secs
    out["gpu_msamples_s"] = 512 * 512 * 64 / min(secs) * 1e-6
    out["note"] = "each run: parse the scene file, create a context, build the BVH, render 64 spp, encode and write the .hdr"
    xml4 = scenes.write_cornell(d, 512, 512, 4, 8, name="cornell4")
    cpu, ref = reference_render(xml4, 512 * 512, 4, os.path.join(d, "cpu"))
    if cpu:
        out.update({"cpu_baseline": cpu, "mean_radiance_gpu": float(img.mean()), "mean_radiance_cpu_4spp": float(ref.mean())})
    return out


RENDER_CONFIGS = {
    "c3": dict(index=2, kind="cornell", variant="diffuse", W=1920, H=1080, depth=16, what="Cornell box (diffuse + area light)"),
    "c4": dict(index=3, kind="cornell", variant="glossy", W=1920, H=1080, depth=16, what="Cornell box with glossy microfacet + dielectric surfaces"),
    "c5": dict(index=4, kind="env", nu=2500, nv=2000, W=3840, H=2160, depth=16, what="10,000,002-triangle torus + ground under an environment map (rough dielectric)"),
}


def render_frame(torch, dist, capi, partition, scenes, local, rank, world, name, spp_total, comm_id, cpu_spp, film_reduce="ipc"):
    """One BASELINE render configuration, timed as one frame: every rank renders its share of the sample indices,
    then ONE NCCL reduce of the RGBW film to rank 0.  Wall clock between barriers, max over ranks."""
    cfg = RENDER_CONFIGS[name]
    W, H, depth = cfg["W"], cfg["H"], cfg["depth"]
    rctx = capi.Context(local)
    if world > 1:
        rctx.comm_init(comm_id, world, rank)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        rctx.sync()

    # warm-up: scene upload, BVH build (on the device), kernels, queues, and the communicator's first collective
    t_setup = time.perf_counter()
    if cfg["kind"] == "cornell":
        cam = scenes.CORNELL_CAMERA
        begin_kw = {}
        capi.cornell_render(rctx, W, H, 2, max_depth=depth, seed=7, variant=cfg["variant"], first=rank, stride=world)
    else:
        cam = scenes.ENV_CAMERA
        begin_kw = {"filter": "tent"}
        capi.envscene_render(rctx, W, H, 1, max_depth=depth, nu=cfg["nu"], nv=cfg["nv"], seed=7, first=rank, stride=world)
    setup_s = time.perf_counter() - t_setup
    bst = rctx.stats()
    c2w, r2c = scenes.perspective_camera(scenes.look_at(cam["origin"], cam["target"], cam["up"]), cam["fov"], W, H)
    if world > 1:
        rctx.film_reduce(0)
    rctx.render_begin(W, H, c2w, r2c, max_depth=depth, seed=7, **begin_kw)
    # How the films meet on rank 0: by default rank 0 maps the other ranks' films through CUDA IPC (once, here) and sums them
    # with one kernel over NVLink peer memory (spb_film_reduce_imported) behind a barrier of the job; --film-reduce nccl, or
    # GPUs that cannot map each other, take one ncclReduce (spb_film_reduce) instead
    how = "nccl"
    if world > 1 and film_reduce == "ipc":
        hs = [None] * world
        dist.all_gather_object(hs, rctx.film_export_handle())
        ok = 1
        if rank == 0:
            try:
                rctx.film_import_handles(hs[1:])
                rctx.film_reduce_imported()             # (films are all zero here: warms the kernel and the mappings)
            except capi.SpbError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.broadcast(flag, 0)
        how = "ipc" if int(flag.item()) else "nccl"
    sync()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        rctx.render_samples(*partition.sample_partition(spp_total, rank, world))
        if world > 1 and how == "ipc":
            dist.barrier()                              # every rank's samples are in its film
            if rank == 0:
                rctx.film_reduce_imported()
        elif world > 1:
            rctx.film_reduce(0)
        sync()
        dt = time.perf_counter() - t0
    st = rctx.render_stats()
    t = torch.tensor([dt, st["render_ms"], st["reduce_ms"], -st["render_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec, render_ms_max, reduce_ms_max, render_ms_min = float(t[0]), float(t[1]), float(t[2]), -float(t[3])
    rays = st["rays_closest"] + st["rays_shadow"] + st["rays_mis"]
    out = None
    if rank == 0:
        img = rctx.film_resolve()
        film = rctx.film_read()
        P = max(st["paths"], 1)
        # algorithmic bytes of the streaming loop per sample ("render roofline"): per closest-hit ray 32 B ray in +
        # 16 B hit out (extend), 32 + 32 + 16 B in (shade) and 64 B ray + state written by the kernel that produced it; per
        # shadow / MIS ray a 48 B record written and read; 16 B film reduction per path end, unoccluded light sample and MIS hit
        b_sample = (st["rays_closest"] * (48 + 80 + 64) + (st["rays_shadow"] + st["rays_mis"]) * (96 + 16) + st["paths"] * 16) / P
        peak, peak_src = peaks()
        achieved = b_sample * W * H * spp_total / sec * 1e-9 / world
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("render_dram_bytes_per_sample", {}).get(name)
        out = {"config": "BASELINE configs[%d]: %s, %dx%d, %d spp total, depth %d, spp partitioned over %d GPU(s), NCCL film reduce to rank 0 inside the frame"
                         % (cfg["index"], cfg["what"], W, H, spp_total, depth, world),
               "triangles": bst["n_tris"], "triangle_records": "float32 (48 B)" if bst["tri_format"] == 0 else "float64 (80 B) + rounded float32 (48 B)",
               "bvh_build_s": bst["build_seconds"], "setup_s_rank0": setup_s,
               "scaling": "strong", "spp_total": spp_total, "spp_rank0": partition.sample_partition(spp_total, 0, world)[1],
               "seconds": sec, "msamples_s": W * H * spp_total / sec * 1e-6,
               "allreduce_ms": reduce_ms_max,
               "allreduce": ("ncclReduce(sum, f32, root 0) of the %d MB RGBW film; device time incl. waiting for the slowest rank, max over ranks" if how == "nccl" else
                             "one kernel on rank 0 that reads the other ranks' %d MB RGBW films through CUDA-IPC peer mappings (NVLink) and adds them to its own, "
                             "behind a barrier of the job (inside the frame's time, not inside allreduce_ms)") % (W * H * 16 >> 20),
               "render_ms_slowest_rank": render_ms_max, "render_ms_fastest_rank": render_ms_min,
               "rank0": {"rays": rays, "rays_per_sample": rays / P, "mrays_s": rays / (st["render_ms"] * 1e-3) * 1e-6,
                         "iterations": st["iterations"], "kernel_launches": st["kernel_launches"]},
               "film_weight_per_pixel": float(film[..., 3].mean()), "mean_radiance": float(img.mean()), "clocks": clk.summary(),
               "roofline": {"bound": "hbm", "bytes_per_sample": b_sample, "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                            "note": "per GPU; queue records only -- BVH nodes and triangles come out of L2"}}
        if cfg["kind"] == "env":
            gp = os.path.join(ROOT, "tests", "golden", "baseline_c5_small_ref.npz")
            if os.path.exists(gp):
                g = np.load(gp)
                ref_mean = float(g["runs"].astype(np.float64).mean())
                out["mean_radiance_reference"] = ref_mean
                out["mean_radiance_ratio"] = float(img.mean()) / ref_mean
                out["mean_radiance_reference_note"] = ("mean of 4 runs of the unmodified reference on the same scene with 100 k triangles at 480x270, 64 spp "
                                                       "(tests/golden/baseline_c5_small_ref.npz; the image itself is relMSE-checked in tests/test_render_gpu.py)")
        if world == 1 and cpu_spp > 0 and cfg["kind"] == "cornell":
            import tempfile
            d = tempfile.mkdtemp(prefix="spb_%s_" % name)
            xml = scenes.write_cornell(d, W, H, cpu_spp, depth, variant=cfg["variant"], name=name)
            try:
                cpu, ref = reference_render(xml, W * H, cpu_spp, os.path.join(d, "cpu"))
                if cpu:
                    out["cpu_baseline"] = cpu
                    out["mean_radiance_cpu_%dspp" % cpu_spp] = float(ref.mean())
            except Exception as exc:      # noqa: BLE001
                out["cpu_baseline"] = {"error": repr(exc)}
    if world > 1:                   # the mappings go before the films they map do
        if rank == 0 and how == "ipc":
            rctx.film_import_handles([])
        dist.barrier()
    rctx.close()
    return out


def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref), all host threads, on a
    bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    from spica_b200 import scenes
    v, f, tris = make_scene()
    cores = os.cpu_count() or 1
    sample = args.ref_rays
    stride = N_RAYS // sample
    rays = scenes.incoherent_rays(N_RAYS, v.min(0), v.max(0), seed=2)[::stride]
    kind = "reference" if ob.have_ref() else "port"
    if kind == "reference":
        # one process: BVH built once, then W untimed + K timed passes of BVHAccel::intersect
        info, _, _ = ob.ref_raycast(rays, tris=tris, threads=cores, repeat=args.steps, warmup=args.warmup)
        ms = 1e3 * info["trace_s"]
    else:
        nodes = ob.bvh_build(tris)
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter(); ob.trace_closest(nodes, tris, rays, threads=cores); dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
    val = len(rays) / (ms * 1e-3) * 1e-6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "raycast 1M-triangle torus, incoherent closest-hit", "triangles": len(tris),
                       "rays_per_step": len(rays), "note": "bounded strided sample of the 16,777,216-ray set per step"},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "sample": "%d of %d incoherent rays (stride %d), BVHAccel::intersect on %d threads" % (len(rays), N_RAYS, stride, cores)},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def bind_to_gpu_numa(local):
    """Multi-GPU runs: keep this rank (and so the first touch of its pinned host buffers) on the CPU socket its GPU
    hangs off, when the process is allowed to run there.  Returns a short note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        node = int(open("/sys/bus/pci/devices/%04x:%s/numa_node" % (int(dom, 16), rest.lower())).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "gpu on numa node %d, none of its cpus in this process's cpuset" % node
        os.sched_setaffinity(0, allowed)
        return "bound to numa node %d (%d cpus)" % (node, len(allowed))
    except Exception as e:      # noqa: BLE001
        return "not bound (%s)" % type(e).__name__


_REAL_STDOUT = None


def claim_stdout():
    """bench.py's stdout carries exactly ONE line, the JSON record.  Libraries below us print there too (NCCL's version
    banner, the C++ host's progress lines): from here on fd 1 is stderr, and emit() writes the record to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--ref-rays", type=int, default=1 << 21)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--max-leaf", type=int, default=0, help="triangles per leaf, 1..3 (0 = library default, 3)")
    ap.add_argument("--c3-spp", type=int, default=1024, help="total spp of the C3 frame (BASELINE configs[2]: 1024; 0 = skip)")
    ap.add_argument("--c4-spp", type=int, default=256, help="total spp of the C4 frame (BASELINE configs[3]: 256; 0 = skip)")
    ap.add_argument("--c5-spp", type=int, default=-1, help="total spp of the C5 frame (BASELINE configs[4]: 256 on 8 GPUs; -1 = 256 when N = 8, else skip)")
    ap.add_argument("--cpu-render-spp", type=int, default=2, help="spp of the reference renderer's run beside C3 / C4 at N = 1 (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--film-reduce", choices=["ipc", "nccl"], default="ipc", help="N > 1: how the films meet on rank 0 (ipc = a kernel over CUDA-IPC peer mappings, falls back to nccl)")
    ap.add_argument("--chunk", type=int, default=0, help="rays per pipelined chunk of the host-buffer call (0 = library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from spica_b200 import capi, partition, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa(local) if world > 1 else "single process, not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    v, f, tris = make_scene()
    lo, hi = v.min(0), v.max(0)
    ctx = capi.Context(local)
    ctx.set_triangles(tris)
    builders = None
    if world == 1:
        # the same mesh through the HOST builder first (the r01 default): its build time, its trace rate and its bytes, to put
        # beside the device builder's below ("C2 Mrays/s on the GPU-built tree >= 95 % of the host-SAH tree")
        probe = scenes.incoherent_rays(1 << 22, lo, hi, seed=2)
        d_probe = torch.from_numpy(probe).cuda(); d_ph = torch.empty((len(probe), 4), dtype=torch.float32, device="cuda")

        def probe_rate():
            ms = []
            for _ in range(4):
                ctx.trace_closest_dev(d_probe, len(probe), d_ph)
                ms.append(ctx.counters()["last_kernel_ms"])
            return len(probe) / min(ms[1:]) * 1e-3

        ctx.build(max_leaf_tris=args.max_leaf, builder=2)
        hs = ctx.stats()
        host_rate = probe_rate()
        host_bytes = tuple(bytes(x) for x in ctx.export_bvh()[1:])
        ctx.build(max_leaf_tris=args.max_leaf)
        first_build_s = ctx.stats()["build_seconds"]                    # the process's first device build: kernels loaded, arena allocated
        ctx.build(max_leaf_tris=args.max_leaf)
        ds = ctx.stats()
        builders = {"device_sah": {"build_s": ds["build_seconds"], "first_build_s": first_build_s, "wide_nodes": ds["n_wide_nodes"], "sah_cost": ds["sah_cost"], "probe_mrays_s": probe_rate()},
                    "host_sah": {"build_s": hs["build_seconds"], "wide_nodes": hs["n_wide_nodes"], "sah_cost": hs["sah_cost"], "probe_mrays_s": host_rate, "host_threads": os.cpu_count()},
                    "wide_bvh_bytes_equal": tuple(bytes(x) for x in ctx.export_bvh()[1:]) == host_bytes,
                    "probe": "%d incoherent closest-hit rays, best of 3 launches" % len(probe)}
        builders["device_vs_host_rate"] = builders["device_sah"]["probe_mrays_s"] / host_rate
        del d_probe, d_ph
    else:
        ctx.build(max_leaf_tris=args.max_leaf)
    if args.variant >= 0:
        ctx.set_option("trace_variant", args.variant)
    if args.chunk > 0:
        ctx.set_option("chunk_rays", args.chunk)
    st = ctx.stats()

    n = args.rays
    # weak scaling: every rank traces its own n rays of the same distribution (distinct counters)
    rays = scenes.incoherent_rays(n, lo, hi, seed=2, start=partition.ray_shard(n, rank)[0])
    pin_rays = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
    pin_rays.numpy()[:] = rays
    pin_hits = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_rays.copy_(pin_rays)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- device-resident pass: `value`
    for _ in range(args.warmup):
        ctx.trace_closest_dev(d_rays, n, d_hits)
    launches0 = ctx.counters()["kernel_launches"]
    barrier()
    kernel_ms = []
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            ctx.trace_closest_dev(d_rays, n, d_hits)      # synchronous; timed by CUDA events on its stream
            kernel_ms.append(ctx.counters()["last_kernel_ms"])
    barrier()
    launches = ctx.counters()["kernel_launches"] - launches0
    total_ms = float(np.sum(kernel_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * n / (ms_per_step * 1e-3) * 1e-6

    # ---- end to end: pinned host rays -> hits back on the host, through the public C-ABI call
    for _ in range(2):
        ctx.trace_closest(pin_rays, out=pin_hits)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.trace_closest(pin_rays, out=pin_hits)
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * n / float(t.item()) * 1e-6
    e2e_s_max = float(t.item())

    # ---- the ceiling of that call: the same bytes over PCIe with no kernel at all -- H2D of the rays and D2H of the hit
    # records on two streams at once, on all ranks at once (with several GPUs the host-memory path is shared)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    def pcie_pass():
        with torch.cuda.stream(s_in):
            d_rays.copy_(pin_rays, non_blocking=True)
        with torch.cuda.stream(s_out):
            pin_hits.copy_(d_hits, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    pcie_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        pcie_pass()
    pcie_s = (time.perf_counter() - t0) / 3
    t = torch.tensor([pcie_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pcie_s_max = float(t.item())

    # ---- BASELINE configs[2] and [3] on all ranks: the path tracer, 1920x1080, depth 16, the configuration's TOTAL sample
    # count partitioned over the GPUs (strong scaling), one NCCL reduce of the film per frame inside the timed region
    renders = {}
    c5_spp = args.c5_spp if args.c5_spp >= 0 else (256 if world == 8 else 0)
    for name, spp_total in (("c3", args.c3_spp), ("c4", args.c4_spp), ("c5", c5_spp)):
        if spp_total > 0:
            comm_id = None
            if world > 1:                       # a fresh communicator per configuration (the context owns and destroys it)
                ids = [capi.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                comm_id = ids[0]
            renders["render_" + name] = render_frame(torch, dist, capi, partition, scenes, local, rank, world, name, spp_total,
                                                     comm_id, 0 if args.no_cpu_baseline else args.cpu_render_spp, args.film_reduce)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- side measurements on rank 0 (not part of `value`)
    extra = dict(renders)
    extra["builders"] = builders
    pri = scenes.primary_rays(4096, 4096)[: n]
    d_rays.copy_(torch.from_numpy(pri)); torch.cuda.synchronize()
    for _ in range(2):
        ctx.trace_closest_dev(d_rays, len(pri), d_hits)
    extra["primary_mrays_s"] = len(pri) / (ctx.counters()["last_kernel_ms"] * 1e-3) * 1e-6
    anyr = scenes.incoherent_rays(n, lo, hi, seed=2, anyhit=True)
    d_rays.copy_(torch.from_numpy(anyr)); torch.cuda.synchronize()
    d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ctx.trace_any_dev(d_rays, n, d_occ)
    extra["anyhit_mrays_s"] = n / (ctx.counters()["last_kernel_ms"] * 1e-3) * 1e-6

    # ---- BASELINE configs[0] (Cornell box 512x512, 64 spp, depth 8): GPU through the reference-facing host
    # (scene file -> C++ plugin surface -> C ABI) next to the reference's own CPU renderer on a bounded sample
    if world == 1 and not args.no_cpu_baseline:
        try:
            extra["cornell_c1"] = cornell_c1_side_by_side()
        except Exception as exc:                # a side measurement must not sink the headline number
            extra["cornell_c1"] = {"error": repr(exc)}

    # ---- roofline (dominant kernel = the closest-hit traversal kernel; one launch per step)
    vis = visits()["incoherent"]
    tri_bytes = 48 if st["tri_format"] == 0 else 96
    b_ray = 32 + 16 + vis["V_int"] * 64 + vis["V_leaf"] * tri_bytes
    peak, peak_src = peaks()
    per_launch_ms = float(np.mean(kernel_ms))
    achieved = b_ray * n / (per_launch_ms * 1e-3) * 1e-9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": peak_src, "bytes_per_ray": b_ray,
            "V_int": vis["V_int"], "V_leaf": vis["V_leaf"], "kernel": "traceCoopPairKernel (closest, trace_variant 5)",
            "kernel_ms": per_launch_ms}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        roof["traffic"] = tj.get("dram_bytes_per_launch")
        # The walk is bound by issue slots, not by bytes (the tree lives in L2; DRAM sees the rays and the hits): the
        # honest ceiling is warp instructions per ray (ncu, same kernel and workload) against the issue slots of the
        # machine, 4 schedulers x SMs x the SM clock observed during the timed region.
        wi = tj.get("warp_inst_per_launch")
        sm_mhz = clk.summary().get("sm_mhz")
        if wi and sm_mhz:
            slots = torch.cuda.get_device_properties(local).multi_processor_count * 4 * sm_mhz * 1e6
            roof["instruction"] = {"warp_inst_per_ray": wi / tj.get("rays_per_launch", n), "issue_slots_per_s": slots,
                                   "achieved_warp_inst_per_s": wi / tj.get("rays_per_launch", n) * n / (per_launch_ms * 1e-3),
                                   "frac": wi / tj.get("rays_per_launch", n) * n / (per_launch_ms * 1e-3) / slots,
                                   "source": tj.get("source")}

    # ---- CPU baseline + parity gates on strided subsets of all three C2 ray sets (rank 0, N = 1 only): the UNMODIFIED
    # reference's BVHAccel::intersect answers the same rays; ids must be equal, except for rays whose two candidate
    # triangles are hit at EXACTLY the same double t (own-built tree: the lower index wins; reference: its visit order) --
    # those are counted and listed, not hidden
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as ob
        cores = os.cpu_count() or 1
        have_ref = ob.have_ref()
        nodes = None if have_ref else ob.bvh_build(tris)

        def ref_closest(sub):
            if have_ref:
                info, p, t = ob.ref_raycast(sub, tris=tris, threads=cores)
                return info["mrays_s"], p, t
            t0 = time.perf_counter(); p, t, _, _ = ob.trace_closest(nodes, tris, sub, threads=cores)
            return len(sub) / (time.perf_counter() - t0) * 1e-6, p, t

        def ref_any(sub):
            if have_ref:
                return ob.ref_raycast(sub, mode="any", tris=tris, threads=cores)[1]
            return ob.trace_any(nodes, tris, sub, threads=cores)

        def closest_gate(sub):
            rate, prim_ref, t_ref = ref_closest(sub)
            hits = ctx.trace_closest(sub)
            diff = np.nonzero(hits["prim"] != prim_ref)[0]
            ties = []
            if len(diff):
                h64 = ctx.trace_closest(np.ascontiguousarray(sub[diff].astype(np.float64)))     # exact double t of the GPU's winner
                ties = [int(i) for k, i in enumerate(diff) if prim_ref[i] >= 0 and h64["prim"][k] >= 0 and h64["t"][k] == t_ref[i]]
            h = (prim_ref >= 0) & (hits["prim"] == prim_ref)
            rel = np.abs(hits["t"][h].astype(np.float64) - t_ref[h]) / t_ref[h]
            return rate, {"checked_rays": len(sub), "prim_id_mismatches": int(len(diff) - len(ties)), "exact_tie_rays": len(ties),
                          "exact_tie_ray_indices": ties[:16], "max_rel_t_err": float(rel.max()) if h.any() else 0.0}

        sample = min(args.ref_rays, n)
        stride = n // sample
        cpu_val, par_inc = closest_gate(np.ascontiguousarray(rays[::stride]))
        kind = "reference" if have_ref else "port"
        cpu = {"value": cpu_val, "unit": "Mrays/s", "cores": cores, "kind": kind,
               "sample": "%d of %d incoherent rays (stride %d), BVHAccel::intersect on %d threads" % (sample, n, stride, cores)}
        sub_n = min(sample, 1 << 20)
        _, par_pri = closest_gate(np.ascontiguousarray(pri[:: max(1, len(pri) // sub_n)]))
        any_sub = np.ascontiguousarray(anyr[:: max(1, n // sub_n)])
        occ_ref = ref_any(any_sub)
        occ = ctx.trace_any(any_sub)
        par_any = {"checked_rays": len(any_sub), "flag_mismatches": int((occ != occ_ref).sum()), "occluded_fraction": float(occ_ref.mean())}
        parity = {"closest_incoherent": par_inc, "closest_primary": par_pri, "any_hit": par_any,
                  "oracle": kind, "tree": "own-built (exact-t ties -> lower primitive index)",
                  # the r01 keys, for continuity: the incoherent closest-hit gate
                  "checked_rays": par_inc["checked_rays"], "prim_id_mismatches": par_inc["prim_id_mismatches"], "max_rel_t_err": par_inc["max_rel_t_err"]}

    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "raycast 1M-triangle torus, 16,777,216 incoherent rays/GPU/step, closest-hit",
                       "triangles": len(tris), "rays_per_step_per_gpu": n, "bvh": "8-wide compressed, %d nodes, %d B nodes + %d B triangles" % (st["n_wide_nodes"], st["node_bytes"], st["tri_bytes"]),
                       "l2": "streamed inputs+outputs (%d MB/step) exceed the 126 MB L2; the BVH is meant to stay resident" % ((n * 48) >> 20),
                       "parallelism": "rays sharded over %d GPU(s), scene replicated, no collective" % world,
                       "host_numa": "rank 0: " + numa_note},
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 16,
                    "pcie_ceiling": {"value": world * n / pcie_s_max * 1e-6, "unit": "Mrays/s", "frac": pcie_s_max / e2e_s_max,
                                     "what": "the same pinned buffers copied H2D (rays) and D2H (hit records) concurrently with no kernel, all ranks at once, max over ranks",
                                     "h2d_gb_s_per_gpu": n * 32 / pcie_s_max * 1e-9, "d2h_gb_s_per_gpu": n * 16 / pcie_s_max * 1e-9}},
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu,
            "parity": parity, "extra": extra, "bvh_build_s": st["build_seconds"],
            "bvh_builder": {0: "device binned SAH + device 8-wide collapse", 1: "device LBVH", 2: "host binned SAH"}.get(st["builder"], "adopted")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
