"""bench.py's host-side helpers on a machine without a GPU: they must degrade to notes, not raise."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_numa_binding_is_best_effort():
    b = _bench()
    before = os.sched_getaffinity(0)
    note = b.bind_to_gpu_numa(0)
    assert isinstance(note, str) and note
    assert os.sched_getaffinity(0) <= before          # never widens the affinity, never empties it
    assert os.sched_getaffinity(0)


def test_clock_sampler_without_samples():
    b = _bench()
    with b.ClockSampler(0) as clk:
        pass
    s = clk.summary()
    assert set(s) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert s["source"] in ("nvml", "nvidia-smi") and isinstance(s["reasons"], list)


def test_visit_constants_of_the_roofline():
    b = _bench()
    v = b.visits()["incoherent"]
    # B_ray = 32 + 16 + V_int * 64 + V_leaf * 48 with the oracle-measured constants
    assert 20 < v["V_int"] < 60 and 2 < v["V_leaf"] < 10
    assert abs(32 + 16 + v["V_int"] * 64 + v["V_leaf"] * 48 - v["bytes_per_ray_f32verts"]) < 1e-6
    assert abs(v["bytes_per_ray_f32verts"] - 2326.51) < 0.01


def test_stdout_carries_exactly_the_json_line(tmp_path):
    """claim_stdout() + emit(): whatever libraries print to fd 1 afterwards (NCCL's banner, the C++ host's progress lines) goes to
    stderr; stdout gets the one record."""
    import json
    import subprocess
    import sys
    code = ("import importlib.util, os, sys\n"
            "spec = importlib.util.spec_from_file_location('b', %r); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
            "b.claim_stdout()\n"
            "print('a library banner')\n"
            "os.write(1, b'a line written by C code\\n')\n"
            "b.emit({'metric': 'x', 'value': 1.5})\n") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("\n") == 1 and json.loads(r.stdout) == {"metric": "x", "value": 1.5}
    assert "a library banner" in r.stderr and "written by C code" in r.stderr


def test_render_configurations_are_the_baseline_ones():
    b = _bench()
    c = b.RENDER_CONFIGS
    assert (c["c3"]["W"], c["c3"]["H"], c["c3"]["depth"], c["c3"]["variant"]) == (1920, 1080, 16, "diffuse")
    assert (c["c4"]["W"], c["c4"]["H"], c["c4"]["depth"], c["c4"]["variant"]) == (1920, 1080, 16, "glossy")
    assert (c["c5"]["W"], c["c5"]["H"], 2 * c["c5"]["nu"] * c["c5"]["nv"]) == (3840, 2160, 10_000_000)
    import json
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "1920x1080, 1024 spp, depth 16" in base["configs"][2] and "256 spp" in base["configs"][3] and "3840x2160, 256 spp" in base["configs"][4]
    # the traffic record bench.py reads for `roofline.traffic` and the instruction roofline
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert t["dram_bytes_per_launch"] == t["dram_bytes_read"] + t["dram_bytes_write"]
    assert 200 < t["warp_inst_per_launch"] / t["rays_per_launch"] < 400
    assert set(t["render_dram_bytes_per_sample"]) >= {"c3", "c4"}


def test_sample_partition_covers_every_sample_once():
    from spica_b200 import partition
    for spp, world in ((1024, 8), (256, 8), (256, 3), (5, 8), (1, 2)):
        seen = []
        for r in range(world):
            first, count, stride = partition.sample_partition(spp, r, world)
            seen += [first + k * stride for k in range(count)]
        assert sorted(seen) == list(range(spp))
