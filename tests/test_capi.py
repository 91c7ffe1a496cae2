"""The drop-in boundary: libspica_b200.so loads and exports exactly what include/spica_b200.h
declares; without a device it refuses to work (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re

import pytest

from spica_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "spica_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libspica_b200.so does not export " + s
    assert sorted(capi.SYMBOLS) == sorted(s for s in syms), "capi.SYMBOLS out of sync with the header"
    assert L.spb_version() == 3


def test_record_sizes_match_header():
    assert capi.RAY_F32.itemsize == 32 and capi.RAY_F64.itemsize == 64
    assert capi.HIT.itemsize == 16 and capi.HIT_F64.itemsize == 32 and capi.IMPORT_NODE.itemsize == 64


def test_no_device_means_no_context():
    from tests.conftest import has_cuda_device
    if has_cuda_device():
        pytest.skip("a device is present")
    with pytest.raises(capi.SpbError) as ei:
        capi.Context(0)
    assert "no CUDA device" in str(ei.value) or "CPU" in str(ei.value)


def test_product_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under spica_b200/ may import, link or exec it
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "spica_b200")):
        if "lib" in dp.split(os.sep):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cpp", ".h", ".cc", "Makefile")):
                if re.search(r"(from|import)\s+oracle|libspica_oracle|spica_oracle\.h|oracle/_|raycast_ref|tests\.emul|libspb_emul",
                             open(os.path.join(dp, fn), errors="ignore").read()):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
