"""CPU checks of the drop-in boundary: libhalab200.so loads, exports exactly what include/halab200.h declares,
and fails loudly (status + message, no fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "halab200.h")).read() + open(os.path.join(ROOT, "include", "halab200_dist.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hala_b200 import capi
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(capi.lib, n), f"{n} declared in include/halab200.h but not exported"
    from hala_b200 import dist as hbdist
    assert sorted(list(capi.SIGNATURES) + list(hbdist.DIST_SIGNATURES)) == names, "ctypes tables and include/*.h disagree"
    assert b"sm_100a" in capi.lib.hb_version()


def test_library_is_sm100a_native():
    """The shipped cubin must be sm_100a SASS (no PTX-JIT fallback to another arch)."""
    import subprocess
    from hala_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_no_silent_cpu_fallback():
    """Without a GPU every compute entry point reports an error; nothing computes on the host."""
    from hala_b200 import capi
    n = C.c_int(-1)
    assert capi.lib.hb_device_count(C.byref(n)) == 0
    if n.value > 0:
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    assert capi.lib.hb_ctx_create(0, C.byref(ctx)) != 0
    assert capi.lib.hb_last_error() != b""
    import hala_b200 as hb
    with pytest.raises(hb.HalaB200Error):
        hb.gpu_engine(0)


def test_argument_errors_are_reported():
    from hala_b200 import capi
    assert capi.lib.hb_ctx_create(0, None) == 2          # HB_ERR_ARG
    assert b"invalid argument" in capi.lib.hb_last_error()
    assert capi.lib.hb_ctx_set_pointer_mode(None, 0) == 2


def test_enumerations_agree_with_the_header():
    """The ctypes mirror restates the header's enumerations by value: keep them equal."""
    from hala_b200 import capi
    text = open(os.path.join(ROOT, "include", "halab200.h")).read()
    enums = {}
    for body in re.findall(r"enum\s*\{([^}]*)\}", text):
        for name, val in re.findall(r"(HB_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            enums[name] = int(val)
    assert len(enums) >= 15
    for name, val in enums.items():
        if hasattr(capi, name):
            assert getattr(capi, name) == val, name
    for name in ("HB_F32", "HB_F64", "HB_C32", "HB_C64", "HB_H2D", "HB_D2H", "HB_D2D", "HB_TRANS_SCATTER", "HB_TRANS_CHECKED", "HB_TRANS_FROZEN"):
        assert name in enums and hasattr(capi, name), name


def test_transpose_entry_points_reject_null_handles():
    """hb_csr_set_transpose_mode / hb_csr_values_changed / hb_csr_transpose_info validate their handle before touching the device."""
    from hala_b200 import capi
    assert capi.lib.hb_csr_set_transpose_mode(None, capi.HB_TRANS_FROZEN) == 2
    assert capi.lib.hb_csr_values_changed(None) == 2
    assert capi.lib.hb_csr_transpose_info(None, None, None, None) == 2
    assert b"invalid argument" in capi.lib.hb_last_error()


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under hala_b200/ may import, link or mention it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hala_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"import\s+oracle|from\s+oracle|oracle[/.]binding|liboracle|libhala_ref|oracle/_ref", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
