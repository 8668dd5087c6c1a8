"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/evostencils_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from evostencils_b200 import backend, oplist as ol, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "evostencils_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evo_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(backend.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_all_symbols():
    lib = backend.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.evo_abi_version() == ol.ABI_VERSION


def test_struct_layout_matches_header():
    # sizes computed by hand from include/evostencils_b200.h (natural alignment, 8-byte doubles)
    assert C.sizeof(ol.CEvoOp) == 8 * 4 + 8 * 4 + 8 * 3 * 4 + 2 * 8
    assert C.sizeof(ol.CEvoLevelOperator) == 8 + 2 * 2 * 27 * 2 * 8
    assert C.sizeof(ol.CEvoProblemDesc) == 8 * 4 + 3 * 8 + 2 * 27 * 8
    assert C.sizeof(ol.CEvoSolveParams) == 8 + 4 * 4
    assert C.sizeof(ol.CEvoSolveResult) == 2 * 4 + 4 * 8 + 8


def test_no_cpu_fallback():
    lib = backend.load_library()
    if lib.evo_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(backend.BackendError):
        backend.DeviceProblem(problems.Poisson2D(3, 4))
    # the raw entry point reports EVO_ERR_NO_DEVICE rather than computing on the host
    desc = backend.make_desc(problems.Poisson2D(3, 4))
    h = C.c_void_p()
    assert lib.evo_problem_create(C.byref(desc), C.byref(h)) == -3
    assert b"no CPU fallback" in lib.evo_last_error() or b"cudaGetDeviceCount" in lib.evo_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "evostencils_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f
