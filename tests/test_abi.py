"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/ptk.h
declares; ctypes structs match the header; construction fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ptk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptk_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ptudes_lab_b200 import _ffi
    lib = _ffi.load()
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_ffi.SYMBOLS) == names
    assert lib.ptk_version() == 100


def test_default_config_matches_kiss_icp_defaults():
    from ptudes_lab_b200 import _ffi
    lib = _ffi.load()
    c = _ffi.PtkConfig()
    lib.ptk_default_config(C.byref(c))
    assert (c.max_range, c.min_range, c.max_points_per_voxel, c.deskew) == (100.0, 5.0, 20, 1)
    assert (c.initial_threshold, c.min_motion_th, c.max_iterations, c.convergence_eps) == (2.0, 0.1, 500, 1e-4)
    assert c.batch == 1


def test_struct_sizes_match_header():
    from ptudes_lab_b200 import _ffi
    assert C.sizeof(_ffi.PtkStats) == 8 * 4 + 4 * 8 + 2 * 4
    assert C.sizeof(_ffi.PtkConfig) == 80


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ptudes_lab_b200 import odometry
    with pytest.raises(RuntimeError) as e:
        odometry.Odometry()
    assert "PTK_E_CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ptudes_lab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
