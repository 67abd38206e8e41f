"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/mcaq_b200.h declares (no compute calls; there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcaq_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MCAQ_API\s+[\w\s\*]+?\b(mcaq_\w+|launch_spatial_quantization)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    from mcaq_yolo_b200 import _lib
    return _lib.LIB_PATH


def test_header_declares_the_path():
    syms = declared_symbols()
    for name in ("launch_spatial_quantization", "mcaq_reduce_planes", "mcaq_morph_phi", "mcaq_tile_quantize",
                 "mcaq_tile_quantize_train_bwd", "mcaq_bit_mapper", "mcaq_soft_mask", "mcaq_complexity"):
        assert name in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/mcaq_b200.h but not exported"


def test_ctypes_table_matches_header(lib_path):
    from mcaq_yolo_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    handle = _lib.load()
    assert handle.mcaq_abi_version() == 1
    assert handle.mcaq_tile_size(80, 8) == 8 and handle.mcaq_tile_size(20, 8) == 4
    assert handle.mcaq_tile_size(160, 4) == 32 and handle.mcaq_tile_size(50, 8) == 4
    assert b"invalid argument" in handle.mcaq_error_string(-1)


def test_library_is_sm100a_only(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_in_product():
    """The product never imports the oracle and refuses CPU tensors."""
    import torch
    pkg = os.path.join(ROOT, "mcaq_yolo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "mcaq_oracle" not in src and "import oracle" not in src, fn
    from mcaq_yolo_b200 import ops
    with pytest.raises(RuntimeError):
        ops.reduce_planes(torch.zeros(1, 4, 8, 8))
