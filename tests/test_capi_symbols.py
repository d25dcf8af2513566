"""The C-ABI library loads and exports every symbol include/pino_locoman_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pino_locoman_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_header_symbols():
    from pino_locoman_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names


def test_struct_sizes_match_header():
    from pino_locoman_b200 import _lib
    from pino_locoman_b200.utils.robot import RobotDesc
    sizes = (ctypes.c_int32 * 3)()
    _lib.load().plm_abi_struct_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(RobotDesc), ctypes.sizeof(_lib.OcpDesc), ctypes.sizeof(_lib.Dims)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pino_locoman_b200.handle import Handle
    from pino_locoman_b200.utils.robot import B2
    r = B2()
    r.set_gait_sequence("trot", 0.8)
    with pytest.raises(RuntimeError):
        Handle(r, "whole_body_rnea", 5, max_batch=1)
