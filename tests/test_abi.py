"""The C-ABI library loads and exports every symbol that include/immunostruct_b200.h declares
(no compute calls: this runs on the CPU-only box)."""
import ctypes
import os
import re

from conftest import ROOT
from immunostruct_b200 import _C, build


def header_symbols():
    text = open(os.path.join(ROOT, "include", "immunostruct_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|int64_t)\s+(is_\w+)\s*\(", text, flags=re.M)))


def test_library_exports_every_declared_symbol():
    path = build.build()                      # no-op when the in-tree .so is current
    lib = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 19
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_C.exported_symbols()) == declared


def test_sources_are_built_for_sm_100a():
    assert "arch=compute_100a,code=sm_100a" in " ".join(build.FLAGS)
    assert "-lineinfo" in build.FLAGS
