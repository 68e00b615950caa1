"""The C-ABI library loads and exports exactly what include/pfst_sm100.h declares.
No compute entry point is called here (this runs on a CPU-only host)."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "pfst_sm100.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pfst_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    from pfst_b200 import build, _lib
    build.build()
    lib = _lib.load()
    assert b"sm_100a" in lib.pfst_version()
    assert lib.pfst_error_string(0) == b"ok"
    assert b"invalid" in lib.pfst_error_string(-1)


def test_every_declared_symbol_is_exported_and_bound():
    from pfst_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (pfst_[a-z0-9_]+)", out)))
    assert exported == names, "library exports symbols the header does not declare (or vice versa)"


def test_only_sm100a_code_in_library():
    from pfst_b200 import build, _lib
    build.build()
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_only_entry_points():
    from pfst_b200 import ops
    a, b = ops.ema_coeffs(5000, 0.999)
    import numpy as np
    assert a == np.float32(0.999) and b == np.float32(1.0 - 0.999)
    a, b = ops.ema_coeffs(1, 0.999)
    assert a == 0.5 and b == 0.5


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from pfst_b200 import ops
    with pytest.raises(ops.PfstError):
        ops.pseudo_label(torch.zeros(1, 2, 4, 4), 0.5)


def test_product_never_imports_oracle():
    for p in (ROOT / "pfst_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
        assert "tests.golden" not in src, p
