"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/srfrd_b200.h declares with the arity _lib.py binds (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from srfrd_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "srfrd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"SRFRD_API\s+(?:const\s+char\*|int)\s+(srfrd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def test_library_exports_every_declared_symbol(lib):
    protos = header_prototypes()
    assert len(protos) >= 28
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/srfrd_b200.h but not exported"


def test_python_bindings_match_header_arity(lib):
    protos = header_prototypes()
    protos.pop("srfrd_last_error")
    assert set(protos) == set(lib.SIGNATURES), set(protos) ^ set(lib.SIGNATURES)
    for name, n in protos.items():
        assert len(lib.SIGNATURES[name]) == n, f"{name}: header has {n} args, _lib.py binds {len(lib.SIGNATURES[name])}"
    lib.load()


def test_struct_layouts_match_header(lib):
    # sizes a C compiler gives the two structs in the header (LP64)
    assert ctypes.sizeof(lib.GemmEpilogue) == 6 * 8 + 4 * 4 + 4 + 4 + 8 + 8
    assert ctypes.sizeof(lib.CastDesc) == 56


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from srfrd_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.weight_sums(torch.zeros(4, dtype=torch.int64), None, None, torch.zeros(2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "srfrd_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), f"{fn} references the oracle"
