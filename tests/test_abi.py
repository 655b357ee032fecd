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


def test_struct_layouts_match_header(lib, tmp_path):
    """sizeof / field offsets the C compiler gives the header's structs == the ctypes mirrors in _lib.py."""
    import shutil
    import subprocess
    assert ctypes.sizeof(lib.GemmEpilogue) == 128 and ctypes.sizeof(lib.CastDesc) == 56      # LP64
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "srfrd_b200.h"\n'
        "int main(void) {\n"
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(srfrd_gemm_epilogue_t), offsetof(srfrd_gemm_epilogue_t, ldc),\n'
        "         offsetof(srfrd_gemm_epilogue_t, drop_seed), offsetof(srfrd_gemm_epilogue_t, ln_out_bf16),\n"
        "         offsetof(srfrd_gemm_epilogue_t, ln_eps), sizeof(srfrd_cast_desc_t));\n  return 0;\n}\n")
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    G = lib.GemmEpilogue
    assert got == [ctypes.sizeof(G), G.ldc.offset, G.drop_seed.offset, G.ln_out_bf16.offset, G.ln_eps.offset,
                   ctypes.sizeof(lib.CastDesc)]


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
