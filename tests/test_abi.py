"""The C-ABI boundary: libtqb200.so loads, exports every symbol include/tqb200.h declares, and the ctypes
prototypes cover exactly that set.  No compute calls (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tqb200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"TQ_API\s+[\w\s\*]+?\b(tq_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from torchquad_b200 import build

    if not os.path.exists(build.LIB):
        build.build()
    return build.LIB


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 27
    for must in ["tq_philox_uniform", "tq_mc_sample", "tq_sum_columns", "tq_vegas_map_forward", "tq_vegas_map_accumulate",
                 "tq_vegas_map_update", "tq_vegas_strat_nh", "tq_vegas_strat_sample", "tq_vegas_strat_accumulate",
                 "tq_vegas_strat_update", "tq_nc_grid_points", "tq_nc_contract", "tq_fused_mc", "tq_fused_vegas", "tq_fused_nc"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/tqb200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (tq_[a-z0-9_]+)", out)))
    assert exported == declared_symbols(), "exported tq_* symbols and header declarations differ"
    assert not re.search(r" T (?!tq_)\w*torch", out), "no torch types/symbols may leak into the C ABI"


def test_ctypes_prototypes_match_header(lib_path):
    from torchquad_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _lib.PROTOTYPES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, ctypes {len(argtypes)}"
    lib = _lib.load()
    assert lib.tq_version() == 100
    assert lib.tq_workspace_bytes() >= 1 << 20
    assert ctypes.sizeof(_lib.tq_integrand) == 16 + 8 * (32 * 4 + 8) + 8


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torchquad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_cpu_tensors_fail_loudly():
    import torch

    import torchquad_b200 as tq

    with pytest.raises(RuntimeError, match="CUDA only"):
        tq.MonteCarlo().integrate(lambda x: x.sum(1), 2, 100, torch.tensor([[0.0, 1.0]] * 2), seed=0)
    with pytest.raises(RuntimeError, match="CUDA only"):
        tq.VEGASMap(10, 2, "torch", torch.float64, device="cpu").get_X(torch.rand(4, 2, dtype=torch.float64))


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/tqb200.h must compile as C (it is the FFI contract) and the ctypes mirrors must have the C layout."""
    from torchquad_b200 import _lib

    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "tqb200.h"\n'
        "int main(void) {\n"
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tq_integrand), sizeof(tq_vegas_state), sizeof(tq_vegas_result),\n'
        "         offsetof(tq_vegas_state, edges_layout), offsetof(tq_vegas_result, results), offsetof(tq_vegas_result, status));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.tq_integrand), ctypes.sizeof(_lib.tq_vegas_state), ctypes.sizeof(_lib.tq_vegas_result),
            _lib.tq_vegas_state.edges_layout.offset, _lib.tq_vegas_result.results.offset, _lib.tq_vegas_result.status.offset]
    assert got == want
