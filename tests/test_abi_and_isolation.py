"""CPU checks of the drop-in boundary: libepb200.so loads and exports every symbol include/epb200.h declares (no
compute call is made: there is no GPU here), the ctypes table mirrors the header one to one, the product package
never imports the oracle, and the product path fails loudly without a CUDA device."""

import ast
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "epb200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(epb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from echopype_b200 import _lib
    from echopype_b200.build import build

    build()  # nvcc cross-compiles for sm_100a without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/epb200.h but not exported by libepb200.so"
    assert lib.epb_version() == 100
    lib.epb_pipeline_workspace_bytes.restype = ctypes.c_longlong
    lib.epb_pipeline_workspace_bytes.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int]
    assert lib.epb_pipeline_workspace_bytes(4, 100, 5) == 256 + 4 * 20 * 144


def test_ctypes_table_mirrors_header():
    from echopype_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()  # resolves every symbol and sets argtypes


def test_bad_arguments_return_error_codes_without_a_gpu():
    """Argument validation happens before any CUDA call: NULL pointers give EPB_E_BADARG and a message."""
    from echopype_b200 import _lib

    lib = _lib.load()
    rc = lib.epb_sv_power(None, None, None, None, None, 1, 1, 4, None)
    assert rc == -1
    assert b"NULL" in lib.epb_last_error()
    rc = lib.epb_bin_finalize(None, None, None, 0, 1, ctypes.c_float(0.0), 1, None)
    assert rc == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "echopype_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dp, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                if any(m == "oracle" or m.startswith("oracle.") or m == "oracle_glue" for m in mods):
                    bad.append(os.path.join(dp, f))
    assert not bad, f"product modules import the test oracle: {bad}"


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import echopype_b200 as ep
    from echopype_b200 import synth
    from echopype_b200._lib import EpbError

    ed = synth.make_ek60(2, 8, 64, seed=1)
    with pytest.raises(EpbError, match="no CPU fallback"):
        ep.calibrate.compute_Sv(ed)
    with pytest.raises(EpbError, match="no CPU fallback"):
        ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=5, range_sample_num=30)
