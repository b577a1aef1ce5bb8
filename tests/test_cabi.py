"""CPU: the C-ABI library loads and exports every symbol include/zkb200.h declares; without a GPU it refuses to work
(no CPU fallback) instead of silently computing elsewhere.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from zk_stark_project_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = L.load()
    declared = header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/zkb200.h but not exported by libzkb200.so"
    assert sorted(L.EXPORTS) == declared, "zk_stark_project_b200.lib.EXPORTS is out of sync with include/zkb200.h"


def test_struct_layouts_match_header():
    # zkb_transcript: 3*32 + 3*16 + 8 + 16*32 + 16*16 + 8 + 256*4 + 8
    assert C.sizeof(L.Transcript) == 96 + 48 + 8 + 512 + 256 + 8 + 1024 + 8
    assert C.sizeof(L.StageTimes) == 13 * 4
    assert C.sizeof(L.AirDesc) == 8 + 8 + 8 * 4 + 8 * 2 + 8 * 4 + 8 * 2


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the refusal path is exercised on the CPU-only box")
    lib = L.load()
    h = C.c_void_p()
    rc = lib.zkb_ctx_create(C.c_int32(0), None, C.byref(h))
    assert rc == -2 and not h.value  # ZKB_ERR_CUDA
    assert b"CUDA" in lib.zkb_last_error(None) or b"device" in lib.zkb_last_error(None)
    with pytest.raises(L.ZkbError):
        L.Context(0)


def test_product_does_not_import_oracle():
    """The shipped package must not reach into oracle/ (the oracle is test infrastructure only)."""
    pkg = os.path.join(ROOT, "zk_stark_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and '"../oracle' not in text and "oracle/" not in text.replace("oracle/)", ""), f
