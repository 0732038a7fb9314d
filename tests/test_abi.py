"""The C-ABI library loads on a machine without a GPU and exports every symbol include/emba_b200.h declares; the
ctypes signature table covers exactly the declared functions; without a CUDA device the product path fails loudly
(no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "emba_b200.h")).read()
    return sorted(set(re.findall(r"EMBA_API\s+(?:const\s+char\*|int)\s+(emba_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from emba_b200 import capi

    lib = capi.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/emba_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names
    assert lib.emba_version().startswith(b"emba_b200")


def test_every_entry_point_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "emba_b200.h")).read()
    assert src.count("model.cpp:") >= 10 and "solver.cpp:11-368" in src and "model.h:" in src


def test_no_cpu_fallback_without_gpu():
    from emba_b200 import capi

    lib = capi.load()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("CUDA device present")
    lut = np.zeros(3 * 4 * 4)
    cfg = capi.Config(4, 4, 8, 4, 0.2, capi.ptr(lut), 0)
    h = C.c_void_p()
    assert lib.emba_create(C.byref(cfg), C.byref(h)) == -2  # EMBA_E_CUDA, nothing computed on the CPU
    from emba_b200.legm import Engine
    with pytest.raises(capi.EmbaError):
        Engine(4, 4, lut, 0.2, 8, 4)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "emba_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("cpu oracle", ""), f"{f} references oracle/"
