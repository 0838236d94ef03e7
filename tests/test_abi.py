"""The C-ABI library builds, loads and exports every symbol include/asm_b200.h declares (CPU: no compute)."""
import ctypes as C
import os
import re

import pytest

from activesetmethods_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "asm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in include/asm_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == syms, "capi.SIGNATURES and the header disagree"


def test_default_params(built_lib):
    p = capi.default_params()
    assert p.eps_rel == 1e-8 and p.check_every == 64 and p.max_iter == 2000000
    assert C.sizeof(capi.LpParams) == 8 * 2 + 8 + 4 * 4 + 8 * 6 + 8 * 4 + 4 * 2 + 8 * 2
    assert p.ipm_max_iter == 200 and p.ipm_refine == 0 and p.ipm_reg == 1e-8 and p.ipm_prox == 1e-7
    assert C.sizeof(capi.LpInfo) == 4 + 4 + 8 + 8 * 5


def test_no_cpu_fallback(built_lib):
    """Without a device, creating a handle fails loudly (ASM_E_CUDA) instead of computing on the host."""
    if built_lib.asm_device_count() > 0:
        pytest.skip("a GPU is present")
    from activesetmethods_b200.sublp import SubLp
    import numpy as np
    with pytest.raises(capi.AsmError) as e:
        SubLp(2, 1, np.array([[1, 1], [1, 2]]), [-1, -1], [1, 1], [0.0], [0.0])
    assert e.value.code == capi.E_CUDA


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline may touch oracle/."""
    pkg = os.path.join(ROOT, "activesetmethods_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "slp_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
