"""The C-ABI library loads on a CPU-only box, exports every symbol the header declares,
and refuses (loudly) to run without a B200 — there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu
from protocol_b200 import abi
from protocol_b200._lib import LIB_PATH, PrimeMatchError, load


def header_symbols():
    text = open(os.path.join(ROOT, "include", "prime_match.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(LIB_PATH)
    declared = header_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in prime_match.h but not exported"
    assert sorted(abi.EXPORTS) == declared


def test_abi_version_and_struct_sizes():
    lib = load()
    assert lib.pm_abi_version() == abi.PM_ABI_VERSION
    assert C.sizeof(abi.PmAsk) == 32 and C.sizeof(abi.PmGpuOpt) == 32
    assert C.sizeof(abi.PmCfg) == 40
    assert C.sizeof(abi.PmStats) == 88


def test_product_does_not_reference_the_oracle():
    """No file under protocol_b200/ may import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "protocol_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pm_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


@pytest.mark.skipif(has_gpu(), reason="CPU-only behaviour")
def test_engine_fails_loudly_without_gpu():
    from protocol_b200.engine import Engine

    with pytest.raises(PrimeMatchError) as ei:
        Engine()
    assert ei.value.status == abi.PM_E_NO_DEVICE
    assert "no CPU path" in str(ei.value) or "sm_100" in str(ei.value)


def test_integration_doc_binds_every_export():
    """INTEGRATION.md shows the reference-side (Rust) binding of the boundary: every exported entry point is named."""
    import os
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    missing = [name for name in abi.EXPORTS if name not in doc]
    assert not missing, missing
