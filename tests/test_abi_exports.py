"""CPU: the C-ABI library builds, loads and exports every symbol include/cavit.h declares
(no compute calls — there is no GPU here)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cavit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cavit_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from cavit import _abi
    L = _abi.lib()
    declared = _declared()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (cavit_[a-z0-9_]+)", out))
    for name in declared:
        assert name in exported, f"{name} declared in include/cavit.h but not exported"
        assert hasattr(L, name)
    assert set(_abi.EXPORTS) == set(declared), "ctypes binding table and header disagree"
    assert L.cavit_abi_version() == _abi.ABI_VERSION


def test_library_is_sm100a_only_and_uses_tcgen05_tma():
    from cavit import _abi
    sass = subprocess.run(["cuobjdump", "-sass", _abi.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        return  # cuobjdump unavailable
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_no_gpu_means_loud_failure():
    import pytest
    import torch
    from cavit import _abi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_abi.CavitError):
        _abi.require_device(0)
