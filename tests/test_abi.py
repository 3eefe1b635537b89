"""CPU: the C-ABI library loads and exports every symbol include/p2vit_b200.h declares (no compute calls)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "p2vit_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2v_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = header_symbols()
    assert "p2v_gemm_i8" in syms and "p2v_attention_i8" in syms and len(syms) >= 16


def test_library_exports_every_header_symbol():
    from p2vit_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = _lib.load()
    for s in header_symbols():
        assert hasattr(lib, s), "library does not export %s" % s
    assert sorted(_lib.SYMBOLS) == header_symbols(), "ctypes table and header diverge"
    assert lib.p2v_abi_version() == 3
    assert lib.p2v_launch_count() >= 0


def test_library_is_blackwell_native():
    """SASS evidence: tcgen05.mma (UTC*MMA), TMA (UTMALDG), TMEM loads (LDTM) are in the shipped binary."""
    from p2vit_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("library not built")
    try:
        sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for mnemonic in ("UTCIMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    import torch
    from p2vit_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.quantize(torch.zeros(4, 8), torch.tensor([0.5]))
