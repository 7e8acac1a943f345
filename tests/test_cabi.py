"""The C-ABI library loads on a CPU-only box and exports exactly what include/atmvfi.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "atmvfi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(atmvfi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    from atmvfi import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/atmvfi.h but not exported"
    assert sorted(_lib.ALL_SYMBOLS) == declared, "python binding and header disagree"
    lib.atmvfi_abi_version.restype = ctypes.c_int
    assert lib.atmvfi_abi_version() == 3


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from atmvfi import _lib
    from network_lite import Network
    net = Network()
    with pytest.raises(_lib.AtmvfiError):
        net(torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "atm-vfi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "atmvfi_oracle" not in text and "import oracle" not in text and "emul_ops" not in text, f


def test_binding_argument_counts_match_header():
    """Every ctypes prototype in atmvfi/_lib.py takes as many arguments as the declaration in include/atmvfi.h (a count that
    drifts makes ctypes push garbage into the trailing parameters without any error)."""
    from atmvfi import _lib
    src = open(os.path.join(ROOT, "include", "atmvfi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    decl = {}
    for m in re.finditer(r"\b(atmvfi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    checked = 0
    for name, argt in _lib.PROTOTYPES.items():
        assert name in decl, name
        assert len(argt) == decl[name], f"{name}: binding has {len(argt)} arguments, header declares {decl[name]}"
        checked += 1
    for name, (argt, _) in _lib._SPECIAL.items():
        assert name in decl, name
        assert len(argt) == decl[name], f"{name}: binding has {len(argt)} arguments, header declares {decl[name]}"
        checked += 1
    assert checked >= 40
