"""C-ABI checks that need no GPU: the built library loads, exports every entry point ``include/cinema_b200.h`` declares,
the ctypes binding (cinema_b200/_C.py) names the same set with the same number of arguments, the product path refuses
CPU tensors / a missing library loudly (no fallback), and nothing in the package imports the oracle."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

from cinema_b200 import _C

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "cinema_b200.h").read_text()


def _declarations() -> dict[str, int]:
    """entry point -> number of parameters, parsed from the header (comments stripped)."""
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*)\s*(cb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_header_declares_the_documented_entry_points():
    decl = _declarations()
    assert len(decl) >= 28
    for name in ("cb_gemm_bf16", "cb_attention_fwd", "cb_attention_bwd", "cb_layernorm_fwd", "cb_layernorm_bwd", "cb_patchify",
                 "cb_gather_patches", "cb_masked_mse_fwd", "cb_rope_apply", "cb_adamw_flat", "cb_last_error"):
        assert name in decl, name


def test_library_loads_and_exports_every_declared_symbol():
    assert _C.LIB_PATH.exists(), "build the library first: python -m cinema_b200.build"
    handle = ctypes.CDLL(str(_C.LIB_PATH))  # loading must not need a GPU
    missing = [n for n in _declarations() if not hasattr(handle, n)]
    assert not missing, missing
    handle.cb_last_error.restype = ctypes.c_char_p
    assert isinstance(handle.cb_last_error(), bytes)


def test_ctypes_binding_matches_the_header():
    decl = _declarations()
    assert set(_C._SIGNATURES) == set(decl)
    for name, (_, argtypes) in _C._SIGNATURES.items():
        assert len(argtypes) == decl[name], (name, len(argtypes), decl[name])
    lib = _C.lib()  # sets restype / argtypes on every symbol
    assert lib.cb_gemm_bf16.argtypes is not None


def test_every_entry_point_cites_the_reference_it_replaces():
    """Each declaration is preceded by a comment naming the reference call site (cinema/...:line) or states that it is a
    support routine of the library itself."""
    blocks = re.split(r"\n(?=/\*)", HEADER)
    cited = sum(1 for b in blocks if re.search(r"cb_[a-z0-9_]+\s*\(", b) and re.search(r"cinema/[a-z_/]+\.py:\d+", b))
    assert cited >= 15, cited


def test_product_path_refuses_cpu_tensors_and_missing_library(monkeypatch, tmp_path):
    x = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _C.gemm(x, x, torch.zeros(8, 8, dtype=torch.bfloat16))
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", tmp_path / "libcinema_b200.so")
    with pytest.raises(_C.KernelLibraryMissing):
        _C.lib()


def test_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may use it."""
    offenders = []
    for path in (ROOT / "cinema_b200").rglob("*.py"):
        if re.search(r"^\s*(from|import)\s+oracle\b", path.read_text(), flags=re.M):
            offenders.append(str(path))
    assert not offenders, offenders
    for path in (ROOT / "cinema_b200").rglob("*.py"):
        assert "/root/reference" not in path.read_text(), path
