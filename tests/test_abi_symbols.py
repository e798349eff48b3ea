"""CPU checks of the drop-in boundary: the shared library loads and exports every symbol that
include/banet_abi.h declares; argument errors come back as codes + messages (no compute calls here)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT
from banet_b200 import _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "banet_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(banet_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES.keys())


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/banet_abi.h but not exported"


def test_version_and_error_paths_without_gpu():
    lib = _lib.load()
    assert lib.banet_abi_version() == 1
    assert lib.banet_mlp_param_count(128) == 20 * 128 * 128 + 10 * 128 + 1
    # bad arguments are rejected before any CUDA call
    rc = lib.banet_eqc_fwd(None, None, None, 1, 1, 1, 1, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.banet_last_error()
    assert lib.banet_eqc_workspace_bytes(2, 4096, 128, 134) > 0
    lv = _lib.BanetLevel(32, 307200, 128, 128, 480, 640, 384, 1, 1, 1, 1, 1, 1, 0, 0)
    ws = lib.banet_lm_build_workspace_bytes(ctypes.byref(lv), 0)
    assert 0 < ws < (1 << 30)
    lv_bad = _lib.BanetLevel(32, 307200, 128, 128, 480, 640, 100, 1, 1, 1, 1, 1, 1, 0, 0)
    rc = lib.banet_lm_build(ctypes.byref(lv_bad), None, None, None, 0, None, None, None, None, None, 0, None)
    assert rc == -1 and b"conv2_channels" in lib.banet_last_error()


def test_product_never_touches_the_oracle():
    """The product path must not import, call or fall back to anything under oracle/ -- nor may scripts/ or the C headers: only tests/,
    __graft_entry__.smoke() and bench.py's CPU legs use the checker."""
    for top in ("banet_b200", "scripts", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                    assert "ba_oracle" not in txt and "from helpers import" not in txt, f"{f} references the oracle"
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    body = entry[:entry.index("def smoke")]                       # build() may compile the checker; only smoke() may run it
    assert not re.search(r"^\s*(from|import)\s+oracle\b", body, flags=re.M)


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from banet_b200 import ops
    with pytest.raises(_lib.BanetError):
        ops.equation_construction(torch.zeros(1, 4, 2, 6), torch.zeros(1, 4, 3, 2), torch.zeros(1, 4, 3, 1))
