"""CPU-side checks: the C-ABI library loads and exports every symbol include/disenlink_b200.h
declares (no compute calls without a GPU); the host-side mirror of model.py keeps the reference's
API and state_dict keys; the product never routes through the oracle or a CPU fallback."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "disenlink_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from disenlink_b200 import _lib, build
    path = build.build()
    handle = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    handle.dl_abi_version.restype = ctypes.c_int
    assert handle.dl_abi_version() == 2
    handle.dl_error_string.restype = ctypes.c_char_p
    assert b"symmetric" in handle.dl_error_string(-4)
    # size queries are pure host functions
    handle.dl_csr_build_workspace_bytes.restype = ctypes.c_size_t
    handle.dl_csr_build_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    assert handle.dl_csr_build_workspace_bytes(1000, 100) >= 2 * 2000 * 8


def test_struct_layout_matches_header():
    from disenlink_b200._lib import DlGraph
    txt = open(os.path.join(ROOT, "include", "disenlink_b200.h")).read()
    body = txt[txt.index("typedef struct dl_graph {"):txt.index("} dl_graph;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(\w+);", body)
    assert fields == [f[0] for f in DlGraph._fields_]
    assert ctypes.sizeof(DlGraph) == 8 * len(fields)


def test_product_never_touches_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libdisen_oracle|dlo_", re.M)
    files = glob.glob(os.path.join(ROOT, "disenlink_b200", "**", "*.*"), recursive=True)
    files = [f for f in files if f.endswith((".py", ".cu", ".cuh", ".h"))]
    assert files
    for f in files:
        assert not pat.search(open(f).read()), f"{f} references the oracle"


def test_missing_extension_fails_loudly(monkeypatch):
    from disenlink_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdisenlink_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    from disenlink_b200.model import Disentangle
    from disenlink_b200.graph import Graph
    m = Disentangle(6, 8, 4, nfactor=2, beta=0.5)
    x = torch.randn(5, 6)
    adj = torch.eye(5)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, adj)
    with pytest.raises(RuntimeError, match="CUDA"):
        Graph.from_edges(torch.tensor([0, 1]), torch.tensor([1, 2]), 3)


@pytest.mark.parametrize("fixture", ["module_small", "module_nhid1"])
def test_state_dict_is_interchangeable_with_the_reference(fixture):
    from disenlink_b200.model import Disentangle
    g = load_golden(fixture)
    m = Disentangle(int(g["F"]), int(g["nhid"]), int(g["d"]), nfactor=int(g["K"]), beta=float(g["beta"]),
                    t=1)
    ref_keys = sorted(k[3:] for k in g if k.startswith("sd."))
    assert sorted(m.state_dict().keys()) == ref_keys
    m.load_state_dict({k: torch.from_numpy(g["sd." + k]) for k in ref_keys}, strict=True)
    # the projection (library GEMM) reproduces the reference's Z, checked through H = beta*Z on an
    # edgeless graph would need the GPU; here only shapes and parameter registration
    assert len(list(m.parameters())) == len(ref_keys)
    x = torch.from_numpy(g["x"])
    Z = m.project(x)
    assert tuple(Z.shape) == (int(g["N"]), int(g["K"]), int(g["d"]))


def test_reference_class_names_are_importable():
    import disenlink_b200.model as M
    for name in ("Factor", "Factor2", "Dec", "Dec2", "Disentangle_layer", "Disentangle_out_layer",
                 "Disentangle"):
        assert hasattr(M, name)
    lay = M.Disentangle_layer(3, 0.7, t=2)
    assert (lay.nfactor, lay.beta, lay.temperature) == (3, 0.7, 2)


def test_one_minus_beta_matches_python_double_rounding():
    from disenlink_b200.ops import one_minus
    from oracle import oracle
    for b in (0.5, 0.6, 0.7, 0.8, 0.9, 0.1):
        assert np.float32(one_minus(b)) == oracle.one_minus(b)


def test_ctypes_signatures_match_the_header_parameter_lists():
    """Every declaration of include/disenlink_b200.h against the ctypes signature the host side binds it
    with: same number of parameters, same class (pointer / int / int64 / float / size) in every position --
    a drifted binding would pass garbage across the ABI without any error."""
    from disenlink_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "disenlink_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    decls = re.findall(r"\b(dl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
    assert len(decls) >= 40
    scalar = {"int64_t": "i64", "int": "i32", "int32_t": "i32", "float": "f32", "double": "f64", "size_t": "u64",
              "uint64_t": "u64", "uint32_t": "u32", "dl_stream_t": "ptr", "cudaStream_t": "ptr"}

    def header_kind(p):
        p = " ".join(p.split())
        if p in ("", "void"):
            return None
        if "*" in p or "[" in p:
            return "ptr"
        base = re.sub(r"\b(const|struct)\b", "", p).split()[0]
        return scalar[base]

    def ctypes_kind(a):
        if a in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(a, (ctypes._Pointer, ctypes.Array)):
            return "ptr"
        return {ctypes.c_int64: "i64", ctypes.c_int: "i32", ctypes.c_float: "f32", ctypes.c_double: "f64",
                ctypes.c_size_t: "u64", ctypes.c_uint64: "u64", ctypes.c_uint32: "u32"}[a]

    for name, params in decls:
        want = [k for k in (header_kind(p) for p in params.split(",")) if k is not None]
        got = [ctypes_kind(a) for a in _lib.SIGNATURES[name][1]]
        assert want == got, f"{name}: header {want} vs ctypes {got}"
