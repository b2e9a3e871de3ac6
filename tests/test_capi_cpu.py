"""The C-ABI library loads on a CPU-only box and exports every symbol include/pbmc.h declares.
No compute call is made here (argument validation returns before any CUDA call)."""
import ctypes as C
import os
import re

import pytest

from pbml_mantle_convection_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G

    G.build()
    return L.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pbmc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pbmc_[a-z_0-9A-Z]+)\s*\(", txt)))


def test_header_and_binding_agree(lib):
    decl = declared_symbols()
    assert set(decl) == set(L.SIGNATURES), (set(decl) ^ set(L.SIGNATURES))
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in pbmc.h but not exported by libpbmc.so"


def test_struct_sizes(lib):
    for name, st in L._STRUCTS.items():
        assert lib.pbmc_sizeof(name.encode()) == C.sizeof(st)
    assert lib.pbmc_sizeof(b"nope") == 0
    assert C.sizeof(L.Member) == 32


def test_version_and_errors(lib):
    assert lib.pbmc_version() == 3
    assert lib.pbmc_error_string(0) == b"ok"
    assert b"workspace" in lib.pbmc_error_string(-5)


def test_argument_validation_without_gpu(lib):
    assert lib.pbmc_pack_nchw(None, None, 1, 1, 4, 4, None) == -3
    assert lib.pbmc_conv_fwd(None, None) == -3
    assert lib.pbmc_advect_diffuse(None, None, None, None, None, None, None, 1, 0.1, 0.99, 0.0, None, None, None, 1, 8, 8,
                                   None) == -3
    assert lib.pbmc_head(None, None, None, 1.0, 0, 1, None, None, None, None, 1, 8, 8, None) == -3
    n = L.Net()
    assert lib.pbmc_workspace_bytes(C.byref(n), 1, 8, 8) == 0  # invalid (all-zero) network
    n.levels, n.repeats, n.c_i, n.c_h, n.c_o, n.ksize = 6, 4, 7, 16, 2, 3
    b1 = lib.pbmc_workspace_bytes(C.byref(n), 1, 512, 512)
    b2 = lib.pbmc_workspace_bytes(C.byref(n), 2, 512, 512)
    assert b1 > 512 * 512 * 16 * 4 * 10 and b2 > 1.9 * b1
    assert lib.pbmc_workspace_bytes(C.byref(n), 1, 16, 16) == 0  # 6 levels do not fit a 16x16 grid
    # pbmc_rollout refuses a dt history that the requested steps would overrun (step i writes row i - 1) -- checked
    # before anything is dereferenced, so dummy non-null addresses are enough here
    d = C.c_void_p(0x1000)
    args = lambda first, k, rows: (d, C.byref(n), d, d, d, d, d, d, 0.01, 0.99, 1, d, 2, first, k, d, rows, d, d, d, d, d, 1 << 30,
                                   1, 512, 512, None)
    assert lib.pbmc_rollout(*args(2, 8, 8)) == -1   # rows 1..8 of an 8-row buffer: one past the end
    assert lib.pbmc_rollout(*args(1, 8, 7)) == -1


def test_python_wrappers_refuse_cpu_tensors(lib):
    import torch

    from pbml_mantle_convection_b200 import ops

    with pytest.raises(L.PbmcError):
        ops.pack_nchw(torch.zeros(1, 4, 8, 8))
