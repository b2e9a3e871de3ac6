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


def _primary_net_desc(levels=6, repeats=4, flags=0):
    """pbmc_net of the primary configuration with dummy (never dereferenced) weight pointers: enough for the host-side queries."""
    n = L.Net()
    n.levels, n.repeats, n.c_i, n.c_h, n.c_o, n.ksize = levels, repeats, 7, 16, 2, 3
    n.pad_mode, n.head_kind, n.p_pred, n.conv_impl, n.flags = L.PAD["replicate"], L.HEAD_CURL, 1, L.CONV_IMPL["auto"], flags
    def lay(cin_blks, cout):
        x = L.Layer()
        x.wpk = x.wpk_row = x.bias = x.gamma = x.beta = 16
        x.cin_blks, x.cout, x.ksize = cin_blks, cout, 3
        return x
    n.conv0 = lay(2, 16)
    for l in range(levels):
        for r in range(repeats):
            n.trunk[l * L.MAX_REPEATS + r] = lay(4, 16)
    n.conv1, n.conv2, n.conv3 = lay(4 * levels + 2, 16), lay(4, 16), lay(4, 2)
    return n


def test_trunk_cta_budgets_are_balanced_on_finish_times(lib):
    """Host-side scheduling of the pyramid levels (api.cu: level_cta_budgets): the persistent trunk kernels of all levels must
    be resident at once, so the budgets sum to <= 148; at 512^2 the level-0 kernel keeps five staging rounds (22 rows per
    CTA) and the coarse levels get what makes them finish, up-sampling included, with it."""
    n = _primary_net_desc()
    b = (C.c_int * L.MAX_LEVELS)()
    assert lib.pbmc_trunk_cta_budgets(C.byref(n), 1, 512, 512, b) == 1
    assert list(b)[:6] == [96, 30, 10, 5, 4, 2] and sum(b) <= 148
    for (B, H, W) in [(1, 128, 506), (1, 256, 256), (1, 128, 128), (1, 64, 96), (3, 64, 96), (2, 256, 256)]:
        assert lib.pbmc_trunk_cta_budgets(C.byref(n), B, H, W, b) == 1, (B, H, W)
        bud = list(b)[:6]
        assert all(x > 0 for x in bud) and sum(bud) <= 148, (B, H, W, bud)
        for l, x in enumerate(bud):  # whole column strips, and never more than one CTA per 8 rows (or the whole level)
            strips = B * (((W >> l) + 127) // 128)
            assert x % strips == 0 and x // strips <= max(1, -(-(H >> l) // 8)), (B, H, W, l, bud)
    # a 32-member ensemble does not fit 148 co-resident CTAs: one launch per layer, level 0 runs in waves (no budget)
    assert lib.pbmc_trunk_cta_budgets(C.byref(n), 32, 256, 256, b) == 0 and b[0] == 0
    # the per-layer flag turns the persistent kernels off and keeps the row-proportional shares
    nf = _primary_net_desc(flags=L.TRUNK_MODE["per_layer"])
    assert lib.pbmc_trunk_cta_budgets(C.byref(nf), 1, 512, 512, b) == 0 and list(b)[:6] == [108, 27, 6, 3, 2, 2]
    assert lib.pbmc_trunk_cta_budgets(None, 1, 512, 512, b) < 0
