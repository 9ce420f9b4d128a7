"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/mpbp.h
declares, its host-only entry points work without a GPU, and GPU entry points fail loudly."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge._load_build_module().build()
    from mp_block_preconditioners_b200 import _cabi
    return _cabi.load()


def test_every_declared_symbol_is_exported(lib):
    from mp_block_preconditioners_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "mpbp.h")).read()
    declared = set(re.findall(r"\b(mpbp_[A-Za-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header(lib):
    """sizeof via a default-filled struct: the fields written by the C side land where ctypes reads them."""
    from mp_block_preconditioners_b200._cabi import Config, GmresOpts
    cfg = Config()
    assert lib.mpbp_config_default(C.byref(cfg)) == 0
    # reference defaults: solve.py:291-297, preconditioner.py:299
    assert (cfg.n, cfg.xi, cfg.eta_n, cfg.eta_s, cfg.c, cfg.d_u, cfg.d_p, cfg.d_div) == (16, 1.0, 100.0, 1.0, 1.0, -1.0, 1.0, -1.0)
    assert (cfg.rank, cfg.nranks, cfg.omega, cfg.nu1, cfg.nu2, cfg.n_coarse) == (0, 1, 0.8, 2, 2, 4)
    assert cfg.operators_only == 0 and cfg.workspace is None and cfg.workspace_bytes == 0
    o = GmresOpts()
    assert lib.mpbp_gmres_opts_default(C.byref(o)) == 0
    assert (o.rtol, o.restart, o.maxiter, o.side, o.use_precond) == (1e-8, 20, 150, 1, 1)


def test_workspace_sizing_and_argument_errors(lib):
    from mp_block_preconditioners_b200._cabi import Config
    cfg = Config()
    lib.mpbp_config_default(C.byref(cfg))
    need = C.c_size_t()
    sizes = []
    for n in (16, 64, 256, 4096):
        cfg.n = n
        assert lib.mpbp_plan_workspace_bytes(C.byref(cfg), C.byref(need)) == 0
        sizes.append(need.value)
    assert sizes == sorted(sizes)
    # 4096^2: ~26 level-0 velocity-sized buffers; must fit comfortably in 180 GB
    assert 4e9 < sizes[-1] < 40e9
    cfg.n = 1
    assert lib.mpbp_plan_workspace_bytes(C.byref(cfg), C.byref(need)) == -1
    assert b"n must be" in lib.mpbp_last_error_string()
    cfg.n = 4094  # 2 x 2047: cannot be coarsened to a small dense problem
    assert lib.mpbp_plan_workspace_bytes(C.byref(cfg), C.byref(need)) == -1
    cfg.n = 64
    cfg.nranks, cfg.rank = 3, 0
    assert lib.mpbp_plan_workspace_bytes(C.byref(cfg), C.byref(need)) != 0  # no unique id / not divisible
    assert lib.mpbp_apply_A(None, None, None, None) == -1


def test_slab_level_shapes_multi_rank(lib):
    """Per-rank workspace of the slab decomposition shrinks ~1/P (host logic of the hierarchy builder)."""
    from mp_block_preconditioners_b200._cabi import Config
    idbuf = C.create_string_buffer(128)
    need = C.c_size_t()
    sizes = {}
    for P in (1, 2, 4, 8):
        cfg = Config()
        lib.mpbp_config_default(C.byref(cfg))
        cfg.n, cfg.nranks, cfg.rank = 4096, P, P - 1
        cfg.nccl_unique_id = C.cast(idbuf, C.c_void_p)
        assert lib.mpbp_plan_workspace_bytes(C.byref(cfg), C.byref(need)) == 0, lib.mpbp_last_error_string()
        sizes[P] = need.value
    # (slab-distributed levels carry no omega/diag array for the fused pre-smoothing, so a rank's share is a
    # little below 1/P of the single-GPU workspace)
    for P in (2, 4, 8):
        assert 0.85 < sizes[P] * P / sizes[1] < 1.10, sizes


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mp_block_preconditioners_b200._cabi import Config
    cfg = Config()
    lib.mpbp_config_default(C.byref(cfg))
    h = C.c_void_p()
    rc = lib.mpbp_plan_create(C.byref(h), C.byref(cfg))
    assert rc != 0 and not h.value
    import mp_block_preconditioners_b200 as mp
    with pytest.raises(RuntimeError):
        mp.MultiphaseBlockPreconditioner(16, 1.0, 100.0, 1.0).get_big_A_matrix(1, -1)


def test_product_does_not_import_oracle():
    """The product package must not reference oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "mp-block-preconditioners_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mpbp_oracle" not in src and "oracle/" not in src, f


def test_get_thn_vals_matches_reference_table():
    """MultiphaseBlockPreconditioner.get_thn_vals (preconditioner.py:26-84): the six volume fractions around every
    u-face of an 8 x 8 grid, both phases, vs the table produced by the reference's own method (host-only code)."""
    import numpy as np
    import mp_block_preconditioners_b200 as mp
    from conftest import golden
    kat = golden("known_answers.npz")
    if "thn_vals_n8" not in kat.files:
        import pytest
        pytest.skip("fixture predates get_thn_vals")
    bp = mp.MultiphaseBlockPreconditioner(8, 1.0, 1.0, 1.0)
    got = np.array([[[bp.get_thn_vals(8, r, c, bool(ph)) for c in range(8)] for r in range(8)] for ph in (0, 1)])
    assert np.abs(got - kat["thn_vals_n8"]).max() < 1e-15
