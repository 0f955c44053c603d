"""The C-ABI library loads and exports exactly what include/staticfusion_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "staticfusion_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_loader_agree(sf_mod):
    from staticfusion_b200 import _lib
    assert header_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(sf_mod):
    L = C.CDLL(sf_mod.LIB_PATH)
    for name in header_functions():
        assert hasattr(L, name), name


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "staticfusion_b200.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "std::" not in src


def test_abi_version_and_default_params(sf_mod):
    from staticfusion_b200 import _lib
    assert _lib.lib().sf_abi_version() == 1
    p = sf_mod.default_params(240, 320)
    # StaticFusion-datasets.cpp:79-94 / FrontEnd.cpp:61,1130
    assert (p.rows, p.cols, p.ctf_levels, p.max_iter_per_level, p.max_iter_irls) == (240, 320, 5, 3, 6)
    assert p.use_motion_filter == 1 and p.enable_segmentation == 1
    assert abs(p.fovh - 62.5 * 3.141592653589793 / 180) < 1e-6
    for k, v in dict(k_photometric_res=0.15, irls_delta_threshold=0.0015, kc_cauchy=0.5, kb=1.5, kz=1.5, lambda_reg=0.35,
                     lambda_prior=0.5, previous_speed_const_weight=0.1, previous_speed_eig_weight=2.0,
                     outer_exit_threshold=0.04).items():
        assert abs(getattr(p, k) - v) < 1e-7, k
    assert sf_mod.default_params(480, 640).ctf_levels == 6 and sf_mod.default_params(960, 1280).ctf_levels == 7


@pytest.mark.parametrize("bad", [dict(rows=0), dict(ctf_levels=0), dict(ctf_levels=9), dict(cols=322), dict(max_iter_irls=0),
                                 dict(ctf_levels=1), dict(rows=250)])
def test_invalid_parameters_rejected_before_touching_cuda(sf_mod, bad):
    p = sf_mod.default_params(240, 320)
    for k, v in bad.items():
        setattr(p, k, v)
    with pytest.raises(sf_mod.SfError) as e:
        sf_mod.StaticFusionSolver(p)
    assert e.value.code == -1  # SF_E_INVALID


def test_fails_loudly_without_a_gpu(sf_mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sf_mod.SfError) as e:
        sf_mod.StaticFusionSolver(sf_mod.default_params(240, 320))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under staticfusion_b200/ may import, include, link or dlopen it."""
    pkg = os.path.join(ROOT, "staticfusion_b200")
    pat = re.compile(r"(from|import)\s+oracle|#include\s*[\"<][^\">]*oracle|sf_oracle|libsf_oracle|oracle/|orc_[a-z_]+\(")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(txt), (dirpath, f, pat.search(txt).group(0))


def test_division_free_pixel_split_is_exact():
    """sf_device.cuh split_rc: p / cols == (p * ceil(2^32 / cols)) >> 32 for every pixel of every level width the library
    accepts (cols % 4 == 0, up to 2048 x 2048), so its correction steps never fire."""
    import numpy as np

    for cols in (20, 40, 80, 160, 320, 640, 1280, 2048, 100, 36, 1000):
        magic = (2**32 + cols - 1) // cols
        assert magic < 2**32
        p = np.arange(0, min(cols * 2048, 2**22), dtype=np.uint64)
        q = (p * np.uint64(magic)) >> np.uint64(32)
        assert np.array_equal(q, p // np.uint64(cols)), cols
