"""The C-ABI library builds for sm_100a, loads without a GPU and exports every declared symbol."""
import ctypes as C
import os
import subprocess

import pytest

from graal_b200 import _lib


def test_library_builds_and_exports_all_declared_symbols():
    path = _lib.build()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in include/graal_b200.h but not exported" % name
    assert set(declared) == set(_lib._SIGS), "ctypes signatures out of sync with the header"
    assert b"sm_100a" in lib.graal_version()


def test_library_targets_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.graal_ctx_create(0, C.byref(ctx))
    assert rc != 0 and b"no CPU path" in lib.graal_last_error()
    from graal_b200.sampler import sampler, GraalError
    from graal_b200.level import build_synthetic_pyramid, prepare_sampler_inputs
    pyr = build_synthetic_pyramid([30_000, 20_000], 30, 2, seed=1, cis_rowsum=20.0, v_inter=0.01)
    with pytest.raises(GraalError):
        sampler.from_inputs(prepare_sampler_inputs(pyr, 1))


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "graal_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
