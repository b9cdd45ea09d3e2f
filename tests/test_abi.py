"""The C-ABI library builds, loads on a CPU-only machine and exports every symbol declared in
include/spyb200.h (no compute calls without a GPU).  Also: the product never imports the oracle."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def libpath():
    from syncopy_b200 import build
    return build.build()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spyb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(spyb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_header_symbols(libpath):
    lib = ctypes.CDLL(libpath)
    names = _declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in spyb200.h but not exported"


def test_python_prototypes_cover_header(libpath):
    from syncopy_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()
    lib = _lib.load()
    assert lib.spyb_version() == 100
    assert lib.spyb_max_fft_len(1) == 16384


def test_no_gpu_fails_loudly(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from syncopy_b200 import _lib
    from syncopy_b200.engine import Engine
    with pytest.raises(_lib.SpybError):
        Engine(0)
    with pytest.raises(_lib.SpybError):
        _lib.init(0)
    import numpy as np
    from syncopy_b200 import compute_functions as cf
    x = np.zeros((64, 2), dtype=np.float32)
    # the dry run needs no device ...
    shp, dt = cf.cross_spectra_cF(x, 100., noCompute=True)
    assert shp == (1, 33, 2, 2)
    # ... the compute call must raise instead of silently falling back to NumPy
    with pytest.raises(_lib.SpybError):
        cf.cross_spectra_cF(x, 100.)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "syncopy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "from .. import oracle" not in src and "/root/reference" not in src, f
