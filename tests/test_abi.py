"""The C-ABI library loads and exports every symbol include/glc_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import subprocess

import pytest

from galacticus_b200 import abi


@pytest.fixture(scope="module")
def lib_path():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "galacticus_b200", "libglcb200.so")
    if not os.path.exists(path):
        subprocess.run([os.path.join(root, "build.sh")], check=True)
    return path


def test_exports_every_declared_symbol(lib_path):
    L = C.CDLL(lib_path)
    assert len(abi.DECLARED_FUNCTIONS) >= 15
    for name in abi.DECLARED_FUNCTIONS:
        assert hasattr(L, name), f"{name} declared in include/glc_b200.h but not exported"


def test_abi_version_and_defaults(lib_path):
    L = C.CDLL(lib_path)
    L.glc_abi_version.restype = C.c_int
    assert L.glc_abi_version() == abi.GLC_ABI_VERSION
    p = abi.glc_params()
    assert L.glc_params_default(C.byref(p), abi.GLC_MODEL_STANDARD) == 0
    # parameters/quickTest.xml
    assert p.odeToleranceAbsolute == 0.01 and p.odeToleranceRelative == 0.01
    assert p.recycledFraction == 0.46 and p.metalYield == 0.035
    assert p.fbDiskVelocityCharacteristic == 250.0 and p.fbSpheroidVelocityCharacteristic == 100.0
    assert L.glc_params_default(C.byref(p), 99) != 0


def test_defaults_match_oracle(lib_path, oracle_lib):
    L = C.CDLL(lib_path)
    for model in (abi.GLC_MODEL_BOX, abi.GLC_MODEL_STANDARD):
        p = abi.glc_params()
        L.glc_params_default(C.byref(p), model)
        q = oracle_lib.params_default(model)
        for name, _ in abi.glc_params._fields_:
            assert getattr(p, name) == getattr(q, name), name


def test_no_device_fails_loudly(lib_path):
    """Without a GPU the product path must refuse to run, not fall back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = C.CDLL(lib_path)
    h = C.c_void_p()
    rc = L.glc_evolver_create(C.byref(h), 0)
    assert rc != 0 and not h.value


def test_product_does_not_import_oracle():
    """The product package may never route through oracle/ (it is the checker only)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "galacticus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read().lower()
                assert "oracle" not in src and "liborc" not in src, f"{f} references the oracle"


def test_fortran_interface_matches_header():
    """integration/B200_interface.F90 (the ISO_C_BINDING module of INTEGRATION.md) is generated from include/glc_b200.h:
    the committed file must be what the generator produces now, every struct field of the C-ABI must appear in its bind(c)
    type in header order, and the per-node status codes must be the reference's errorStatus* values
    (source/error/_module.F90:66-75 = GSL error codes)."""
    import importlib.util
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_fortran_interface", os.path.join(root, "scripts", "gen_fortran_interface.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(os.path.join(root, "integration", "B200_interface.F90")).read()
    assert committed == gen.generate(), "run scripts/gen_fortran_interface.py"
    block = re.search(r"type, bind\(c\) :: glcParams(.*?)end type glcParams", committed, flags=re.S).group(1)
    fields = re.findall(r"::\s*(\w+)", block)
    assert fields == [name for name, _ in abi.glc_params._fields_]
    assert abi.GLC_STATUS_SUCCESS == 0 and abi.GLC_STATUS_FAIL == -1 and abi.GLC_STATUS_UNDERFLOW == 15
    assert abi.GLC_STATUS_XCPU == 1025
    for f in ("node_evolver_B200.F90", "evolver_B200.F90"):
        src = open(os.path.join(root, "integration", f)).read()
        for call in re.findall(r"\b(glc_\w+)\s*\(", src):
            assert call in abi.DECLARED_FUNCTIONS, f"{f} calls {call}, which include/glc_b200.h does not declare"
