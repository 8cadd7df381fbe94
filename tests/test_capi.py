"""The C ABI library loads and exports every symbol include/emulator_b200.h declares (CPU)."""
import os
import re

import numpy as np
import pytest

from emulator_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "emulator_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.eb200_version() >= 100


def test_result_length_macro_matches_binding():
    text = open(os.path.join(ROOT, "include", "emulator_b200.h")).read()
    assert "#define EB200_RESULT_LEN(d) ((d) + 5)" in text
    assert _lib.result_len(9) == 14
    assert f"#define EB200_MAX_DIM {_lib.MAX_DIM}" in text


def test_argument_validation_without_gpu():
    lib = _lib.load()
    from ctypes import byref, c_void_p

    h = c_void_p()
    x = np.zeros((4, 2))
    noise = np.zeros(4)
    assert lib.eb200_gp_create(byref(h), 0, 0, 2, _lib.as_double_ptr(x), _lib.as_double_ptr(noise)) != 0
    assert "n must be" in _lib.last_error()
    assert lib.eb200_gp_create(byref(h), 0, 4, 17, _lib.as_double_ptr(x), _lib.as_double_ptr(noise)) != 0
    assert "dimension" in _lib.last_error()
    assert lib.eb200_gp_destroy(None) == 0


@pytest.mark.skipif(_lib.load().eb200_device_count() > 0, reason="a GPU is visible")
def test_product_fails_loudly_without_a_gpu():
    """No CPU fallback: building a GP without a CUDA device must raise, not degrade."""
    from emulator_b200 import george_compat as george
    from emulator_b200.device import DeviceProblem

    with pytest.raises(_lib.NativeLibraryError):
        DeviceProblem(np.zeros((4, 2)), np.ones(4))
    gp = george.GP(1.0 * george.kernels.ExpSquaredKernel(np.ones(2), ndim=2))
    with pytest.raises(_lib.NativeLibraryError):
        gp.compute(np.zeros((4, 2)), 0.1)


def test_product_never_imports_the_oracle():
    import ast

    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "emulator_b200")):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(base, f)).read())
                for node in ast.walk(tree):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        mods = [node.module]
                    bad += [(f, m) for m in mods if m.split(".")[0] == "oracle"]
    assert bad == []


def test_kernel_expression_guard():
    from emulator_b200.george_compat import kernels

    k = 1 ** 2 * kernels.ExpSquaredKernel(np.ones(3), ndim=3)
    assert kernels.device_form(k) == 3
    assert np.allclose(k.get_parameter_vector(), [np.log(1 / 3), 0, 0, 0])
    assert k.get_parameter_names()[0] == "k1:log_constant" and k.get_parameter_names()[3] == "k2:metric:log_M_2_2"
    import copy

    k2 = copy.deepcopy(k)
    k2.set_parameter_vector(np.arange(4.0))
    assert np.allclose(k.get_parameter_vector(), [np.log(1 / 3), 0, 0, 0])
    with pytest.raises(NotImplementedError):
        kernels.device_form(kernels.ExpSquaredKernel(np.ones(3), ndim=3))
    with pytest.raises(NotImplementedError):
        kernels.ExpSquaredKernel(1.0, ndim=3)
