"""CPU-only: the device math helpers that replace IEEE divide / sqrt chains / exp(c*log x) in the CUDA kernels (MUFU-pipe
diet), compiled for the host by the dev emulator build (tests/emu) and compared with libm.
Bounds are the ones DESIGN.md states: <= 24 eps (5e-15 relative) for the dyadic powers, 3e-14 relative for the rain / ice exponents
(FP32-accurate ln x in the correction for the REAL(4) rounding of 0.8 and 0.24), 4 ulp for the one-division pairs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common as cm

ULP = 2.220446049250313e-16


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(cm.ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    lib = C.CDLL(cm.EMU_LIB)
    lib.crm_emu_math.restype = C.c_double
    lib.crm_emu_math.argtypes = [C.c_int, C.c_double, C.c_double]
    return lib


def _xs():
    rng = np.random.default_rng(42)
    return np.concatenate([10.0 ** rng.uniform(-14, 0, 4000), [1e-8, 1.000000001e-8, 1e-3, 0.5, 1.0]])


@pytest.mark.parametrize("which,expo,sel", [(0, 1 / 16, 0.), (1, 1 / 8, 0.), (2, 13 / 16, 0.), (3, 7 / 8, 0.), (4, 21 / 32, 0.), (4, 0.5, 1.),
                                            (5, 11 / 16, 0.), (5, 0.5, 1.)])
def test_dyadic_powers(emu, which, expo, sel):
    worst = max(abs(emu.crm_emu_math(which, float(x), sel) / float(x) ** expo - 1.0) for x in _xs())
    assert worst <= 24 * ULP, worst / ULP


def test_dyadic_powers_outside_fp32_range_use_sqrt_chain(emu):
    for x in (1e-40, 3e-33, 1e35):
        assert abs(emu.crm_emu_math(0, x, 0.) / x ** (1 / 16) - 1.0) <= 4 * ULP
        assert abs(emu.crm_emu_math(2, x, 0.) / x ** (13 / 16) - 1.0) <= 4 * ULP


@pytest.mark.parametrize("which,c", [(6, float(np.float32(0.8)) / 4), (7, (3 + float(np.float32(0.8))) / 4), (8, (5 + float(np.float32(0.8))) / 8),
                                     (9, float(np.float32(0.24)))])
def test_rain_and_ice_exponents(emu, which, c):
    worst = max(abs(emu.crm_emu_math(which, float(x), c) / float(x) ** c - 1.0) for x in _xs())
    assert worst <= 3e-14, worst


def test_one_division_pairs(emu):
    rng = np.random.default_rng(7)
    for _ in range(2000):
        t, p = float(rng.uniform(200., 310.)), float(rng.uniform(50., 1050.))
        a, b = emu.crm_emu_math(10, t, p), emu.crm_emu_math(11, t, p)
        assert abs(a - b) <= 4 * ULP * abs(b), (t, p, a, b)
        x, c = float(10.0 ** rng.uniform(-12, 2)), float(10.0 ** rng.uniform(-10, 3))
        a, b = emu.crm_emu_math(12, x, c), emu.crm_emu_math(13, x, c)   # ratio2 returns min(1, n/d) for both quotients
        assert abs(a - b) <= 4 * ULP * abs(b), (x, c, a, b)
    # zero numerators give exact zeros; ratios >= 1 come back as exactly 1 without a division
    assert emu.crm_emu_math(12, 0.0, 5.0) == 0.0
    assert emu.crm_emu_math(12, 50.0, 5.0) == 1.0 + 1e3 * 1.0
