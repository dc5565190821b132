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


# ---- property fuzz (hypothesis) around the switch points of the helpers ----------------------------------------------
# rpow_ok(x): 1e-30 < x < 1e30 selects the Newton-seeded path, everything else the sqrt chain / pow(); ratio2: n < d
# selects the shared-denominator quotient, n >= d returns exactly 1.  The properties must hold on both sides of each switch
# and AT it (no jump larger than the stated bound when x crosses the switch).
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

_SET = dict(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
_near_switch = st.one_of(
    st.floats(min_value=0.5, max_value=2.0).map(lambda s: 1e-30 * s), st.floats(min_value=0.5, max_value=2.0).map(lambda s: 1e30 * s),
    st.sampled_from([1e-30, np.nextafter(1e-30, 1.0), np.nextafter(1e-30, 0.0), 1e30, np.nextafter(1e30, 0.0), np.nextafter(1e30, np.inf)]),
    st.floats(min_value=-38.0, max_value=38.0).map(lambda e: 10.0 ** e),   # the float32 range of the seed, both paths
    st.floats(min_value=1e-300, max_value=1e-39), st.floats(min_value=1e39, max_value=1e150))   # outside the FP32 range of the seed


@settings(**_SET)
@given(x=_near_switch)
def test_fuzz_dyadic_powers_across_the_path_switch(emu, x):
    x = float(x)
    for which, expo, sel in ((0, 1 / 16, 0.), (1, 1 / 8, 0.), (2, 13 / 16, 0.), (3, 7 / 8, 0.), (4, 21 / 32, 0.), (4, 0.5, 1.), (5, 11 / 16, 0.), (5, 0.5, 1.)):
        ref = x ** expo
        assert abs(emu.crm_emu_math(which, x, sel) / ref - 1.0) <= 24 * ULP, (which, x)


@settings(**_SET)
@given(x=_near_switch)
def test_fuzz_rain_and_ice_exponents_across_the_path_switch(emu, x):
    x = float(x)
    for which, c in ((6, float(np.float32(0.8)) / 4), (7, (3 + float(np.float32(0.8))) / 4), (8, (5 + float(np.float32(0.8))) / 8), (9, float(np.float32(0.24)))):
        # outside the seeded range the helper is exp(c*log x): its error grows with |c ln x| (2 ulp of the exponent argument)
        assert abs(emu.crm_emu_math(which, x, c) / x ** c - 1.0) <= max(3e-14, 2 * ULP * abs(c * np.log(x))), (which, x)


@settings(**_SET)
@given(d=st.floats(min_value=-10.0, max_value=3.0).map(lambda e: 10.0 ** e), k=st.integers(min_value=-6, max_value=6),
       scale=st.one_of(st.just(1.0), st.floats(min_value=1e-3, max_value=1e3)))
def test_fuzz_ratio2_at_the_limiter_switch(emu, d, k, scale):
    """crm_emu_math(12, x, c) = q1 + 1e3*q2 with (q1, q2) = ratio2(x, c, 3x, 7c+1): numerators within a few ulp of their denominators
    (the n < d switch), exactly equal, and far away; min(1, n/d) of the literal divisions is the reference (case 13)."""
    c = float(d)
    x = c * float(scale)
    for _ in range(abs(k)):
        x = float(np.nextafter(x, np.inf if k > 0 else 0.0))
    a, b = emu.crm_emu_math(12, x, c), emu.crm_emu_math(13, x, c)
    assert abs(a - b) <= 4 * ULP * abs(b), (x, c, a, b)
    if x >= c:                                   # first limiter inactive: exactly 1, no quotient
        q2 = min(1.0, (3 * x) / (7 * c + 1))
        assert abs((a - 1.0) - 1e3 * q2) <= 1e3 * 8 * ULP
    # monotone through the switch: q1(x) <= 1 always
    assert a - 1e3 * min(1.0, (3 * x) / (7 * c + 1)) <= 1.0 + 1e3 * 8 * ULP


@settings(**_SET)
@given(t=st.floats(min_value=185.0, max_value=320.0), p=st.floats(min_value=20.0, max_value=1080.0))
def test_fuzz_qsat_pair_including_the_cold_branch_and_e_over_p_clamp(emu, t, p):
    # sat.F90: dt > -80 polynomial / Goff-Gratch style cold branch at 193.16 K; max(e, p-e) clamp when e > p/2 (hot, low pressure)
    a, b = emu.crm_emu_math(10, float(t), float(p)), emu.crm_emu_math(11, float(t), float(p))
    assert abs(a - b) <= 4 * ULP * abs(b), (t, p, a, b)
