"""The engine's restatement of glibc's pow (csrc/ee_pow_glibc.h) against the live libm of this machine (CPU only).

The ship step-size controller's `err.powf(-1/7)` is libm's pow in the reference (Rust's f64::powf); the GPU evaluates the
same algorithm with the same tables.  Here the HOST build of that header is compared bit for bit with libm.so.6's pow on
tens of millions of inputs: the controller's own domain, values around 1, subnormals, random exponents.  If the installed
libm is not glibc's table-driven pow (or its tables changed) this test fails instead of the GPU silently drifting."""
import ctypes
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HDR = ROOT / "ephemeris-explorer_b200" / "csrc" / "ee_pow_glibc.h"

SRC = r'''
#include <cmath>
#include <cstdint>
#include <cstring>
#include "%s"
static uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static double from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static uint64_t rng;
static uint64_t xs() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }
extern "C" double ee_pow_glibc_host(double x, double y) { return ee::pow_glibc(x, y); }
// mode 0: x log-uniform in [e^-span, e^span], fixed y;  mode 1: x within 2^20 ulp of 1, fixed y;
// mode 2: random bit patterns for x (all finite positive doubles incl. subnormals), y uniform in [-8, 8], restricted to
//         the exactness domain |y log x| < 500
extern "C" int64_t ee_pow_compare(int64_t n, uint64_t seed, int mode, double y0, double span, double* first_bad) {
    rng = seed ? seed : 88172645463325252ull;
    int64_t bad = 0;
    for (int64_t k = 0; k < n; ++k) {
        double x, y = y0;
        const double u = (double)(xs() >> 11) / 9007199254740992.0;
        if (mode == 0) {
            x = std::exp((2.0 * u - 1.0) * span);
        } else if (mode == 1) {
            x = from_bits(bits(1.0) + (xs() %% 2097152) - 1048576);
        } else {
            x = from_bits(xs() & 0x7fffffffffffffffull);
            y = (2.0 * u - 1.0) * 8.0;
            if (!(x == x) || !(x < 1.7e308) || !(std::fabs(y * std::log(x)) < 500.0)) continue;
        }
        const double a = std::pow(x, y), b = ee::pow_glibc(x, y);
        if (bits(a) != bits(b) && !(a != a && b != b)) {
            if (bad == 0 && first_bad) { first_bad[0] = x; first_bad[1] = y; first_bad[2] = a; first_bad[3] = b; }
            ++bad;
        }
    }
    return bad;
}
'''


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("pow")
    src = d / "pow_host.cpp"
    src.write_text(SRC % HDR)
    so = d / "libpowhost.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(so), "-lm"])
    lib = ctypes.CDLL(str(so))
    lib.ee_pow_glibc_host.restype = ctypes.c_double
    lib.ee_pow_glibc_host.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.ee_pow_compare.restype = ctypes.c_int64
    lib.ee_pow_compare.argtypes = [ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                   ctypes.POINTER(ctypes.c_double)]
    return lib


def compare(lib, n, seed, mode, y, span=0.0):
    bad = (ctypes.c_double * 4)()
    k = lib.ee_pow_compare(n, seed, mode, y, span, bad)
    assert k == 0, "%d mismatches, first: x=%r y=%r libm=%r ours=%r" % (k, bad[0], bad[1], bad[2], bad[3])


def test_controller_domain_is_bit_exact(lib):
    """err.powf(-1/k) for the embedded orders of the reference's adaptive methods (k = 7 for Verner87; 4..8 for the others),
    err log-uniform over e^+-300 (every finite positive err a step can produce)."""
    for k in (7, 4, 5, 6, 8):
        compare(lib, 4_000_000, k, 0, -1.0 / k, 300.0)
    compare(lib, 4_000_000, 99, 0, -1.0 / 7.0, 3.0)   # the values a healthy controller actually sees: err ~ 1e-1..1e1
    compare(lib, 2_000_000, 7, 1, -1.0 / 7.0)          # err within 1e-10 of 1


def test_other_exponents_and_random_arguments(lib):
    for y in (0.5, 1.0 / 3.0, -0.2, 2.5, -3.0, 7.0):
        compare(lib, 1_000_000, 1234, 0, y, 60.0)
    compare(lib, 10_000_000, 4321, 2, 0.0)


def test_special_values_follow_c99(lib):
    import math
    libm = ctypes.CDLL("libm.so.6")
    libm.pow.restype = ctypes.c_double
    libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    y = -1.0 / 7.0
    for x in (0.0, 1.0, math.inf, 5e-324, 1e-310, 2.2250738585072014e-308, 1.7976931348623157e308):
        a, b = libm.pow(x, y), lib.ee_pow_glibc_host(x, y)
        assert a == b or (math.isnan(a) and math.isnan(b)), (x, a, b)
    assert math.isnan(lib.ee_pow_glibc_host(math.nan, y))
    # outside the exactness domain (|y log x| >= 512) the result only has to clamp like glibc's would
    assert lib.ee_pow_glibc_host(1e300, -3.0) == 0.0 and lib.ee_pow_glibc_host(1e-300, -3.0) == math.inf
