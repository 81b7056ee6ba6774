"""Mathematical known-answer checks of the generated coefficient tables (oracle/ee_oracle_coeffs.h and the engine's copy):
order conditions that any correct transcription of the reference's methods must satisfy, independent of the reference
text.  Also checks that the two generated headers are identical."""
import re
from fractions import Fraction
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
HDR = ROOT / "oracle" / "ee_oracle_coeffs.h"


def table(name, text=None):
    text = text or HDR.read_text()
    m = re.search(r"%s\[\d+\] = \{[^\n]*\n(.*?)\};" % re.escape(name), text, re.S)
    assert m, name
    return np.array([float.fromhex(x) for x in re.findall(r"(-?0x[0-9a-f.]+p[+-]\d+),", m.group(1))])


def scalar(name):
    m = re.search(r"%s = (-?0x[0-9a-f.]+p[+-]\d+);" % re.escape(name), HDR.read_text())
    return float.fromhex(m.group(1))


def test_engine_and_oracle_headers_are_the_same_tables():
    a = HDR.read_text()
    b = (ROOT / "ephemeris-explorer_b200" / "csrc" / "ee_coeffs.h").read_text()
    assert a == b


def test_verner87_order_conditions():
    A = table("EE_V87_A").reshape(13, 13)
    B, C, E = table("EE_V87_B"), table("EE_V87_C"), table("EE_V87_E")
    assert np.allclose(A.sum(axis=1), C, atol=2e-15)           # row sums
    assert np.all(np.triu(A) == 0.0)                           # explicit
    Bh = B - E                                                 # embedded 7th-order weights
    for k in range(8):                                         # quadrature conditions sum b_i c_i^k = 1/(k+1)
        assert abs(B @ C**k - 1.0 / (k + 1)) < 5e-15, k        # order 8 main
    for k in range(7):
        assert abs(Bh @ C**k - 1.0 / (k + 1)) < 5e-15, k       # order 7 embedded
    assert abs(B @ (A @ C) - 1.0 / 6.0) < 5e-15                # a tree condition beyond pure quadrature
    assert abs(B @ (A @ (A @ C)) - 1.0 / 24.0) < 5e-15
    assert abs(Bh @ C**7 - 1.0 / 8.0) > 1e-6                   # ... and the embedded method really is of lower order


def test_blanes_moan_6b_is_a_symmetric_consistent_composition():
    a, b = table("EE_BM6B_A"), table("EE_BM6B_B")
    assert abs(a.sum() - 1.0) < 1e-15 and abs(b.sum() - 1.0) < 1e-15
    assert a[6] == 0.0 and np.array_equal(a[:6], a[5::-1]) and np.array_equal(b, b[::-1])  # palindromic, FSAL


def test_blanes_moan_14a_is_a_symmetric_consistent_composition():
    """methods.rs:1730-1774: 15 kick/drift pairs; kicks start with b_0 = 0 (which is what makes the FSAL skip of stage 0
    exact) and the scheme is time-symmetric: drifts palindromic a_s = a_{14-s}, kicks b_s = b_{15-s} for s >= 1."""
    a, b = table("EE_BM14A_A"), table("EE_BM14A_B")
    assert len(a) == 15 and len(b) == 15
    assert abs(a.sum() - 1.0) < 1e-15 and abs(b.sum() - 1.0) < 1e-15
    assert b[0] == 0.0 and np.array_equal(a, a[::-1]) and np.array_equal(b[1:], b[:0:-1])


def test_quinlan_tremaine_12_and_stormer_13_consistency():
    """y_{n+1} + sum_j alpha_j y_{n+1-j} = h^2 sum_j beta_j a_{n+1-j} must be exact for y = t^q, q = 0 .. order+1 (the
    tables hold exact integers, so this is checked in rational arithmetic)."""
    for tag, order in (("QT12", 12), ("ST13", 13)):
        nalpha = table("EE_%s_NEG_ALPHA" % tag)
        beta_n = table("EE_%s_BETA" % tag)
        beta_d = round(1.0 / scalar("EE_%s_INV_BETA_D" % tag))
        assert all(float(x).is_integer() for x in nalpha) and all(float(x).is_integer() for x in beta_n)
        assert scalar("EE_%s_INV_BETA_D" % tag) == 1.0 / float(beta_d)
        alpha = [Fraction(1)] + [Fraction(-int(x)) for x in nalpha]          # alpha_0 .. alpha_order
        beta = [Fraction(0)] + [Fraction(int(x), beta_d) for x in beta_n]    # beta_0 = 0 (explicit) .. beta_order
        for q in range(order + 2):
            lhs = sum(alpha[j] * Fraction(-j) ** q for j in range(order + 1))
            rhs = sum(beta[j] * q * (q - 1) * Fraction(-j) ** (q - 2) for j in range(1, order + 1)) if q >= 2 else Fraction(0)
            assert lhs == rhs, (tag, q, lhs, rhs)
        # ... and not beyond: the error constant is non-zero at q = order + 2
        q = order + 2
        lhs = sum(alpha[j] * Fraction(-j) ** q for j in range(order + 1))
        rhs = sum(beta[j] * q * (q - 1) * Fraction(-j) ** (q - 2) for j in range(1, order + 1))
        assert lhs != rhs, tag


def test_cowell_velocity_coefficients_sum():
    # dy = (y_n - y_{n-1})/h + h * sum c_j a_{n-j}: for constant acceleration the correction must be h*a/2
    for order in (12, 13):
        c = table("EE_COWELL%d_BETA" % order) * scalar("EE_COWELL%d_INV_BETA_D" % order)
        assert abs(c.sum() - 0.5) < 1e-14
        # exact for a(t) = t: y = t^3/6, velocity at t_n = t_n^2/2 (n = 0: t = 0): (0 - (-h)^3/6)/h + h*sum c_j*(-j h) = 0
        j = np.arange(order)
        assert abs(1.0 / 6.0 - (c * j).sum()) < 1e-13
