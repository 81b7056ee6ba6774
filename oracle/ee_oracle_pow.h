// ee_oracle_pow.h -- ORACLE copy (test infrastructure) of the portable pow the GPU engine uses; selectable against
// libm's pow with ora_set_pow_mode().  Same operations in the same order as the device version.
#pragma once
#include <cmath>
namespace ora_pow {
#define PQUAL static inline
#define PFMA(a, b, c) std::fma((a), (b), (c))
#define PFREXP(x, e) std::frexp((x), (e))
#define PFLOOR(x) std::floor(x)
#define PLDEXP(x, n) std::ldexp((x), (n))
#define PNAME pow_portable
// Portable pow(x, y) for x > 0 built from IEEE-754 +, -, *, / and fma only (double-double arithmetic), so that the
// CPU and the GPU produce the SAME bits.  Why it exists: the ship step-size controller computes
// `fac * err.powf(-1/7)` (integration/src/runge_kutta/mod.rs:238); Rust's powf is the platform libm's pow, whose last
// bit is not portable (glibc vs CUDA libdevice differ), and the embedded error estimate cancels ~8 digits, so a 1-ulp
// difference in one step factor grows to ~1e-8 in the next -- the accepted-step sequence is only reproducible if pow
// is.  Accuracy: log via atanh series and exp via Taylor series in double-double, relative error < 2^-60 before the
// final rounding, i.e. the result is the correctly rounded value except when the exact result lies within ~2^-8 ulp
// of a rounding boundary.
struct PDD {
    double hi, lo;
};
PQUAL PDD p_two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return PDD{s, e};
}
PQUAL PDD p_quick_two_sum(double a, double b) {
    double s = a + b;
    double e = b - (s - a);
    return PDD{s, e};
}
PQUAL PDD p_two_prod(double a, double b) {
    double p = a * b;
    double e = PFMA(a, b, -p);
    return PDD{p, e};
}
PQUAL PDD p_add(PDD a, PDD b) {
    PDD s = p_two_sum(a.hi, b.hi);
    double e = s.lo + (a.lo + b.lo);
    return p_quick_two_sum(s.hi, e);
}
PQUAL PDD p_add_d(PDD a, double b) {
    PDD s = p_two_sum(a.hi, b);
    double e = s.lo + a.lo;
    return p_quick_two_sum(s.hi, e);
}
PQUAL PDD p_mul(PDD a, PDD b) {
    PDD p = p_two_prod(a.hi, b.hi);
    double e = p.lo + (a.hi * b.lo + a.lo * b.hi);
    return p_quick_two_sum(p.hi, e);
}
PQUAL PDD p_mul_d(PDD a, double b) {
    PDD p = p_two_prod(a.hi, b);
    double e = p.lo + a.lo * b;
    return p_quick_two_sum(p.hi, e);
}
PQUAL PDD p_div(PDD a, PDD b) {
    double q1 = a.hi / b.hi;
    PDD r = p_add(a, p_mul_d(b, -q1));
    double q2 = r.hi / b.hi;
    r = p_add(r, p_mul_d(b, -q2));
    double q3 = r.hi / b.hi;
    PDD q = p_quick_two_sum(q1, q2);
    return p_add_d(q, q3);
}
PQUAL PDD p_div_d(PDD a, double b) { return p_div(a, PDD{b, 0.0}); }

PQUAL double PNAME(double x, double y) {
    if (x != x || y != y) return x + y;
    if (y == 0.0) return 1.0;
    if (x == 0.0) return y < 0.0 ? 1.0 / 0.0 : 0.0;
    if (x < 0.0) return 0.0 / 0.0;          // not needed by the controller (err >= 0)
    if (x > 1.7976931348623157e308) return y < 0.0 ? 0.0 : x;
    int e = 0;
    double m = PFREXP(x, &e);               // x = m * 2^e, m in [0.5, 1)
    if (m < 0.70710678118654757) {
        m = m * 2.0;
        e -= 1;
    }                                        // m in [sqrt(1/2), sqrt(2))
    // ln m = 2 atanh(s), s = (m - 1)/(m + 1), |s| <= 0.1716
    PDD num = p_two_sum(m, -1.0);
    PDD den = p_two_sum(m, 1.0);
    PDD s = p_div(num, den);
    PDD s2 = p_mul(s, s);
    PDD acc = p_div_d(PDD{1.0, 0.0}, 29.0);
    for (int k = 13; k >= 0; --k) {
        acc = p_mul(acc, s2);
        acc = p_add(acc, p_div_d(PDD{1.0, 0.0}, (double)(2 * k + 1)));
    }
    PDD lnm = p_mul(p_mul_d(s, 2.0), acc);
    const PDD inv_ln2 = PDD{0x1.71547652b82fep+0, 0x1.777d0ffda0d24p-56};
    const PDD ln2 = PDD{0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56};
    PDD l2 = p_add_d(p_mul(lnm, inv_ln2), (double)e);  // log2(x)
    PDD t = p_mul_d(l2, y);
    if (t.hi > 1100.0) return 1.0 / 0.0;
    if (t.hi < -1100.0) return 0.0;
    double n = PFLOOR(t.hi + 0.5);
    PDD r = p_add_d(t, -n);                  // |r| <= 0.5
    PDD u = p_mul(r, ln2);
    PDD ex = PDD{1.0, 0.0};
    for (int k = 20; k >= 1; --k) {          // exp(u) = 1 + u/1 (1 + u/2 (1 + ... ))
        ex = p_mul(ex, p_div_d(u, (double)k));
        ex = p_add_d(ex, 1.0);
    }
    return PLDEXP(ex.hi + ex.lo, (int)n);
}
#undef PQUAL
#undef PFMA
#undef PFREXP
#undef PFLOOR
#undef PLDEXP
#undef PNAME
}  // namespace ora_pow
