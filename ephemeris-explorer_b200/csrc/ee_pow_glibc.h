// ee_pow_glibc.h -- glibc's pow(x, y), restated operation by operation, for the ship step-size controller.
//
// Why: `IController::step` computes `fac * err.powf(-1/k)` (integration/src/runge_kutta/mod.rs:238) and Rust's f64::powf IS
// the platform libm's pow.  The embedded error estimate cancels ~8 digits, so one differing ulp in a step factor becomes
// 1e-8 in the next step size: to reproduce the reference AS BUILT on Linux/x86-64 the engine must produce glibc's bits.
// This is glibc >= 2.28's table-driven pow (sysdeps/ieee754/dbl-64/e_pow.c: log_inline -> exp_inline), in the FMA
// variant the ifunc resolver picks on every AVX2 machine (__pow_fma), INCLUDING the contractions GCC applied when that
// variant was compiled (read off the disassembly of the installed libm.so.6: t1, lo1, the log polynomial's outer
// multiply-add, elo, z + Shift, r, the exp polynomial and scale + scale*tmp are single fused operations).  Tables:
// ee_pow_glibc_tables.h, dumped from the installed libm by tools/gen_glibc_pow_tables.py.
// Pinned by tests/test_pow_cpu.py: the host build of this file equals the live libm's pow bit for bit on millions of
// inputs (the controller's domain y = -1/7, values around 1, subnormals, random x and y).
//
// Exactness domain: finite x > 0 (normal or subnormal) and |y log x| < 512, which holds for err^(-1/k) with any finite
// positive err.  Beyond it (results above 2^738 or below 2^-738) the function returns +inf / +0 where glibc returns a
// finite huge / tiny value; the controller clamps both to fac_max / fac_min identically.  x = 0, inf, NaN follow C99.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define EE_POW_TABLE static __device__ const
#define EE_POW_QUAL __device__ __forceinline__
#define EE_POW_BITS(x) ((uint64_t)__double_as_longlong(x))
#define EE_POW_FROM_BITS(u) __longlong_as_double((long long)(u))
#define EE_POW_FMA(a, b, c) fma((a), (b), (c))
#else
#include <math.h>
#include <string.h>
#define EE_POW_TABLE static const
#define EE_POW_QUAL static inline
static inline uint64_t ee_pow_bits_(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}
static inline double ee_pow_from_bits_(uint64_t u) {
    double x;
    memcpy(&x, &u, 8);
    return x;
}
#define EE_POW_BITS(x) ee_pow_bits_(x)
#define EE_POW_FROM_BITS(u) ee_pow_from_bits_(u)
#define EE_POW_FMA(a, b, c) fma((a), (b), (c))
#endif

namespace ee {
#include "ee_pow_glibc_tables.h"

EE_POW_QUAL double pow_glibc(double x, double y) {
    const double kInf = EE_POW_FROM_BITS(0x7ff0000000000000ull);
    uint64_t ix = EE_POW_BITS(x);
    const uint64_t iy = EE_POW_BITS(y);
    const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
        // special cases, restricted to what a step-size controller can produce (x >= 0 or NaN)
        if (x != x || y != y) return x + y;
        if (ix == 0x3ff0000000000000ull) return 1.0;
        if (2 * iy == 0) return 1.0;
        if (2 * ix == 0) return (iy >> 63) ? kInf : 0.0;  // pow(+-0, y)
        if (ix == 0x7ff0000000000000ull) return (iy >> 63) ? 0.0 : kInf;
        if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
            if ((topy & 0x7ff) < 0x3be) return 1.0;  // |y| < 2^-65
            return ((ix > 0x3ff0000000000000ull) == (topy < 0x800)) ? kInf : 0.0;
        }
        if (topx & 0x800) return EE_POW_FROM_BITS(0x7ff8000000000000ull);  // negative finite x: not needed here
        if (topx == 0) {  // subnormal x: normalise
            ix = EE_POW_BITS(x * 0x1p52);
            ix &= 0x7fffffffffffffffull;
            ix -= 52ull << 52;
        }
    }
    // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1), z in [0x1.69555p-1, 0x1.69555p0)
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> (52 - 7)) % 128);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffull << 52));
    const double z = EE_POW_FROM_BITS(iz), kd = (double)k;
    const double invc = kPowLogTab[i][0], logc = kPowLogTab[i][1], logctail = kPowLogTab[i][2];
    const double r = EE_POW_FMA(z, invc, -1.0);  // exact
    const double t1 = EE_POW_FMA(kd, kPowLn2Hi, logc);
    const double t2 = t1 + r;
    const double lo1 = EE_POW_FMA(kd, kPowLn2Lo, logctail);
    const double lo2 = t1 - t2 + r;
    const double ar = kPowLogPoly[0] * r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double hi = t2 + ar2;
    const double lo3 = EE_POW_FMA(ar, r, -ar2);
    const double lo4 = t2 - hi + ar2;
    const double q3 = EE_POW_FMA(r, kPowLogPoly[6], kPowLogPoly[5]);
    const double q2 = EE_POW_FMA(q3, ar2, EE_POW_FMA(r, kPowLogPoly[4], kPowLogPoly[3]));
    const double q1 = EE_POW_FMA(ar2, q2, EE_POW_FMA(r, kPowLogPoly[2], kPowLogPoly[1]));
    const double lo = EE_POW_FMA(ar3, q1, ((lo1 + lo2) + lo3) + lo4);
    const double lhi = hi + lo;
    const double llo = hi - lhi + lo;
    // ---- pow: y * log(x) in two pieces
    const double ehi = y * lhi;
    const double elo = EE_POW_FMA(y, llo, EE_POW_FMA(lhi, y, -ehi));
    // ---- exp_inline(ehi, elo)
    const uint32_t abstop = (uint32_t)(EE_POW_BITS(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + ehi;         // |y log x| < 2^-54
        return (EE_POW_BITS(ehi) >> 63) ? 0.0 : kInf;                 // |y log x| >= 512: see the header comment
    }
    double kd2 = EE_POW_FMA(ehi, kExpInvLn2N, kExpShift);
    const uint64_t ki = EE_POW_BITS(kd2);
    kd2 -= kExpShift;
    double rr = EE_POW_FMA(kd2, kExpNegLn2HiN, ehi);
    rr = EE_POW_FMA(kd2, kExpNegLn2LoN, rr);
    rr += elo;
    const unsigned idx = 2 * (unsigned)(ki % 128);
    const uint64_t top = ki << (52 - 7);
    const double tail = EE_POW_FROM_BITS(kExpTab[idx]);
    const uint64_t sbits = kExpTab[idx + 1] + top;
    const double r2 = rr * rr;
    const double p23 = EE_POW_FMA(rr, kExpPoly[1], kExpPoly[0]);
    const double tr = tail + rr;
    const double p45 = EE_POW_FMA(rr, kExpPoly[3], kExpPoly[2]);
    const double s1 = EE_POW_FMA(p23, r2, tr);
    const double tmp2 = EE_POW_FMA(p45, r2 * r2, s1);
    const double scale = EE_POW_FROM_BITS(sbits);
    return EE_POW_FMA(tmp2, scale, scale);
}
}  // namespace ee
#undef EE_POW_TABLE
#undef EE_POW_QUAL
#undef EE_POW_BITS
#undef EE_POW_FROM_BITS
#undef EE_POW_FMA
