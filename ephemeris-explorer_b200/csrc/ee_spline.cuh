// ee_spline.cuh -- device-side evaluation of the piecewise-polynomial ephemeris, reference operation order.
//   UniformSpline::get_polynomial      ephemeris/src/trajectory.rs:561-569 (+ index rules :600-617, span :625-627)
//   Polynomial::eval / eval_and_deriv  ephemeris/src/trajectory.rs:360-385, eval_slice_horner :398-410
//   UniformSpline::state_vector        ephemeris/src/trajectory.rs:455-470
#pragma once
#include "ee_common.cuh"

namespace ee {

struct EphemView {
    int64_t nb;
    const double* mu;
    const double* start;
    const double* interval;
    const int64_t* npoly;
    const int64_t* first;
    const double* coef;    // [total][9][3]
    const int32_t* ncoef;  // [total]
};

// Some((polynomial index, normalised time)) or None
__device__ __forceinline__ bool spline_locate(const EphemView& e, int64_t b, double at, int64_t* pidx, double* tau) {
    const double start = e.start[b], interval = e.interval[b];
    const int64_t np = e.npoly[b];
    const double local = xsub(at, start);
    const double span = xmul(interval, (double)np);
    if (signbit(local) || local > span) return false;  // time.is_negative() || time > self.span()
    const double q = ceil(xdiv(local, interval));
    int64_t idx = (int64_t)q;  // `as usize` (q >= 0)
    idx = idx > 0 ? idx - 1 : 0;  // saturating_sub(1)
    const double lp = xsub(local, xmul(interval, (double)idx));
    *tau = xdiv(lp, interval);
    if (idx >= np) return false;  // polynomials.get(idx)?
    *pidx = e.first[b] + idx;
    return true;
}

__device__ __forceinline__ D3 ld_coef(const double* c, int i) { return {c[3 * i], c[3 * i + 1], c[3 * i + 2]}; }

// Polynomial::eval: result = 0; for c in coeffs.rev(): result = result * t + c
__device__ __forceinline__ D3 poly_eval(const double* c, int nc, double t) {
    D3 r = {0.0, 0.0, 0.0};
    for (int i = nc - 1; i >= 0; --i) r = xadd3(xmul3(r, t), ld_coef(c, i));
    return r;
}

// Polynomial::eval_and_deriv
__device__ __forceinline__ void poly_eval_and_deriv(const double* c, int nc, double t, D3* ev, D3* de) {
    const D3 zero = {0.0, 0.0, 0.0};
    const D3 first = nc ? ld_coef(c, 0) : zero;
    const D3 last = nc ? ld_coef(c, nc - 1) : zero;
    D3 eval = last, deriv = last;
    for (int i = nc - 2; i >= 1; --i) {
        eval = xadd3(xmul3(eval, t), ld_coef(c, i));
        deriv = xadd3(xmul3(deriv, t), eval);
    }
    eval = xadd3(xmul3(eval, t), first);
    *ev = eval;
    *de = deriv;
}

__device__ __forceinline__ bool spline_position(const EphemView& e, int64_t b, double at, D3* pos) {
    int64_t p;
    double tau;
    if (!spline_locate(e, b, at, &p, &tau)) return false;
    *pos = poly_eval(e.coef + 27 * p, e.ncoef[p], tau);
    return true;
}

__device__ __forceinline__ bool spline_state_vector(const EphemView& e, int64_t b, double at, D3* pos, D3* vel) {
    int64_t p;
    double tau;
    if (!spline_locate(e, b, at, &p, &tau)) return false;
    D3 ev, de;
    poly_eval_and_deriv(e.coef + 27 * p, e.ncoef[p], tau, &ev, &de);
    *pos = ev;
    *vel = xdiv3(de, e.interval[b]);  // dx/dt = dx/dtau / interval
    return true;
}

}  // namespace ee
