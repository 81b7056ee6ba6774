// ee_ships.cu -- massless-ship propagator: the adaptive Runge-Kutta methods a flight plan can select, over the
// device-resident spline ephemeris, with the app's trajectory analytics.
//
// Reference functions restated here (file:line under the reference tree):
//   SpacecraftPropagator::{new, step}            ephemeris/src/propagators/spacecraft.rs:453-477, :599-615   (a17)
//   Timeline::new / segment_idx_at               ephemeris/src/propagators/spacecraft.rs:131-170
//   SpacecraftModel::eval (1st / 2nd order form) ephemeris/src/propagators/spacecraft.rs:283-332             (a13)
//   Bodies::acceleration / acceleration_at       ephemeris_explorer/src/dynamics/spacecraft.rs:71-74, :218-229 (a14)
//   ReferenceFrame::transform / TNB              ephemeris_explorer/src/dynamics/spacecraft.rs:240-293
//   ERK<C,[_;S]>::{advance, error, undo_step}    integration/src/runge_kutta/explicit.rs:73-140              (a15)
//   ERKNG<Fine45,..>::{advance, error}           integration/src/runge_kutta/nystrom/explicit_generalized.rs:77-175
//   IntegrationMethod (the eight tableaux)       ephemeris_explorer/src/flight_plan.rs:175-184, integration/src/methods.rs:92-1658
//   AdaptiveRungeKuttaIntegrator::advance        integration/src/runge_kutta/mod.rs:396-440                  (a16)
//   IController::step                            integration/src/runge_kutta/mod.rs:225-243
//   AbsTol::err_over_tol                         ephemeris_explorer/src/dynamics/spacecraft.rs:615-640
//   CubicHermiteSplineSolout                     ephemeris/src/propagators/spacecraft.rs:645-695             (a18)
//   SpacecraftSolout (SOI transitions, apsides)  ephemeris_explorer/src/dynamics/spacecraft.rs:76-161, :302-586
//   RelativeTrajectory::state_vector             ephemeris/src/trajectory.rs:315-335
//
// Mapping: one WARP per ship, lane b owns body b.  1 024 ships are 1.7 warps per scheduler, so the kernel is bound by the
// dependency chain of one right-hand side; the attempt is laid out to keep that chain short (see k_ships_step_to): body
// positions for all stage times up front and in lock-step, a per-warp polynomial cache in shared memory, running stage rows
// shared by the lanes, the 32 pulls added in body order by one lane per component (the reference's summation order).  Same
// operations as the reference in a different schedule: every knot and every event is bit-identical to the scalar path.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>

#include "ee_coeffs.h"
#include "ee_engine.h"
#include "ee_pow.cuh"
#include "ee_pow_glibc.h"
#include "ee_ships.h"
#include "ee_spline.cuh"

namespace ee {

// Every selectable adaptive method (ephemeris_explorer/src/flight_plan.rs:175-184) in one constant-memory layout; the
// kernel is instantiated per (stage count, FSAL, kind) and reads its coefficients from c_rk[method].
struct RkDev {
    int stages, fsal, kord, kind;  // kord = min(ORDER, ORDER_EMBEDDED); kind 0 = ERK, 1 = ERKNG (Fine45)
    double a[120], a2[120];        // strictly lower triangle, entry (s, j) at s(s-1)/2 + j; a2 = AV of an ERKNG method
    double b[16], b2[16], c[16], e[16], e2[16];
};
__constant__ RkDev c_rk[EE_RK_METHODS];

struct ShipParams {
    double h_init, h_max, tol_pos, tol_vel, fac_min, fac_max, fac;
    uint32_t n_max, pow_mode;
};

struct ShipsView {
    int64_t n;
    double* time;
    double* bound;
    double* state;  // [n][6]
    double* next_h;
    uint32_t* rk_i;
    uint32_t* n_att;
    int32_t* cur_seg;
    int32_t* status;
    int64_t* n_knots;
    unsigned long long* rhs_evals;
    const int64_t* seg_off;
    const int32_t* seg_burn;
    const double* seg_end;
    const double* seg_acc;
    const int32_t* seg_ref;
    double* knots;  // [n][kcap][7]
    int64_t kcap;
    double* fsal_k;  // [n][6] last slope of an FSAL method, carried from launch to launch
    // SpacecraftSolout analytics (optional): SOI transitions and apsides, per ship sorted by time
    int analytics;
    const double* soi_r;  // [nb]
    int64_t tr_cap, ap_cap;
    double* tr_time;      // [n][tr_cap]
    int32_t* tr_body;
    int32_t* n_tr;
    double* ap_time;      // [n][ap_cap]
    double* ap_dist;
    int32_t* ap_body;
    int32_t* ap_kind;     // 0 = periapsis, 1 = apoapsis
    int32_t* n_ap;
};

constexpr int kShipWarps = 4;
constexpr unsigned kFull = 0xffffffffu;

// Per-warp scratch in static shared memory.  P holds the running linear combinations of one attempt: row s < STAGES is the
// stage state y_s under construction, row STAGES the new state, row STAGES + 1 the embedded error (see k_ships_step_to).
struct WarpScratch {
    double P[EE_RK_MAX_STAGES + 2][6];
    double a[32][3];
    double k6[6];  // the current slope, for the lanes' row updates
};
// The method's tableau in "column" layout, built in shared memory once per CTA: m[s][r] is the coefficient with which
// slope k_s enters row r (a_rs for a later stage r, b_s for the new state, e_s for the error; 0 for r <= s).  The row
// updates read a different r per lane, which a constant-cache access would serialise; per attempt every warp keeps the
// table times h (the products `h * C::A[s][j]` the reference forms).
constexpr int kColRows = EE_RK_MAX_STAGES + 2;
struct ColTab {
    double m[EE_RK_MAX_STAGES * kColRows];
};

// DVec3::try_normalize
__device__ __forceinline__ bool try_normalize_dev(D3 v, D3* out) {
    const double r = xdiv(1.0, xsqrt(xdot3(v, v)));
    if (isfinite(r) && r > 0.0) {
        *out = xmul3(v, r);
        return true;
    }
    return false;
}

// Body positions at every stage time of one attempt, BEFORE the stages run: they depend on time only, not on the ship, so
// the 2 divisions + Horner chain per look-up leave the stage-to-stage dependency chain, and up to eight stage times are
// evaluated in lock-step (24 independent chains per lane instead of 1).  Same operations per look-up as
// UniformSpline::position (ee_spline.cuh), hence the same bits.
//
// Polynomial cache.  Lane b reading "its" polynomial straight from the table touches 32 different cache lines per load
// instruction (measured: a quarter of the kernel's stall samples sat in these loads).  Instead the warp keeps the current
// polynomial of each of its bodies in shared memory (pc[body][27], stride 27 doubles = conflict-free): when a lane's time
// has moved into another polynomial the WARP loads it with one coalesced 216-byte read (lanes 0..26).  A ship stays inside
// one polynomial of a body for many steps, so refills are rare; a stage time that falls into a different polynomial than
// the cached one (a step straddling a boundary) reads the table directly.
// bp is [STAGES][ngrp][3][32]; tag[g] is this lane's cached polynomial index of body g*32+lane (-1 = none); returns the
// per-lane mask of stages whose look-up succeeded for every body this lane owns.
template <int STAGES, int NG>
__device__ __forceinline__ unsigned ship_body_positions(const EphemView& E, double* __restrict__ bp, double* __restrict__ pc,
                                                        int64_t* tag, int* tag_nc, int ngrp, int lane, int s_first, double time,
                                                        double h, const double* __restrict__ cc) {
    // stage times evaluated in lock-step: all of them for short tableaux, two passes for the long ones (13 -> 7 + 6, 16 -> 8 + 8)
    constexpr int U = STAGES <= 8 ? STAGES : (STAGES + 1) / 2;
    unsigned okmask = 0xffffffffu;
#pragma unroll
    for (int g = 0; g < (NG ? NG : ngrp); ++g) {  // NG > 0: compile-time group count (tag[] stays in registers)
        const int64_t b = (int64_t)g * 32 + lane;
        const bool active = b < E.nb;
        const int64_t bb = active ? b : 0;
        const double start = E.start[bb], interval = E.interval[bb];
        const int64_t np = E.npoly[bb], first = E.first[bb];
        const double span = xmul(interval, (double)np);
        double* mine = pc + ((size_t)g * 32 + lane) * 27;
        for (int s0 = s_first; s0 < STAGES; s0 += U) {
            // UniformSpline::get_polynomial (spline_locate, ee_spline.cuh) for U times at once, straight-line: the two IEEE
            // divisions of each time are independent of the other times' and overlap in the pipe
            double local[U], q[U], tau[U];
            int64_t idx[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s = min(s0 + u, STAGES - 1);
                local[u] = xsub(xadd(time, xmul(h, cc[s])), start);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) q[u] = ceil(xdiv(local[u], interval));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in_span = !(signbit(local[u]) || local[u] > span);  // time.is_negative() || time > self.span()
                int64_t ix = in_span ? (int64_t)q[u] : 0;                      // `as usize` (q >= 0 inside the span)
                ix = ix > 0 ? ix - 1 : 0;                                      // saturating_sub(1)
                ok[u] = in_span && ix < np && s0 + u < STAGES;                 // polynomials.get(idx)?
                idx[u] = ok[u] ? ix : 0;
                tau[u] = xdiv(xsub(local[u], xmul(interval, (double)idx[u])), interval);
            }
            // refill the cache where this group's first time left the cached polynomial (warp-collective, coalesced)
            const bool want = active && ok[0] && idx[0] != tag[g];
            unsigned need = __ballot_sync(kFull, want);
            const int64_t my_poly = first + idx[0];
            if (need) {
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int64_t pi = __shfl_sync(kFull, my_poly, src);
                    if (lane < 27) pc[((size_t)g * 32 + src) * 27 + lane] = E.coef[27 * pi + lane];
                }
                if (want) {
                    tag[g] = idx[0];
                    tag_nc[g] = E.ncoef[my_poly];
                }
                __syncwarp();
            }
            const double* cf[U];
            int nc[U];
            int top = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool hit = idx[u] == tag[g];
                cf[u] = hit ? mine : E.coef + 27 * (first + idx[u]);
                nc[u] = !ok[u] ? 0 : (hit ? tag_nc[g] : E.ncoef[first + idx[u]]);
                top = max(top, nc[u]);
            }
            D3 r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) r[u] = d3(0.0, 0.0, 0.0);
            bool same = true;  // all U times of this lane fall into its cached polynomial (the usual case)
#pragma unroll
            for (int u = 0; u < U; ++u) same = same && (!ok[u] || idx[u] == tag[g]);
            if (__all_sync(kFull, same || !active)) {
                // one coefficient load per level serves the U Horner chains
                const int n0 = active ? tag_nc[g] : 0;
                for (int i = top - 1; i >= 0; --i) {  // Polynomial::eval: result = result * t + c, highest coefficient first
                    const D3 c = ld_coef(mine, i);
                    if (i < n0) {
#pragma unroll
                        for (int u = 0; u < U; ++u) r[u] = xadd3(xmul3(r[u], tau[u]), c);
                    }
                }
            } else {
                for (int i = top - 1; i >= 0; --i) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const D3 nr = xadd3(xmul3(r[u], tau[u]), ld_coef(cf[u], i));
                        if (i < nc[u]) r[u] = nr;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s = s0 + u;
                if (s < STAGES && active) {
                    if (!ok[u]) okmask &= ~(1u << s);
                    double* o = bp + ((size_t)(s * ngrp + g) * 3) * 32 + lane;
                    o[0] = r[u].x;
                    o[32] = r[u].y;
                    o[64] = r[u].z;
                }
            }
        }
    }
    return okmask;
}

// Bodies::acceleration at `pos` from the precomputed body positions of stage s: lane b's pull, then the ordered sum in
// body order.  Lane c < 3 carries component c's running sum through all groups (no per-lane select of a double: the
// compiler turns that into a divergent branch tree); one broadcast at the end.  Warp-collective.
template <int NG>
__device__ __forceinline__ D3 ship_context_acceleration(const EphemView& E, WarpScratch& ws, const double* __restrict__ bp, int ngrp,
                                                        const double* mu_l, int s, int lane, D3 pos) {
    double t = 0.0;  // lanes 0..2: x, y, z
#pragma unroll
    for (int g = 0; g < (NG ? NG : ngrp); ++g) {
        const int64_t b = (int64_t)g * 32 + lane;
        D3 a = {0.0, 0.0, 0.0};
        if (b < E.nb) {  // AccelerationAt::<false>: dir = src - pos; dir * (mu / (n * sqrt(n)))
            const double* q = bp + ((size_t)(s * ngrp + g) * 3) * 32 + lane;
            const D3 dir = xsub3(d3(q[0], q[32], q[64]), pos);
            const double nn = xdot3(dir, dir);
            const double sc = xdiv(mu_l[g], xmul(nn, xsqrt(nn)));
            a = xmul3(dir, sc);
        }
        ws.a[lane][0] = a.x;
        ws.a[lane][1] = a.y;
        ws.a[lane][2] = a.z;
        __syncwarp();
        if (lane < 3) {
            const int cnt = (int)min((int64_t)32, E.nb - (int64_t)g * 32);
            if (cnt == 32) {
#pragma unroll
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    double v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = ws.a[i0 + i][lane];  // eight loads in flight ahead of the dependent adds
#pragma unroll
                    for (int i = 0; i < 8; ++i) t = xadd(t, v[i]);
                }
            } else {
                for (int i = 0; i < cnt; ++i) t = xadd(t, ws.a[i][lane]);
            }
        }
        __syncwarp();
    }
    D3 sum;
    sum.x = __shfl_sync(kFull, t, 0);
    sum.y = __shfl_sync(kFull, t, 1);
    sum.z = __shfl_sync(kFull, t, 2);
    return sum;
}

// Segment::acceleration -> ConstantThrust::acceleration -> ReferenceFrame::transform (TNB).  Warp-uniform.
__device__ __forceinline__ bool ship_manoeuvre_acceleration(const EphemView& E, double ti, const double* yi, bool burn, D3 bacc, int bref,
                                                            D3* out) {
    D3 ma = {0.0, 0.0, 0.0};
    if (burn) {
        D3 cx, cy, cz;  // columns of DMat3::from_cols(x, z, y)
        if (bref < 0) {  // TNB::IDENTITY
            cx = d3(1.0, 0.0, 0.0);
            cy = d3(0.0, 1.0, 0.0);
            cz = d3(0.0, 0.0, 1.0);
        } else {
            D3 rp, rv;
            if (!spline_state_vector(E, bref, ti, &rp, &rv)) return false;
            const D3 relp = xsub3(d3(yi[0], yi[1], yi[2]), rp);
            const D3 relv = xsub3(d3(yi[3], yi[4], yi[5]), rv);
            D3 x, y;
            if (!try_normalize_dev(relv, &x)) return false;
            if (!try_normalize_dev(xcross3(relp, relv), &y)) return false;
            const D3 zc = xcross3(x, y);
            const D3 z = xmul3(zc, xdiv(1.0, xsqrt(xdot3(zc, zc))));
            cx = x;
            cy = z;
            cz = y;
        }
        // DMat3::mul_vec3: x_axis*v.x + y_axis*v.y + z_axis*v.z
        D3 r = xmul3(cx, bacc.x);
        r = xadd3(r, xmul3(cy, bacc.y));
        r = xadd3(r, xmul3(cz, bacc.z));
        ma = r;
    }
    *out = ma;
    return true;
}

// ---- SpacecraftSolout analytics (ephemeris_explorer/src/dynamics/spacecraft.rs:76-161, :536-586), warp-uniform code:
// every lane evaluates the same expressions (spline loads are broadcasts); the per-body sign tests of the SOI search are
// the one lane-parallel part.
struct HermiteD {  // CubicHermite::new / eval / eval_derivative -- ephemeris/src/trajectory.rs:645-698
    double t0, t1;
    D3 a0, a1, a2, a3;
};
__device__ __forceinline__ HermiteD hermite_make(double t0, const double* y0, double t1, const double* y1) {
    HermiteD H;
    H.t0 = t0;
    H.t1 = t1;
    const D3 p0 = d3(y0[0], y0[1], y0[2]), v0 = d3(y0[3], y0[4], y0[5]);
    const D3 p1 = d3(y1[0], y1[1], y1[2]), v1 = d3(y1[3], y1[4], y1[5]);
    H.a0 = p0;
    H.a1 = v0;
    H.a2 = d3(0.0, 0.0, 0.0);
    H.a3 = d3(0.0, 0.0, 0.0);
    const double dt = xsub(t1, t0);
    const bool same = p0.x == p1.x && p0.y == p1.y && p0.z == p1.z && v0.x == v1.x && v0.y == v1.y && v0.z == v1.z;
    if (!(dt == 0.0 && same)) {
        const double r = xdiv(1.0, dt), r2 = xmul(r, r), r3 = xmul(r, r2);
        const D3 dv = xsub3(p1, p0);
        H.a2 = xsub3(xmul3(xmul3(dv, r2), 3.0), xmul3(xadd3(xmul3(v0, 2.0), v1), r));
        H.a3 = xadd3(xmul3(xmul3(dv, r3), -2.0), xmul3(xadd3(v0, v1), r2));
    }
    return H;
}
__device__ __forceinline__ D3 hermite_eval(const HermiteD& H, double t) {
    const double d = xsub(t, H.t0);
    return xadd3(xmul3(xadd3(xmul3(xadd3(xmul3(H.a3, d), H.a2), d), H.a1), d), H.a0);
}
__device__ __forceinline__ D3 hermite_deriv(const HermiteD& H, double t) {
    const double d = xsub(t, H.t0);
    return xadd3(xmul3(xadd3(xmul3(xmul3(H.a3, d), 3.0), xmul3(H.a2, 2.0)), d), H.a1);
}
__device__ __forceinline__ double signum_f64(double x) { return isnan(x) ? x : (signbit(x) ? -1.0 : 1.0); }
// GravitationalBody::soi_distance_squared_at / radial_velocity_at
__device__ __forceinline__ bool ana_f(const EphemView& E, const double* soi_r, bool radial, int64_t b, const HermiteD& H, double t,
                                      double* out) {
    if (radial) {
        D3 bp, bv;
        if (!spline_state_vector(E, b, t, &bp, &bv)) return false;
        const D3 rp = xsub3(hermite_eval(H, t), bp), rv = xsub3(hermite_deriv(H, t), bv);
        *out = xdot3(rp, rv);
        return true;
    }
    D3 bp;
    if (!spline_position(E, b, t, &bp)) return false;
    const D3 d = xsub3(hermite_eval(H, t), bp);
    *out = xsub(xdot3(d, d), xmul(soi_r[b], soi_r[b]));
    return true;
}
// find_zero_crossing's bisection once f0, f1 are known to differ in sign: 0 = none, 1 = ascending, 2 = descending.
// The reference halves [x0, x1] one evaluation at a time (about 19 of them from a 400 s step down to 1 ms).  Here the warp
// evaluates FIVE levels at once: lanes 1..31 are the nodes of the binary tree below the current interval (heap order);
// each computes the end points its node would have -- with the reference's own expression mid = x0 + (x1 - x0) / 2 along
// its path, so they are the very numbers the sequential search would form -- and evaluates f at its midpoint; the warp then
// walks down the tree with the reference's sign rule and termination test.  Same mid points, same decisions, same result;
// four rounds instead of nineteen evaluations.  (x / 2.0 is formed as x * 0.5: identical for every double.)
__device__ int ana_bisect(const EphemView& E, const double* soi_r, bool radial, int64_t b, const HermiteD& H, double t0, double t1,
                          double f0, double* when) {
    const bool ascending = signbit(f0);
    const int lane = threadIdx.x & 31;
    const int depth = 31 - __clz(lane | 1);  // level of this lane's node (root = 0)
    double x0 = t0, x1 = t1;
    int it = 0;
    while (it < 100) {
        double a = x0, c = x1;
        for (int l = depth - 1; l >= 0; --l) {
            const double m = xadd(a, xmul(xsub(c, a), 0.5));
            if ((lane >> l) & 1)
                a = m;
            else
                c = m;
        }
        const double mid = xadd(a, xmul(xsub(c, a), 0.5));
        double fm = 0.0;
        if (lane >= 1) ana_f(E, soi_r, radial, b, H, mid, &fm);
        int node = 1;
        for (int lvl = 0; lvl < 5 && it < 100; ++lvl, ++it) {
            const double m = __shfl_sync(kFull, mid, node);
            const double f = __shfl_sync(kFull, fm, node);
            if (signum_f64(f0) != signum_f64(f)) {
                x1 = m;
                node = 2 * node;
            } else {
                x0 = m;
                f0 = f;
                node = 2 * node + 1;
            }
            if (fabs(xsub(x1, x0)) < 1e-3) {
                *when = x0;
                return ascending ? 1 : 2;
            }
        }
    }
    return 0;
}
// Bodies::soi_at_except + find_soi: closest body whose sphere contains `pos` (first one on ties), -1 = none
__device__ int ana_soi_at_except(const EphemView& E, const double* soi_r, double t, D3 pos, int64_t except) {
    int best = -1;
    double best_d = 0.0;
    for (int64_t b = 0; b < E.nb; ++b) {
        if (b == except) continue;
        D3 bp;
        if (!spline_position(E, b, t, &bp)) continue;
        const D3 d = xsub3(pos, bp);
        const double d2 = xdot3(d, d);
        if (!(d2 < xmul(soi_r[b], soi_r[b]))) continue;
        if (best < 0 || d2 < best_d) {
            best = (int)b;
            best_d = d2;
        }
    }
    return best;
}
// SoiTransitions::insert (lane 0 writes; every lane tracks the count)
// first index whose time is >= t (the Err/Ok position of the reference's binary_search); events arrive almost always in
// time order, so the end of the list is tried first
__device__ __forceinline__ int ana_lower_bound(const double* tt, int n, double t) {
    if (n == 0 || tt[n - 1] < t) return n;
    int lo = 0, hi = n - 1;  // tt[hi] >= t
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tt[mid] < t)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__device__ void ana_insert_transition(double* tt, int32_t* tb, int& n, double t, int body, int lane) {
    const int i = ana_lower_bound(tt, n, t);
    if (i < n && tt[i] == t) {
        if (lane == 0) tb[i] = body;
    } else if (i > 0 && tb[i - 1] == body) {
    } else {
        if (lane == 0) {
            for (int q = n; q > i; --q) {
                tt[q] = tt[q - 1];
                tb[q] = tb[q - 1];
            }
            tt[i] = t;
            tb[i] = body;
        }
        n += 1;
    }
    __syncwarp();
}
// One accepted step's analytics: k0 = (t0, y0) previous knot, k1 = (t1, y1) the knot just pushed.
// UniformSpline::position of this lane's body (group g) at two times in lock-step, reading the warp's polynomial cache
// where the time falls into the cached polynomial (it nearly always does: both times lie inside the step just taken) and
// the table otherwise.  Same operations as spline_position.
__device__ __forceinline__ void ana_positions2(const EphemView& E, const double* __restrict__ pc, const int64_t* tag, const int* tag_nc,
                                               int g, int lane, const double (&t)[2], bool (&ok)[2], D3 (&pos)[2]) {
    const int64_t b = (int64_t)g * 32 + lane;
    const double start = E.start[b], interval = E.interval[b];
    const int64_t np = E.npoly[b], first = E.first[b];
    const double span = xmul(interval, (double)np);
    double local[2], q[2], tau[2];
    const double* cf[2];
    int nc[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) local[u] = xsub(t[u], start);
#pragma unroll
    for (int u = 0; u < 2; ++u) q[u] = ceil(xdiv(local[u], interval));
    int top = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const bool in_span = !(signbit(local[u]) || local[u] > span);
        int64_t ix = in_span ? (int64_t)q[u] : 0;
        ix = ix > 0 ? ix - 1 : 0;
        ok[u] = in_span && ix < np;
        if (!ok[u]) ix = 0;
        tau[u] = xdiv(xsub(local[u], xmul(interval, (double)ix)), interval);
        const bool hit = ix == tag[g];
        cf[u] = hit ? pc + ((size_t)g * 32 + lane) * 27 : E.coef + 27 * (first + ix);
        nc[u] = !ok[u] ? 0 : (hit ? tag_nc[g] : E.ncoef[first + ix]);
        top = max(top, nc[u]);
        pos[u] = d3(0.0, 0.0, 0.0);
    }
    for (int i = top - 1; i >= 0; --i) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const D3 nr = xadd3(xmul3(pos[u], tau[u]), ld_coef(cf[u], i));
            if (i < nc[u]) pos[u] = nr;
        }
    }
}

// cur_t / cur_b mirror the LAST transition (tt[ntr-1], tb[ntr-1]) in registers: on a step without a crossing -- nearly all of
// them -- the sphere the ship is in is known without touching the lists in global memory.
__device__ __forceinline__ void ana_step(const ShipsView& S, const EphemView& E, const double* pc, const int64_t* tag, const int* tag_nc,
                                         int64_t ship, int lane, double t0, const double* y0, double t1, const double* y1, int& ntr,
                                         int& nap, double& cur_t, int& cur_b) {
    const HermiteD H = hermite_make(t0, y0, t1, y1);
    double* tt = S.tr_time + ship * S.tr_cap;
    int32_t* tb = S.tr_body + ship * S.tr_cap;
    bool touched = false;  // a transition was inserted during this step
    const D3 ship0 = hermite_eval(H, t0), ship1 = hermite_eval(H, t1);
    // SOI crossings, bodies in construction order: the end-point signs are found lane-parallel, a crossing is bisected
    // by the whole warp
    for (int64_t base = 0; base < E.nb; base += 32) {
        const int64_t b = base + lane;
        double f0 = 0.0, f1 = 0.0;
        bool cross = false;
        if (b < E.nb) {  // soi_distance_squared_at(t0), (t1): |ship - body|^2 - r^2
            const double tq[2] = {t0, t1};
            bool okq[2];
            D3 bq[2];
            ana_positions2(E, pc, tag, tag_nc, (int)(base / 32), lane, tq, okq, bq);
            const D3 d0 = xsub3(ship0, bq[0]), d1 = xsub3(ship1, bq[1]);
            const double rr = xmul(S.soi_r[b], S.soi_r[b]);
            f0 = xsub(xdot3(d0, d0), rr);
            f1 = xsub(xdot3(d1, d1), rr);
            cross = okq[0] && okq[1] && !(signum_f64(f0) == signum_f64(f1));
        }
        unsigned m = __ballot_sync(kFull, cross);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double g0 = __shfl_sync(kFull, f0, src);
            const int64_t bb = base + src;
            double when = 0.0;
            const int dir = ana_bisect(E, S.soi_r, false, bb, H, t0, t1, g0, &when);
            if (dir == 2) {
                ana_insert_transition(tt, tb, ntr, when, (int)bb, lane);
                touched = true;
            } else if (dir == 1) {
                const int entered = ana_soi_at_except(E, S.soi_r, when, hermite_eval(H, when), bb);
                if (entered >= 0) {
                    ana_insert_transition(tt, tb, ntr, when, entered, lane);
                    touched = true;
                }
            }
        }
    }
    if (touched && ntr > 0) {
        cur_t = tt[ntr - 1];
        cur_b = tb[ntr - 1];
    }
    // apsides inside every sphere occupied during the step: SoiTransitions::starting_at(t0).  Without a new transition and
    // with the last one at or before t0 that is the last entry alone.
    const bool simple = !touched && ntr > 0 && cur_t <= t0;
    int first = 0;
    if (simple) {
        first = ntr - 1;
    } else {
        const int i = ana_lower_bound(tt, ntr, t0);
        first = (i < ntr && tt[i] == t0) ? i : (i == 0 ? 0 : i - 1);
    }
    double* at = S.ap_time + ship * S.ap_cap;
    double* ad = S.ap_dist + ship * S.ap_cap;
    int32_t* ab = S.ap_body + ship * S.ap_cap;
    int32_t* ak = S.ap_kind + ship * S.ap_cap;
    for (int i = first; i < ntr; ++i) {
        const double ti = simple ? cur_t : tt[i];
        const double ta = ti > t0 ? ti : t0;
        const double tbnd = simple ? t1 : (i + 1 < ntr ? tt[i + 1] : t1);
        const int soi = simple ? cur_b : tb[i];
        // radial_velocity_at(ta) and (tbnd): lane 0 evaluates the first, the other lanes the second -- one evaluation's latency
        double fl = 0.0;
        const bool okl = ana_f(E, S.soi_r, true, soi, H, lane == 0 ? ta : tbnd, &fl);
        const unsigned okm = __ballot_sync(kFull, okl);
        if ((okm & 3u) != 3u) continue;
        const double f0 = __shfl_sync(kFull, fl, 0), f1 = __shfl_sync(kFull, fl, 1);
        if (signum_f64(f0) == signum_f64(f1)) continue;
        double when = 0.0;
        const int dir = ana_bisect(E, S.soi_r, true, soi, H, ta, tbnd, f0, &when);
        if (!dir) continue;
        D3 bp;
        if (!spline_position(E, soi, when, &bp)) continue;
        const D3 d = xsub3(bp, hermite_eval(H, when));
        const double dist = xsqrt(xdot3(d, d));
        // Apsides::insert
        const int q = ana_lower_bound(at, nap, when);
        const bool replace = q < nap && at[q] == when;
        if (lane == 0) {
            if (!replace)
                for (int r = nap; r > q; --r) {
                    at[r] = at[r - 1];
                    ad[r] = ad[r - 1];
                    ab[r] = ab[r - 1];
                    ak[r] = ak[r - 1];
                }
            at[q] = when;
            ad[q] = dist;
            ab[q] = soi;
            ak[q] = dir == 1 ? 0 : 1;
        }
        if (!replace) nap += 1;
        __syncwarp();
    }
}

// How the attempt is laid out (same operations as ERK::advance / ERKNG::advance, explicit.rs:73-106 and
// nystrom/explicit_generalized.rs:77-138, so the same bits; a different schedule):
//  * the reference rebuilds every stage state from y: y_s = (((y + k_0 (h a_s0)) + k_1 (h a_s1)) + ...).  Here every row
//    (the later stages, the new state, the error) is a running sum in shared memory; as soon as k_j exists, each lane adds
//    k_j's term to the rows it owns -- in j order, i.e. the reference's order -- so a stage waits for ONE multiply-add after
//    the previous slope instead of an s-term chain, and the 32 lanes share the work instead of repeating it;
//  * the body positions at the stage times do not depend on the ship: ship_body_positions() evaluates them for the whole
//    attempt up front, four stage times in lock-step;
//  * what is left on the stage-to-stage critical path is the pull of the bodies (one sqrt, one division) and the 32-term
//    ordered sum the reference's summation order dictates.
template <int STAGES, bool FSAL, int KIND, bool ANA, int NG>
__global__ void __launch_bounds__(kShipWarps * 32, 2) k_ships_step_to(ShipsView S, EphemView E, ShipParams P, int method, int ngrp,
                                                                   double t_end, int64_t max_steps, double* gscratch) {
    __shared__ WarpScratch scratch[kShipWarps];
    __shared__ ColTab T1, T2;  // T2: the velocity coefficients (AV, BV, EV) of an ERKNG method
    __shared__ double Tc[EE_RK_MAX_STAGES];
    __shared__ ColTab HC1[kShipWarps], HC2[KIND == 1 ? kShipWarps : 1];  // the tables times h (h^2: ERKNG positions) per warp
    // per warp: [STAGES][ngrp][3][32] body positions of the current attempt, then [ngrp][32][27] cached polynomials
    extern __shared__ __align__(16) double bp_all[];
    for (int e = threadIdx.x; e < EE_RK_MAX_STAGES * kColRows; e += blockDim.x) {
        const int sc = e / kColRows, r = e - sc * kColRows;
        double v1 = 0.0, v2 = 0.0;
        if (sc < STAGES && r > sc && r < STAGES + 2) {
            const RkDev& R = c_rk[method];
            v1 = r < STAGES ? R.a[r * (r - 1) / 2 + sc] : (r == STAGES ? R.b[sc] : R.e[sc]);
            v2 = r < STAGES ? R.a2[r * (r - 1) / 2 + sc] : (r == STAGES ? R.b2[sc] : R.e2[sc]);
        }
        T1.m[e] = v1;
        T2.m[e] = v2;
    }
    if (threadIdx.x < EE_RK_MAX_STAGES) Tc[threadIdx.x] = c_rk[method].c[threadIdx.x];
    const int kord_i = c_rk[method].kord;
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t ship = (int64_t)blockIdx.x * kShipWarps + warp;
    if (ship >= S.n) return;
    WarpScratch& ws = scratch[warp];
    // (systems too large for shared memory keep the two caches in a global scratch area instead: same code, slower)
    const size_t per_warp = (size_t)STAGES * ngrp * 96 + (size_t)ngrp * 32 * 27;
    double* bp = gscratch ? gscratch + (size_t)ship * per_warp : bp_all + (size_t)warp * per_warp;
    double* pc = bp + (size_t)STAGES * ngrp * 96;
    ColTab& H1 = HC1[warp];
    ColTab& H2 = HC2[KIND == 1 ? warp : 0];
    constexpr int kMaxGrp = NG ? NG : 32;  // NG = 1 is the compiled-in common case; the general kernel takes up to 1 024 bodies
    int64_t tag[kMaxGrp];
    int tag_nc[kMaxGrp];
    double mu_l[kMaxGrp];
#pragma unroll
    for (int g = 0; g < kMaxGrp; ++g) {
        tag[g] = -1;
        tag_nc[g] = 0;
        const int64_t b = (int64_t)g * 32 + lane;
        mu_l[g] = b < E.nb ? E.mu[b] : 0.0;
    }
    constexpr int kRows = STAGES + 2, kRowY = STAGES, kRowE = STAGES + 1;
    const int rl = lane / 6, cl = lane - 6 * rl;  // this lane's place in the row updates

    double time = S.time[ship], bound = S.bound[ship], next_h = S.next_h[ship];
    double y[6];
    for (int c = 0; c < 6; ++c) y[c] = S.state[6 * ship + c];
    uint32_t rk_i = S.rk_i[ship], n_att = S.n_att[ship];
    int32_t cur = S.cur_seg[ship], status = S.status[ship];
    int64_t nk = S.n_knots[ship];
    unsigned long long evals = S.rhs_evals[ship];
    const int64_t so = S.seg_off[ship];
    double last_t = S.knots[(ship * S.kcap + (nk - 1)) * 7];  // solution.end()
    int ntr = 0, nap = 0;
    double cur_t = 0.0;
    int cur_b = -1;
    if (ANA) {
        ntr = S.n_tr[ship];
        nap = S.n_ap[ship];
        if (ntr > 0) {
            cur_t = S.tr_time[ship * S.tr_cap + ntr - 1];
            cur_b = S.tr_body[ship * S.tr_cap + ntr - 1];
        }
    }
    double kl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // k[STAGES-1] of the last advance (the FSAL slope), kept across launches
    if (FSAL)
        for (int c = 0; c < 6; ++c) kl[c] = S.fsal_k[6 * ship + c];

    int64_t accepted = 0;
    while (status == EE_OK && accepted < max_steps && !(last_t >= t_end) && nk < S.kcap) {
        // a step can add at most one transition and one apsis per body: stop (the caller grows the lists) before overflow
        if (ANA && (ntr + E.nb + 1 > S.tr_cap || nap + E.nb + 2 > S.ap_cap)) break;
        // SpacecraftPropagator::step: a manoeuvre change re-initialises the integrator (spacecraft.rs:599-610)
        if (time >= S.seg_end[so + cur]) {
            cur += 1;
            bound = S.seg_end[so + cur];
            next_h = P.h_init;
            n_att = 0;
            rk_i = 0;
        }
        const bool burn = S.seg_burn[so + cur] != 0;
        const D3 bacc = {S.seg_acc[3 * (so + cur)], S.seg_acc[3 * (so + cur) + 1], S.seg_acc[3 * (so + cur) + 2]};
        const int bref = S.seg_ref[so + cur];
        // AdaptiveRungeKuttaIntegrator::advance; PreviousStep::store keeps (time, state, i, k[last] when FSAL)
        const double prev_t = time;
        double prev_y[6];
        for (int c = 0; c < 6; ++c) prev_y[c] = y[c];
        const uint32_t prev_i = rk_i;
        double prev_kl[6];
        for (int c = 0; c < 6; ++c) prev_kl[c] = kl[c];
        for (;;) {
            if (n_att > P.n_max) {
                status = EE_MAX_ITERATIONS_REACHED;
                break;
            }
            if (xadd(time, next_h) > bound) next_h = xsub(bound, time);
            const double h = next_h;
            if (time >= bound) {
                status = EE_BOUND_REACHED;
                break;
            }
            if (xadd(time, h) == time) {
                status = EE_STEP_SIZE_UNDERFLOW;
                break;
            }
            const double hh = xmul(h, h);
            const bool skip0 = FSAL && rk_i > 0;  // stage 0 takes the previous advance's last slope (k.swap(0, STAGES-1))
            const unsigned okmask = __reduce_and_sync(
                kFull, ship_body_positions<STAGES, NG>(E, bp, pc, tag, tag_nc, ngrp, lane, skip0 ? 1 : 0, time, h, Tc));  // warp-uniform
            // h * coefficient for the whole attempt, one product per lane and entry (the reference forms exactly these
            // products, `h * C::A[s][j]` etc., before multiplying a slope with them)
            for (int e = lane; e < STAGES * kColRows; e += 32) {
                H1.m[e] = xmul(KIND == 1 ? hh : h, T1.m[e]);
                if (KIND == 1) H2.m[e] = xmul(h, T2.m[e]);
            }
            // rows start from y (ERK) or from y + y' (h c_s) and y' (ERKNG); the error row from zero
            for (int c = 0; c < 6; ++c) ws.k6[c] = y[c];  // through shared memory: a select of y[c] by lane would branch six ways
            __syncwarp();
            for (int e = lane; e < kRows * 6; e += 32) {
                const int r = e / 6, c = e - 6 * r;
                const double yc = ws.k6[c];
                double v = yc;
                if (KIND == 1 && c < 3 && r != kRowE) v = xadd(yc, xmul(ws.k6[3 + c], r == kRowY ? h : xmul(h, Tc[r])));
                ws.P[r][c] = r == kRowE ? 0.0 : v;
            }
            __syncwarp();
            bool ok = true;
            double k[6];
            for (int s = 0; s < STAGES; ++s) {
                if (s == 0 && skip0) {
                    for (int c = 0; c < 6; ++c) k[c] = kl[c];
                } else {
                    const double ti = xadd(time, xmul(h, Tc[s]));
                    double yi[6];
                    for (int c = 0; c < 6; ++c) yi[c] = ws.P[s][c];
                    evals += 1;
                    ok = (okmask >> s) & 1u;
                    if (!ok) break;
                    const D3 ctx = ship_context_acceleration<NG>(E, ws, bp, ngrp, mu_l, s, lane, d3(yi[0], yi[1], yi[2]));
                    D3 ma;
                    ok = ship_manoeuvre_acceleration(E, ti, yi, burn, bacc, bref, &ma);
                    if (!ok) break;
                    const D3 acc = xadd3(ctx, ma);
                    k[0] = yi[3];
                    k[1] = yi[4];
                    k[2] = yi[5];
                    k[3] = acc.x;
                    k[4] = acc.y;
                    k[5] = acc.z;
                }
                // k_s's term goes into every row that still needs it: the later stages, the new state, the error.  Lane
                // (rl, cl) owns component cl of rows s + 1 + rl, s + 6 + rl, ...; lanes 30 and 31 idle.
                for (int c = 0; c < 6; ++c) ws.k6[c] = k[c];  // every lane holds the same six values
                __syncwarp();
                if (rl < 5) {
                    const bool second = KIND == 1 && cl >= 3;  // ERKNG velocities: h AV; positions: h^2 AP, both from the slope
                    const double kv = ws.k6[KIND == 0 ? cl : (cl < 3 ? cl + 3 : cl)];
                    const double* hm = (second ? H2.m : H1.m) + s * kColRows;
                    double hc[4], pv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {  // rows s+1+rl, +5, +10, +15: loads first, then the independent updates
                        const int r = s + 1 + rl + 5 * i;  // (no clamped dummy read: it would race with the row's owner)
                        hc[i] = r < kRows ? hm[r] : 0.0;
                        pv[i] = r < kRows ? ws.P[r][cl] : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = s + 1 + rl + 5 * i;
                        if (r < kRows) ws.P[r][cl] = xadd(pv[i], xmul(kv, hc[i]));
                    }
                }
                __syncwarp();
            }
            if (!ok) {
                status = EE_EVAL_FAILED;
                break;
            }
            for (int c = 0; c < 6; ++c) kl[c] = k[c];
            double er[6];
            for (int c = 0; c < 6; ++c) {
                y[c] = ws.P[kRowY][c];
                er[c] = ws.P[kRowE][c];
            }
            __syncwarp();  // the rows are re-initialised by the next attempt
            time = xadd(time, h);
            rk_i += 1;
            n_att += 1;
            // AbsTol::err_over_tol
            const double ea = fmax(fabs(xdiv(er[0], P.tol_pos)), fmax(fabs(xdiv(er[1], P.tol_pos)), fabs(xdiv(er[2], P.tol_pos))));
            const double eb = fmax(fabs(xdiv(er[3], P.tol_vel)), fmax(fabs(xdiv(er[4], P.tol_vel)), fabs(xdiv(er[5], P.tol_vel))));
            const double err = fmax(ea, eb);
            // IController::step with order = LOWER_ORDER
            const double kord = (double)kord_i;
            const double pexp = -xdiv(1.0, kord);
            // err.powf(-1/k): glibc's pow by default (= Rust's powf on Linux, the reference as built), see ee_pow_glibc.h
            const double pw = P.pow_mode == EE_POW_GLIBC ? pow_glibc(err, pexp) : pow_portable(err, pexp);
            const double mfac = xmul(P.fac, pw);
            const double cl = mfac < P.fac_min ? P.fac_min : (mfac > P.fac_max ? P.fac_max : mfac);
            const double nh = xmul(next_h, cl);
            next_h = nh > P.h_max ? P.h_max : nh;
            if (err <= 1.0) break;
            time = prev_t;  // PreviousStep::restore
            for (int c = 0; c < 6; ++c) y[c] = prev_y[c];
            rk_i = prev_i;
            if (FSAL)
                for (int c = 0; c < 6; ++c) kl[c] = prev_kl[c];
        }
        if (status != EE_OK) break;
        // CubicHermiteSplineSolout::solout: one knot per accepted step
        if (lane == 0) {
            double* kn = S.knots + (ship * S.kcap + nk) * 7;
            kn[0] = time;
            for (int c = 0; c < 6; ++c) kn[1 + c] = y[c];
        }
        nk += 1;
        last_t = time;
        accepted += 1;
        if (ANA) ana_step(S, E, pc, tag, tag_nc, ship, lane, prev_t, prev_y, time, y, ntr, nap, cur_t, cur_b);
    }
    if (lane == 0) {
        S.time[ship] = time;
        S.bound[ship] = bound;
        S.next_h[ship] = next_h;
        for (int c = 0; c < 6; ++c) S.state[6 * ship + c] = y[c];
        S.rk_i[ship] = rk_i;
        S.n_att[ship] = n_att;
        S.cur_seg[ship] = cur;
        S.status[ship] = status;
        S.n_knots[ship] = nk;
        S.rhs_evals[ship] = evals;
        if (ANA) {
            S.n_tr[ship] = ntr;
            S.n_ap[ship] = nap;
        }
        if (FSAL)
            for (int c = 0; c < 6; ++c) S.fsal_k[6 * ship + c] = kl[c];
    }
}

// take_solution: the new CubicHermiteSpline starts at the current (time, position, velocity)
__global__ void k_ships_reset_knots(ShipsView S) {
    const int64_t ship = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ship >= S.n) return;
    double* kn = S.knots + ship * S.kcap * 7;
    kn[0] = S.time[ship];
    for (int c = 0; c < 6; ++c) kn[1 + c] = S.state[6 * ship + c];
    S.n_knots[ship] = 1;
}

// UniformSpline::position / state_vector, one thread per (time, body)
__global__ void k_ephem_evaluate(EphemView E, int64_t nt, const double* __restrict__ times, double* __restrict__ pos,
                                 double* __restrict__ vel, int32_t* __restrict__ okf) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nt * E.nb) return;
    const int64_t ti = idx / E.nb, b = idx % E.nb;
    D3 p = {0.0, 0.0, 0.0}, v = {0.0, 0.0, 0.0};
    bool ok;
    if (vel)
        ok = spline_state_vector(E, b, times[ti], &p, &v);
    else
        ok = spline_position(E, b, times[ti], &p);
    okf[idx] = ok ? 1 : 0;
    pos[3 * idx] = p.x;
    pos[3 * idx + 1] = p.y;
    pos[3 * idx + 2] = p.z;
    if (vel) {
        vel[3 * idx] = v.x;
        vel[3 * idx + 1] = v.y;
        vel[3 * idx + 2] = v.z;
    }
}

// RelativeTrajectory::state_vector (ephemeris/src/trajectory.rs:315-335) batched over times -- what the plot sampler
// evaluates for every candidate point (ephemeris_explorer/src/ui/world/plot.rs:326-334).  The trajectory is either body
// `body` of the ephemeris (UniformSpline::state_vector) or, when `knots` is given, a CubicHermiteSpline of `nk` knots
// (trajectory.rs:779-795: a time equal to a knot returns the knot, otherwise the cubic-Hermite segment around it); the
// reference is a body of the ephemeris or none (-1).  One thread per time.
__global__ void k_eval_relative(EphemView E, const double* __restrict__ knots, int64_t nk, int body, int reference, int64_t nt,
                                const double* __restrict__ times, double* __restrict__ pos, double* __restrict__ vel,
                                int32_t* __restrict__ okf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const double t = times[i];
    D3 p = {0.0, 0.0, 0.0}, v = {0.0, 0.0, 0.0};
    bool ok = true;
    D3 rp = {0.0, 0.0, 0.0}, rv = {0.0, 0.0, 0.0};
    if (reference >= 0) ok = spline_state_vector(E, reference, t, &rp, &rv);  // evaluated first, like the reference does
    if (ok) {
        if (knots) {
            // binary_search_by(|(k, _)| k.cmp(&at)): first index whose time is >= t
            int64_t lo = 0, hi = nk;
            while (lo < hi) {
                const int64_t mid = lo + (hi - lo) / 2;
                if (knots[7 * mid] < t)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo < nk && knots[7 * lo] == t) {
                p = d3(knots[7 * lo + 1], knots[7 * lo + 2], knots[7 * lo + 3]);
                v = d3(knots[7 * lo + 4], knots[7 * lo + 5], knots[7 * lo + 6]);
            } else if (lo == 0 || lo >= nk) {
                ok = false;  // i.checked_sub(1)? / hermite3(i - 1)?
            } else {
                const double* k0 = knots + 7 * (lo - 1);
                const double* k1 = knots + 7 * lo;
                const HermiteD H = hermite_make(k0[0], k0 + 1, k1[0], k1 + 1);
                p = hermite_eval(H, t);
                v = hermite_deriv(H, t);
            }
        } else {
            ok = spline_state_vector(E, body, t, &p, &v);
        }
    }
    if (ok) {
        p = xsub3(p, rp);
        v = xsub3(v, rv);
    } else {
        p = v = d3(0.0, 0.0, 0.0);
    }
    okf[i] = ok ? 1 : 0;
    pos[3 * i] = p.x;
    pos[3 * i + 1] = p.y;
    pos[3 * i + 2] = p.z;
    vel[3 * i] = v.x;
    vel[3 * i + 1] = v.y;
    vel[3 * i + 2] = v.z;
}

static EphemView view_of(const Ephem& e) {
    return EphemView{e.nb, e.d_mu.p, e.d_start.p, e.d_interval.p, e.d_npoly.p, e.d_first.p, e.coef.p, e.ncoef.p};
}

void Ephem::evaluate(int64_t nt, const double* times, double* pos, double* vel, int32_t* ok) {
    EE_REQUIRE(nt >= 0 && times && pos && ok, "bad arguments");
    if (nt == 0) return;
    EE_CUDA(cudaSetDevice(device));
    const int64_t tot = nt * nb;
    DBuf<double> d_t((size_t)nt), d_p((size_t)tot * 3), d_v;
    DBuf<int32_t> d_ok((size_t)tot);
    if (vel) d_v.alloc((size_t)tot * 3);
    EE_CUDA(cudaMemcpy(d_t.p, times, (size_t)nt * 8, cudaMemcpyHostToDevice));
    const int B = 128;
    k_ephem_evaluate<<<(unsigned)((tot + B - 1) / B), B>>>(view_of(*this), nt, d_t.p, d_p.p, vel ? d_v.p : nullptr, d_ok.p);
    EE_CUDA(cudaGetLastError());
    count_launch();
    EE_CUDA(cudaMemcpy(pos, d_p.p, d_p.bytes(), cudaMemcpyDeviceToHost));
    if (vel) EE_CUDA(cudaMemcpy(vel, d_v.p, d_v.bytes(), cudaMemcpyDeviceToHost));
    EE_CUDA(cudaMemcpy(ok, d_ok.p, d_ok.bytes(), cudaMemcpyDeviceToHost));
}

static void eval_relative_launch(const Ephem& e, const double* d_knots, int64_t nk, int body, int reference, int64_t nt,
                                 const double* times, double* pos, double* vel, int32_t* ok) {
    EE_REQUIRE(nt >= 0 && times && pos && vel && ok, "bad arguments");
    EE_REQUIRE(reference >= -1 && reference < e.nb, "reference body out of range");
    EE_REQUIRE(d_knots || (body >= 0 && body < e.nb), "body out of range");
    if (nt == 0) return;
    DBuf<double> d_t((size_t)nt), d_p((size_t)nt * 3), d_v((size_t)nt * 3);
    DBuf<int32_t> d_ok((size_t)nt);
    EE_CUDA(cudaMemcpy(d_t.p, times, (size_t)nt * 8, cudaMemcpyHostToDevice));
    const int B = 128;
    k_eval_relative<<<(unsigned)((nt + B - 1) / B), B>>>(view_of(e), d_knots, nk, body, reference, nt, d_t.p, d_p.p, d_v.p, d_ok.p);
    EE_CUDA(cudaGetLastError());
    count_launch();
    EE_CUDA(cudaMemcpy(pos, d_p.p, d_p.bytes(), cudaMemcpyDeviceToHost));
    EE_CUDA(cudaMemcpy(vel, d_v.p, d_v.bytes(), cudaMemcpyDeviceToHost));
    EE_CUDA(cudaMemcpy(ok, d_ok.p, d_ok.bytes(), cudaMemcpyDeviceToHost));
}

void Ephem::evaluate_relative(int body, int reference, int64_t nt, const double* times, double* pos, double* vel, int32_t* ok) {
    EE_CUDA(cudaSetDevice(device));
    eval_relative_launch(*this, nullptr, 0, body, reference, nt, times, pos, vel, ok);
}

void Ships::evaluate_relative(int64_t ship, int reference, int64_t nt, const double* times, double* pos, double* vel, int32_t* ok) {
    EE_REQUIRE(ship >= 0 && ship < n, "ship index out of range");
    EE_CUDA(cudaSetDevice(ephem->device));
    EE_CUDA(cudaStreamSynchronize(stream));
    int64_t nk = 0;
    EE_CUDA(cudaMemcpy(&nk, d_nknots.p + ship, 8, cudaMemcpyDeviceToHost));
    eval_relative_launch(*ephem, knots.p + (size_t)ship * kcap * 7, nk, -1, reference, nt, times, pos, vel, ok);
}

// ---------------------------------------------------------------------------------------------------------
static const double kEpochMin = -std::numeric_limits<double>::max();  // ftime Duration::MIN / MAX
static const double kEpochMax = std::numeric_limits<double>::max();

Ships::Ships(Ephem* eph, int64_t n_, const double* t0, const double* states, const ee_adaptive_params* p,
             const int64_t* burn_off, const double* bstart, const double* bend, const double* bacc, const int32_t* bref)
    : ephem(eph), n(n_) {
    EE_REQUIRE(eph && n >= 1 && t0 && states && p, "bad arguments");
    EE_REQUIRE(p->pow_mode == EE_POW_GLIBC || p->pow_mode == EE_POW_CORRECTLY_ROUNDED, "unknown pow_mode");
    EE_REQUIRE(p->method < EE_RK_METHODS, "unknown adaptive method id (see EE_SHIP_* in ee_b200.h)");
    params = *p;
    EE_CUDA(cudaSetDevice(eph->device));
    EE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    EE_CUDA(cudaEventCreate(&ev0));
    EE_CUDA(cudaEventCreate(&ev1));
    static std::atomic<uint64_t> consts_loaded{0};  // constant memory is per device
    const uint64_t dbit = 1ull << (eph->device & 63);
    if (!(consts_loaded.load(std::memory_order_acquire) & dbit)) {
        std::vector<RkDev> tabs(EE_RK_METHODS);
        for (int m = 0; m < EE_RK_METHODS; ++m) {
            const EeRkTableau& t = EE_RK_TABLE[m];
            RkDev& d = tabs[(size_t)m];
            std::memset(&d, 0, sizeof(d));
            d.stages = t.stages;
            d.fsal = t.fsal;
            d.kord = t.kord;
            d.kind = t.kind;
            const int tri = t.stages * (t.stages - 1) / 2;
            std::copy(t.a, t.a + tri, d.a);
            std::copy(t.b, t.b + t.stages, d.b);
            std::copy(t.c, t.c + t.stages, d.c);
            std::copy(t.e, t.e + t.stages, d.e);
            if (t.kind) {
                std::copy(t.a2, t.a2 + tri, d.a2);
                std::copy(t.b2, t.b2 + t.stages, d.b2);
                std::copy(t.e2, t.e2 + t.stages, d.e2);
            }
        }
        EE_CUDA(cudaMemcpyToSymbol(c_rk, tabs.data(), sizeof(RkDev) * EE_RK_METHODS));
        consts_loaded.fetch_or(dbit, std::memory_order_release);
    }
    // Timeline::new per ship (spacecraft.rs:131-160): sort burns by start, interleave coasts from Epoch::MIN to MAX
    std::vector<int64_t> seg_off((size_t)n + 1, 0);
    std::vector<int32_t> seg_burn, seg_ref, cur((size_t)n);
    std::vector<double> seg_end, seg_acc, bound((size_t)n);
    for (int64_t s = 0; s < n; ++s) {
        struct B {
            double st, en;
            double a[3];
            int32_t ref;
        };
        std::vector<B> burns;
        if (burn_off)
            for (int64_t k = burn_off[s]; k < burn_off[s + 1]; ++k) {
                EE_REQUIRE(bref[k] < eph->nb, "burn reference body out of range");
                burns.push_back({bstart[k], bend[k], {bacc[3 * k], bacc[3 * k + 1], bacc[3 * k + 2]}, bref[k]});
            }
        std::stable_sort(burns.begin(), burns.end(), [](const B& a, const B& b) { return a.st < b.st; });
        double cursor = kEpochMin;
        std::vector<double> ends;
        auto push = [&](bool burn, double end, const double* a, int32_t ref) {
            seg_burn.push_back(burn ? 1 : 0);
            seg_end.push_back(end);
            for (int c = 0; c < 3; ++c) seg_acc.push_back(a ? a[c] : 0.0);
            seg_ref.push_back(ref);
            ends.push_back(end);
        };
        for (const B& b : burns) {
            if (b.st > cursor) push(false, b.st, nullptr, -1);
            cursor = b.en;
            push(true, b.en, b.a, b.ref);
        }
        if (cursor < kEpochMax) push(false, kEpochMax, nullptr, -1);
        seg_off[(size_t)s + 1] = (int64_t)seg_end.size();
        // segment_idx_at: partition_point(seg.end() <= time)
        size_t i = 0;
        while (i < ends.size() && ends[i] <= t0[s]) ++i;
        EE_REQUIRE(i < ends.size(), "initial time beyond the last segment");
        cur[(size_t)s] = (int32_t)i;
        bound[(size_t)s] = ends[i];
    }
    auto up = [&](auto& dbuf, const auto& host) {
        dbuf.alloc(host.size());
        if (!host.empty())
            EE_CUDA(cudaMemcpy(dbuf.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice));
    };
    up(d_seg_off, seg_off);
    up(d_seg_burn, seg_burn);
    up(d_seg_end, seg_end);
    up(d_seg_acc, seg_acc);
    up(d_seg_ref, seg_ref);
    up(d_cur, cur);
    up(d_bound, bound);
    std::vector<double> time(t0, t0 + n), st(states, states + 6 * n), nh((size_t)n, p->h_init);
    up(d_time, time);
    up(d_state, st);
    up(d_next_h, nh);
    std::vector<uint32_t> zu((size_t)n, 0);
    std::vector<int32_t> zs((size_t)n, 0);
    std::vector<int64_t> one((size_t)n, 1);
    std::vector<unsigned long long> zl((size_t)n, 0);
    up(d_rk_i, zu);
    up(d_natt, zu);
    up(d_status, zs);
    up(d_nknots, one);
    up(d_evals, zl);
    std::vector<double> zk((size_t)6 * n, 0.0);
    up(d_fsal_k, zk);
    // CubicHermiteSplineSolout::new_solution: first knot = (t0, position, velocity)
    kcap = 64;
    knots.alloc((size_t)n * kcap * 7);
    std::vector<double> k0((size_t)n * kcap * 7, 0.0);
    for (int64_t s = 0; s < n; ++s) {
        k0[(size_t)(s * kcap * 7)] = t0[s];
        for (int c = 0; c < 6; ++c) k0[(size_t)(s * kcap * 7 + 1 + c)] = states[6 * s + c];
    }
    EE_CUDA(cudaMemcpy(knots.p, k0.data(), k0.size() * 8, cudaMemcpyHostToDevice));
    max_held = 1;
}

Ships::~Ships() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
}

static ShipsView ships_view(Ships& s) {
    ShipsView v{};
    v.n = s.n;
    v.time = s.d_time.p;
    v.bound = s.d_bound.p;
    v.state = s.d_state.p;
    v.next_h = s.d_next_h.p;
    v.rk_i = s.d_rk_i.p;
    v.n_att = s.d_natt.p;
    v.cur_seg = s.d_cur.p;
    v.status = s.d_status.p;
    v.n_knots = s.d_nknots.p;
    v.rhs_evals = s.d_evals.p;
    v.seg_off = s.d_seg_off.p;
    v.seg_burn = s.d_seg_burn.p;
    v.seg_end = s.d_seg_end.p;
    v.seg_acc = s.d_seg_acc.p;
    v.seg_ref = s.d_seg_ref.p;
    v.knots = s.knots.p;
    v.kcap = s.kcap;
    v.fsal_k = s.d_fsal_k.p;
    v.analytics = s.analytics ? 1 : 0;
    v.soi_r = s.d_soi_r.p;
    v.tr_cap = s.tr_cap;
    v.ap_cap = s.ap_cap;
    v.tr_time = s.d_tr_time.p;
    v.tr_body = s.d_tr_body.p;
    v.n_tr = s.d_ntr.p;
    v.ap_time = s.d_ap_time.p;
    v.ap_dist = s.d_ap_dist.p;
    v.ap_body = s.d_ap_body.p;
    v.ap_kind = s.d_ap_kind.p;
    v.n_ap = s.d_nap.p;
    return v;
}

// SpacecraftSolout::new_solution (dynamics/spacecraft.rs:518-533): transitions = [(now, soi_at(now, position))], no apsides
__global__ void k_ships_new_analytics(ShipsView S, EphemView E) {
    const int64_t ship = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ship >= S.n) return;
    const double t = S.time[ship];
    const D3 pos = d3(S.state[6 * ship], S.state[6 * ship + 1], S.state[6 * ship + 2]);
    const int cur = ana_soi_at_except(E, S.soi_r, t, pos, -1);
    S.n_ap[ship] = 0;
    if (cur >= 0) {
        S.tr_time[ship * S.tr_cap] = t;
        S.tr_body[ship * S.tr_cap] = cur;
        S.n_tr[ship] = 1;
    } else {
        S.n_tr[ship] = 0;
    }
}

void Ships::enable_analytics(const double* soi_radius) {
    EE_REQUIRE(soi_radius, "null soi_radius");
    EE_CUDA(cudaSetDevice(ephem->device));
    EE_CUDA(cudaStreamSynchronize(stream));
    d_soi_r.alloc((size_t)ephem->nb);
    EE_CUDA(cudaMemcpy(d_soi_r.p, soi_radius, (size_t)ephem->nb * 8, cudaMemcpyHostToDevice));
    tr_cap = std::max<int64_t>(64, 4 * (ephem->nb + 2));
    ap_cap = std::max<int64_t>(256, 4 * (ephem->nb + 2));
    d_tr_time.alloc((size_t)n * tr_cap);
    d_tr_body.alloc((size_t)n * tr_cap);
    d_ap_time.alloc((size_t)n * ap_cap);
    d_ap_dist.alloc((size_t)n * ap_cap);
    d_ap_body.alloc((size_t)n * ap_cap);
    d_ap_kind.alloc((size_t)n * ap_cap);
    d_ntr.alloc((size_t)n);
    d_nap.alloc((size_t)n);
    analytics = true;
    reset_analytics();
}

void Ships::reset_analytics() {
    k_ships_new_analytics<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(ships_view(*this), view_of(*ephem));
    EE_CUDA(cudaGetLastError());
    count_launch();
    EE_CUDA(cudaStreamSynchronize(stream));
    max_tr = 1;
    max_ap = 0;
}

template <class T>
static void grow_rows(DBuf<T>& buf, int64_t n, int64_t old_cap, int64_t new_cap, int64_t used, cudaStream_t stream) {
    DBuf<T> nb((size_t)n * new_cap);
    if (used > 0)
        EE_CUDA(cudaMemcpy2DAsync(nb.p, (size_t)new_cap * sizeof(T), buf.p, (size_t)old_cap * sizeof(T), (size_t)used * sizeof(T),
                                  (size_t)n, cudaMemcpyDeviceToDevice, stream));
    EE_CUDA(cudaStreamSynchronize(stream));
    buf = std::move(nb);
}

// Lists that are more than half full are doubled before a launch; the kernel stops a ship whose lists could overflow in
// one more step (status stays OK, the caller's step_to loop comes back).
void Ships::ensure_analytics_capacity() {
    if (!analytics) return;
    const int64_t margin = ephem->nb + 2;
    if (2 * (max_tr + margin) > tr_cap) {
        int64_t nc = tr_cap;
        while (2 * (max_tr + margin) > nc) nc *= 2;
        grow_rows(d_tr_time, n, tr_cap, nc, max_tr, stream);
        grow_rows(d_tr_body, n, tr_cap, nc, max_tr, stream);
        tr_cap = nc;
    }
    if (2 * (max_ap + margin) > ap_cap) {
        int64_t nc = ap_cap;
        while (2 * (max_ap + margin) > nc) nc *= 2;
        grow_rows(d_ap_time, n, ap_cap, nc, max_ap, stream);
        grow_rows(d_ap_dist, n, ap_cap, nc, max_ap, stream);
        grow_rows(d_ap_body, n, ap_cap, nc, max_ap, stream);
        grow_rows(d_ap_kind, n, ap_cap, nc, max_ap, stream);
        ap_cap = nc;
    }
}

void Ships::analytics_counts(int32_t* n_tr, int32_t* n_ap) {
    EE_REQUIRE(analytics, "analytics are not enabled on this handle (ee_ships_enable_analytics)");
    EE_CUDA(cudaSetDevice(ephem->device));
    EE_CUDA(cudaStreamSynchronize(stream));
    if (n_tr) EE_CUDA(cudaMemcpy(n_tr, d_ntr.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (n_ap) EE_CUDA(cudaMemcpy(n_ap, d_nap.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
}

template <class T>
static void gather_rows(const DBuf<T>& buf, int64_t n, int64_t cap, const std::vector<int32_t>& cnt, const int64_t* off, T* out) {
    if (!out) return;
    int64_t mx = 0;
    for (int32_t c : cnt) mx = std::max<int64_t>(mx, c);
    if (mx == 0) return;
    std::vector<T> host((size_t)n * mx);
    EE_CUDA(cudaMemcpy2D(host.data(), (size_t)mx * sizeof(T), buf.p, (size_t)cap * sizeof(T), (size_t)mx * sizeof(T), (size_t)n,
                         cudaMemcpyDeviceToHost));
    for (int64_t s = 0; s < n; ++s)
        std::copy(host.begin() + (size_t)(s * mx), host.begin() + (size_t)(s * mx + cnt[(size_t)s]), out + off[s]);
}

void Ships::read_analytics(const int64_t* tr_off, double* tr_time, int32_t* tr_body, const int64_t* ap_off, double* ap_time,
                           double* ap_distance, int32_t* ap_body, int32_t* ap_kind) {
    EE_REQUIRE(analytics, "analytics are not enabled on this handle (ee_ships_enable_analytics)");
    EE_REQUIRE(tr_off && ap_off, "null offsets");
    std::vector<int32_t> ntr((size_t)n), nap((size_t)n);
    analytics_counts(ntr.data(), nap.data());
    for (int64_t s = 0; s < n; ++s) {
        EE_REQUIRE(tr_off[s + 1] - tr_off[s] == ntr[(size_t)s], "transition offsets do not match ee_ships_analytics_counts");
        EE_REQUIRE(ap_off[s + 1] - ap_off[s] == nap[(size_t)s], "apsis offsets do not match ee_ships_analytics_counts");
    }
    gather_rows(d_tr_time, n, tr_cap, ntr, tr_off, tr_time);
    gather_rows(d_tr_body, n, tr_cap, ntr, tr_off, tr_body);
    gather_rows(d_ap_time, n, ap_cap, nap, ap_off, ap_time);
    gather_rows(d_ap_dist, n, ap_cap, nap, ap_off, ap_distance);
    gather_rows(d_ap_body, n, ap_cap, nap, ap_off, ap_body);
    gather_rows(d_ap_kind, n, ap_cap, nap, ap_off, ap_kind);
}

void Ships::ensure_capacity(int64_t extra) {
    const int64_t need = max_held + extra;
    if (need <= kcap) return;
    int64_t ncap = kcap;
    while (ncap < need) ncap *= 2;
    DBuf<double> nk((size_t)n * ncap * 7);
    EE_CUDA(cudaMemcpy2DAsync(nk.p, (size_t)ncap * 56, knots.p, (size_t)kcap * 56, (size_t)max_held * 56, (size_t)n,
                              cudaMemcpyDeviceToDevice, stream));
    EE_CUDA(cudaStreamSynchronize(stream));
    knots = std::move(nk);
    kcap = ncap;
}

void Ships::step_to(double t_end, int64_t max_steps) {
    EE_REQUIRE(max_steps >= 0, "negative max_steps");
    EE_CUDA(cudaSetDevice(ephem->device));
    ensure_capacity(max_steps);
    ShipParams P{params.h_init, params.h_max, params.tol_position, params.tol_velocity,
                 params.fac_min, params.fac_max, params.fac, params.n_max, params.pow_mode};
    ensure_analytics_capacity();
    const unsigned grid = (unsigned)((n + kShipWarps - 1) / kShipWarps);
    const ShipsView sv = ships_view(*this);
    const EphemView evw = view_of(*ephem);
    const int method = (int)params.method;
    EE_CUDA(cudaEventRecord(ev0, stream));
    const int ngrp = (int)((ephem->nb + 31) / 32);
    auto launch = [&](auto kernel, int stages) {
        const size_t per_warp = (size_t)stages * ngrp * 96 + (size_t)ngrp * 32 * 27;
        size_t smem = (size_t)kShipWarps * per_warp * sizeof(double);
        if (ngrp > 32) throw Error(EE_ERR_UNSUPPORTED, "the ship kernel takes ephemerides of up to 1 024 bodies");
        double* scr = nullptr;
        if (smem > 200 * 1024) {  // beyond shared memory (about 100 bodies): the position and polynomial caches move to global memory
            if (scratch.n < (size_t)n * per_warp) scratch.alloc((size_t)n * per_warp);
            scr = scratch.p;
            smem = 0;
        }
        EE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, kShipWarps * 32, smem, stream>>>(sv, evw, P, method, ngrp, t_end, max_steps, scr);
    };
    // one instantiation per (stages, FSAL, kind) x (plain | SpacecraftSolout analytics): the plain kernels make no call and
    // keep everything in registers
#define EE_SHIP_LAUNCH(ST, FS, KD)                                      \
    do {                                                                \
        if (analytics && ngrp == 1)                                     \
            launch(k_ships_step_to<ST, FS, KD, true, 1>, ST);           \
        else if (analytics)                                             \
            launch(k_ships_step_to<ST, FS, KD, true, 0>, ST);           \
        else if (ngrp == 1)                                             \
            launch(k_ships_step_to<ST, FS, KD, false, 1>, ST);          \
        else                                                            \
            launch(k_ships_step_to<ST, FS, KD, false, 0>, ST);          \
    } while (0)
    switch (method) {
        case EE_SHIP_VERNER87:
        case EE_SHIP_DORMAND_PRINCE87: EE_SHIP_LAUNCH(13, false, 0); break;
        case EE_SHIP_CASH_KARP45:
        case EE_SHIP_FEHLBERG45: EE_SHIP_LAUNCH(6, false, 0); break;
        case EE_SHIP_DORMAND_PRINCE54: EE_SHIP_LAUNCH(7, true, 0); break;
        case EE_SHIP_TSITOURAS75: EE_SHIP_LAUNCH(9, false, 0); break;
        case EE_SHIP_VERNER98: EE_SHIP_LAUNCH(16, false, 0); break;
        case EE_SHIP_FINE45: EE_SHIP_LAUNCH(7, true, 1); break;
        default: throw Error(EE_ERR_INVALID, "unknown adaptive method id");
    }
#undef EE_SHIP_LAUNCH
    EE_CUDA(cudaGetLastError());
    EE_CUDA(cudaEventRecord(ev1, stream));
    count_launch();
    EE_CUDA(cudaStreamSynchronize(stream));
    float ms = 0.f;
    EE_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    last_ms = ms;
    // knots actually held (not the max_steps upper bound) decide how much the next call has to grow the buffer
    std::vector<int64_t> nk((size_t)n);
    EE_CUDA(cudaMemcpy(nk.data(), d_nknots.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    max_held = 1;
    for (int64_t v : nk) max_held = std::max(max_held, v);
    if (analytics) {
        std::vector<int32_t> a((size_t)n), b((size_t)n);
        EE_CUDA(cudaMemcpy(a.data(), d_ntr.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        EE_CUDA(cudaMemcpy(b.data(), d_nap.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        max_tr = 1;
        max_ap = 0;
        for (int64_t s = 0; s < n; ++s) {
            max_tr = std::max<int64_t>(max_tr, a[(size_t)s]);
            max_ap = std::max<int64_t>(max_ap, b[(size_t)s]);
        }
    }
}

void Ships::info(int32_t* status, double* time, int64_t* n_knots, uint32_t* n_attempts, uint64_t* rhs_evals) {
    EE_CUDA(cudaSetDevice(ephem->device));
    EE_CUDA(cudaStreamSynchronize(stream));
    if (status) EE_CUDA(cudaMemcpy(status, d_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (time) EE_CUDA(cudaMemcpy(time, d_time.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (n_knots) EE_CUDA(cudaMemcpy(n_knots, d_nknots.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (n_attempts) EE_CUDA(cudaMemcpy(n_attempts, d_natt.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (rhs_evals) EE_CUDA(cudaMemcpy(rhs_evals, d_evals.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
}

void Ships::take_knots(const int64_t* offsets, double* out) {
    EE_CUDA(cudaSetDevice(ephem->device));
    EE_CUDA(cudaStreamSynchronize(stream));
    std::vector<int64_t> nk((size_t)n);
    EE_CUDA(cudaMemcpy(nk.data(), d_nknots.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    int64_t mx = 0;
    for (int64_t s = 0; s < n; ++s) {
        EE_REQUIRE(offsets[s + 1] - offsets[s] == nk[(size_t)s], "knot_offsets do not match ee_ships_info n_knots");
        mx = std::max(mx, nk[(size_t)s]);
    }
    std::vector<double> host((size_t)n * mx * 7);
    EE_CUDA(cudaMemcpy2D(host.data(), (size_t)mx * 56, knots.p, (size_t)kcap * 56, (size_t)mx * 56, (size_t)n,
                         cudaMemcpyDeviceToHost));
    for (int64_t s = 0; s < n; ++s)
        std::copy(host.begin() + (size_t)(s * mx * 7), host.begin() + (size_t)(s * mx * 7 + nk[(size_t)s] * 7),
                  out + offsets[s] * 7);
    k_ships_reset_knots<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(ships_view(*this));
    EE_CUDA(cudaGetLastError());
    count_launch();
    EE_CUDA(cudaStreamSynchronize(stream));
    max_held = 1;
    if (analytics) reset_analytics();  // take_solution replaces the whole SpacecraftSolution (trajectory, transitions, apsides)
}

}  // namespace ee
