// ee_ships.cu -- massless-ship propagator: adaptive Verner 8(7) over the device-resident spline ephemeris.
//
// Reference functions restated here (file:line under the reference tree):
//   SpacecraftPropagator::{new, step}            ephemeris/src/propagators/spacecraft.rs:453-477, :599-615   (a17)
//   Timeline::new / segment_idx_at               ephemeris/src/propagators/spacecraft.rs:131-170
//   SpacecraftModel::eval (first-order form)     ephemeris/src/propagators/spacecraft.rs:283-309             (a13)
//   Bodies::acceleration / acceleration_at       ephemeris_explorer/src/dynamics/spacecraft.rs:71-74, :218-229 (a14)
//   ReferenceFrame::transform / TNB              ephemeris_explorer/src/dynamics/spacecraft.rs:240-293
//   ERK<Verner87,[_;13]>::{advance, error}       integration/src/runge_kutta/explicit.rs:73-132              (a15)
//   AdaptiveRungeKuttaIntegrator::advance        integration/src/runge_kutta/mod.rs:396-440                  (a16)
//   IController::step                            integration/src/runge_kutta/mod.rs:225-243
//   AbsTol::err_over_tol                         ephemeris_explorer/src/dynamics/spacecraft.rs:615-625
//   CubicHermiteSplineSolout                     ephemeris/src/propagators/spacecraft.rs:645-695             (a18)
//
// Mapping: one WARP per ship.  Lane b evaluates body b's spline and its pull on the ship, so the 32 spline lookups
// of one right-hand side run side by side; the 32 contributions are then added in body order by one lane per
// component (the reference sums in construction order), which keeps the result bit-identical to the scalar path.
// Everything else (stage combinations, error norm, controller) is warp-uniform and kept in registers/shared memory.
#include <algorithm>
#include <cmath>
#include <limits>

#include "ee_coeffs.h"
#include "ee_engine.h"
#include "ee_pow.cuh"
#include "ee_pow_glibc.h"
#include "ee_ships.h"
#include "ee_spline.cuh"

namespace ee {

__constant__ double c_v87_a[169];
__constant__ double c_v87_b[13];
__constant__ double c_v87_c[13];
__constant__ double c_v87_e[13];

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
};

constexpr int kShipWarps = 4;
constexpr unsigned kFull = 0xffffffffu;

struct WarpScratch {
    double k[13][6];
    double a[32][3];
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

// SpacecraftModel::eval: dy.velocity = context + manoeuvre; dy.position = y.velocity.  Warp-collective.
__device__ bool ship_rhs(const EphemView& E, WarpScratch& ws, int lane, double ti, const double* yi, bool burn, D3 bacc,
                         int bref, double* kout) {
    const D3 pos = {yi[0], yi[1], yi[2]};
    D3 sum = {0.0, 0.0, 0.0};
    for (int64_t base = 0; base < E.nb; base += 32) {
        const int64_t b = base + lane;
        D3 a = {0.0, 0.0, 0.0};
        bool ok = true;
        if (b < E.nb) {
            D3 bp;
            ok = spline_position(E, b, ti, &bp);
            if (ok) {  // AccelerationAt::<false>: dir = src - pos; dir * (mu / (n * sqrt(n)))
                const D3 dir = xsub3(bp, pos);
                const double nn = xdot3(dir, dir);
                const double s = xdiv(E.mu[b], xmul(nn, xsqrt(nn)));
                a = xmul3(dir, s);
            }
        }
        if (!__all_sync(kFull, ok)) return false;
        ws.a[lane][0] = a.x;
        ws.a[lane][1] = a.y;
        ws.a[lane][2] = a.z;
        __syncwarp();
        double s = lane == 0 ? sum.x : (lane == 1 ? sum.y : sum.z);
        if (lane < 3) {
            const int cnt = (int)min((int64_t)32, E.nb - base);
            for (int i = 0; i < cnt; ++i) s = xadd(s, ws.a[i][lane]);
        }
        sum.x = __shfl_sync(kFull, s, 0);
        sum.y = __shfl_sync(kFull, s, 1);
        sum.z = __shfl_sync(kFull, s, 2);
        __syncwarp();
    }
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
            const D3 relp = xsub3(pos, rp);
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
    const D3 acc = xadd3(sum, ma);
    kout[0] = yi[3];
    kout[1] = yi[4];
    kout[2] = yi[5];
    kout[3] = acc.x;
    kout[4] = acc.y;
    kout[5] = acc.z;
    return true;
}

__global__ void __launch_bounds__(kShipWarps * 32) k_ships_step_to(ShipsView S, EphemView E, ShipParams P, double t_end,
                                                                   int64_t max_steps) {
    __shared__ WarpScratch scratch[kShipWarps];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t ship = (int64_t)blockIdx.x * kShipWarps + warp;
    if (ship >= S.n) return;
    WarpScratch& ws = scratch[warp];

    double time = S.time[ship], bound = S.bound[ship], next_h = S.next_h[ship];
    double y[6];
    for (int c = 0; c < 6; ++c) y[c] = S.state[6 * ship + c];
    uint32_t rk_i = S.rk_i[ship], n_att = S.n_att[ship];
    int32_t cur = S.cur_seg[ship], status = S.status[ship];
    int64_t nk = S.n_knots[ship];
    unsigned long long evals = S.rhs_evals[ship];
    const int64_t so = S.seg_off[ship];
    double last_t = S.knots[(ship * S.kcap + (nk - 1)) * 7];  // solution.end()

    int64_t accepted = 0;
    while (status == EE_OK && accepted < max_steps && !(last_t >= t_end) && nk < S.kcap) {
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
        // AdaptiveRungeKuttaIntegrator::advance
        const double prev_t = time;
        double prev_y[6];
        for (int c = 0; c < 6; ++c) prev_y[c] = y[c];
        const uint32_t prev_i = rk_i;
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
            // ERK::advance: 13 stages, yi rebuilt from y every stage (explicit.rs:85-90)
            bool ok = true;
            for (int s = 0; s < EE_V87_STAGES; ++s) {
                const double ti = xadd(time, xmul(h, c_v87_c[s]));
                double yi[6];
                for (int c = 0; c < 6; ++c) yi[c] = y[c];
                for (int j = 0; j < s; ++j) {
                    const double ha = xmul(h, c_v87_a[s * 13 + j]);
                    for (int c = 0; c < 6; ++c) yi[c] = xadd(yi[c], xmul(ws.k[j][c], ha));
                }
                double kk[6];
                ok = ship_rhs(E, ws, lane, ti, yi, burn, bacc, bref, kk);
                evals += 1;
                if (!ok) break;
                __syncwarp();
                if (lane == 0)
                    for (int c = 0; c < 6; ++c) ws.k[s][c] = kk[c];
                __syncwarp();
            }
            if (!ok) {
                status = EE_EVAL_FAILED;
                break;
            }
            for (int i = 0; i < EE_V87_STAGES; ++i) {
                const double hb = xmul(h, c_v87_b[i]);
                for (int c = 0; c < 6; ++c) y[c] = xadd(y[c], xmul(ws.k[i][c], hb));
            }
            time = xadd(time, h);
            rk_i += 1;
            n_att += 1;
            // RKEmbedded::error + AbsTol
            double er[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < EE_V87_STAGES; ++i) {
                const double he = xmul(h, c_v87_e[i]);
                for (int c = 0; c < 6; ++c) er[c] = xadd(er[c], xmul(ws.k[i][c], he));
            }
            const double ea = fmax(fabs(xdiv(er[0], P.tol_pos)), fmax(fabs(xdiv(er[1], P.tol_pos)), fabs(xdiv(er[2], P.tol_pos))));
            const double eb = fmax(fabs(xdiv(er[3], P.tol_vel)), fmax(fabs(xdiv(er[4], P.tol_vel)), fabs(xdiv(er[5], P.tol_vel))));
            const double err = fmax(ea, eb);
            // IController::step with order = min(8, 7)
            const double kord = (double)EE_V87_ORDER_EMBEDDED;
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

// ---------------------------------------------------------------------------------------------------------
static const double kEpochMin = -std::numeric_limits<double>::max();  // ftime Duration::MIN / MAX
static const double kEpochMax = std::numeric_limits<double>::max();

Ships::Ships(Ephem* eph, int64_t n_, const double* t0, const double* states, const ee_adaptive_params* p,
             const int64_t* burn_off, const double* bstart, const double* bend, const double* bacc, const int32_t* bref)
    : ephem(eph), n(n_) {
    EE_REQUIRE(eph && n >= 1 && t0 && states && p, "bad arguments");
    EE_REQUIRE(p->pow_mode == EE_POW_GLIBC || p->pow_mode == EE_POW_CORRECTLY_ROUNDED, "unknown pow_mode");
    params = *p;
    EE_CUDA(cudaSetDevice(eph->device));
    EE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    EE_CUDA(cudaEventCreate(&ev0));
    EE_CUDA(cudaEventCreate(&ev1));
    static bool consts_loaded[64] = {false};
    if (!consts_loaded[eph->device % 64]) {
        EE_CUDA(cudaMemcpyToSymbol(c_v87_a, EE_V87_A, sizeof(EE_V87_A)));
        EE_CUDA(cudaMemcpyToSymbol(c_v87_b, EE_V87_B, sizeof(EE_V87_B)));
        EE_CUDA(cudaMemcpyToSymbol(c_v87_c, EE_V87_C, sizeof(EE_V87_C)));
        EE_CUDA(cudaMemcpyToSymbol(c_v87_e, EE_V87_E, sizeof(EE_V87_E)));
        consts_loaded[eph->device % 64] = true;
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
    return ShipsView{s.n,          s.d_time.p,   s.d_bound.p,   s.d_state.p,   s.d_next_h.p,  s.d_rk_i.p,
                     s.d_natt.p,   s.d_cur.p,    s.d_status.p,  s.d_nknots.p,  s.d_evals.p,   s.d_seg_off.p,
                     s.d_seg_burn.p, s.d_seg_end.p, s.d_seg_acc.p, s.d_seg_ref.p, s.knots.p,  s.kcap};
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
    const unsigned grid = (unsigned)((n + kShipWarps - 1) / kShipWarps);
    EE_CUDA(cudaEventRecord(ev0, stream));
    k_ships_step_to<<<grid, kShipWarps * 32, 0, stream>>>(ships_view(*this), view_of(*ephem), P, t_end, max_steps);
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
}

}  // namespace ee
