// ee_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain scalar C++ restatement of the reference's gravitational-integration hot path, operation for
// operation and in the reference's evaluation order, compiled with `-O2 -ffp-contract=off` (Rust never
// contracts a*b+c into an FMA, so neither may this file).  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library; the product
// (ephemeris-explorer_b200/) never links, imports or calls it.
//
// PARITY STATUS: **parity unpinned** at one boundary.  The pair-force arithmetic lives in the un-vendored git
// dependency `particular` 0.8.0-dev @ d490707aa3f3c10382f43fc31c3a1e95c4866516 (Cargo.lock:4278-4280), whose
// source is not under /root/reference.  `pair_accel()` / `accel_at()` below restate particular's published
// scalar formula  dir * (mu / (n * sqrt(n)))  with n = dir.dir + s*s, s = 0  (see DESIGN.md); everything else in
// this file follows source that IS in the tree and cites it.  The reference holds no golden vectors for this
// path (its three tests need network access to JPL Horizons); what pins this oracle is listed in
// tests/test_oracle_*.py: Kepler two-body vs the analytic orbit, conservation laws, the reference's own
// `spacecraft_propagation` scenario (ephemeris/tests/spacecraft_propagation.rs:401-483) replayed on the
// checked-in 10-body state.json, and the convergence behaviour asserted in solar_system_convergence.rs:346-357.
//
// All citations are file:line under /root/reference.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <thread>
#include <vector>

#include "ee_oracle_coeffs.h"
#include "ee_oracle_pow.h"

namespace {

// 0 = libm pow (what Rust's f64::powf calls on this platform), 1 = the engine's portable pow (bit-reproducible on the GPU)
int g_pow_mode = 0;
// Which reading of `particular`'s pair kernel (source not in the reference tree, Cargo.lock:4278-4280) the oracle uses:
// 0 = dir * (mu / (n * sqrt(n))) (the published scalar form, default); 1 = one reciprocal, two products.  Run-time so that
// the GPU twin (ee_set_pair_variant) can be checked bit for bit against both.
int g_pair_variant = 0;
inline double ctrl_pow(double x, double y) { return g_pow_mode ? ora_pow::pow_portable(x, y) : std::pow(x, y); }

// ---------------------------------------------------------------------------------------------------------
// glam::DVec3 restated (glam 0.30.10, Cargo.lock:2890-2891): plain {x,y,z}, lane-wise operators.
struct V3 {
    double x, y, z;
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline void st3v(double* p, V3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}
inline bool is_zero(V3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }
inline double dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline double length_recip(V3 a) { return 1.0 / std::sqrt(dot(a, a)); }
// DVec3::try_normalize: None unless 1/len is finite and > 0.
inline bool try_normalize(V3 a, V3* out) {
    double r = length_recip(a);
    if (std::isfinite(r) && r > 0.0) {
        *out = a * r;
        return true;
    }
    return false;
}
inline V3 normalize(V3 a) { return a * length_recip(a); }
const V3 ZERO3 = {0.0, 0.0, 0.0};

// Double<DVec3> -- the compensated state type the reference's OWN convergence test integrates with
// (ephemeris/tests/solar_system_convergence.rs:12-110).  Only used to replay that test's assertion.
struct DD3 {
    V3 value, error;
};
inline DD3 dd_two_sum(V3 a, V3 b) {
    V3 value = a + b;
    V3 v = value - a;
    V3 error = (a - (value - v)) + (b - v);
    return {value, error};
}
inline DD3 dd_fast_two_sum(V3 a, V3 b) {
    V3 value = a + b;
    V3 error = b - (value - a);
    return {value, error};
}
inline DD3 operator+(DD3 a, DD3 b) {
    DD3 s = dd_two_sum(a.value, b.value);
    return dd_fast_two_sum(s.value, (s.error + a.error) + b.error);
}
inline DD3 operator-(DD3 a, DD3 b) {
    DD3 s = dd_two_sum(a.value, -b.value);
    return dd_fast_two_sum(s.value, (s.error + a.error) - b.error);
}
inline DD3 operator*(DD3 a, double r) { return {a.value * r, a.error * r}; }
inline DD3 operator/(DD3 a, double r) { return {a.value / r, a.error / r}; }
inline V3 value_of(const V3& v) { return v; }
inline V3 value_of(const DD3& d) { return d.value; }
template <class V>
inline V wrap(V3 v);
template <>
inline V3 wrap<V3>(V3 v) {
    return v;
}
template <>
inline DD3 wrap<DD3>(V3 v) {
    return DD3{v, {0.0, 0.0, 0.0}};
}

// ---------------------------------------------------------------------------------------------------------
// particular::gravity::newtonian (NOT in tree -- see header).  Softening literal 0.0 at every call site
// (ephemeris/src/propagators/nbody.rs:29, ephemeris_explorer/src/dynamics/spacecraft.rs:73).
//
// AccelerationPaired for (V, f64):  returns (acceleration of self due to other, acceleration of other due to self).
inline void pair_accel(V3 pi, double mui, V3 pj, double muj, V3* ai, V3* aj) {
    V3 dir = pj - pi;
    double n = dot(dir, dir);
    double ns = n + (0.0 * 0.0);
    double mag = ns * std::sqrt(ns);
    if (g_pair_variant == 1) {
        // alternative reading of particular's paired kernel: one reciprocal, two products
        double inv = 1.0 / mag;
        *ai = dir * (muj * inv);
        *aj = -(dir * (mui * inv));
    } else {
        *ai = dir * (muj / mag);
        *aj = -(dir * (mui / mag));
    }
}
// AccelerationAt::<false> for (V, f64) = (source position, source mu): acceleration at `position`.
inline V3 accel_at(V3 src, double mu, V3 position) {
    V3 dir = src - position;
    double n = dot(dir, dir);
    double ns = n + (0.0 * 0.0);
    return dir * (mu / (ns * std::sqrt(ns)));
}

// NewtonianGravity::eval -- ephemeris/src/propagators/nbody.rs:16-39.  Caller has zeroed ddy.
void gravity_eval(const std::vector<V3>& y, const std::vector<double>& mu, std::vector<V3>& ddy) {
    const size_t n = y.size();
    for (size_t i = 0; i < n; ++i) {
        V3 out_i = ZERO3;
        for (size_t j = i + 1; j < n; ++j) {
            V3 ci, cj;
            pair_accel(y[i], mu[i], y[j], mu[j], &ci, &cj);
            out_i = out_i + ci;
            ddy[j] = ddy[j] + cj;
        }
        ddy[i] = ddy[i] + out_i;
    }
}

enum Status : int32_t {  // integration/src/lib.rs:312-318 (+ solout exit, nbody.rs:44-47)
    OK = 0,
    STEP_SIZE_UNDERFLOW = 1,
    MAX_ITERATIONS_REACHED = 2,
    BOUND_REACHED = 3,
    EVAL_FAILED = 4,
    SOLOUT_EXIT = 5,
};

// ---------------------------------------------------------------------------------------------------------
// Polynomial / UniformSpline -- ephemeris/src/trajectory.rs:337-633
struct Poly {
    int n = 0;      // number of coefficients (after trim)
    V3 c[9];        // lowest order first; SmallVec<[DVec3; 8]> (degree <= 8 possible: min(degree, 8))
};

// eval_slice_horner -- trajectory.rs:398-410
inline V3 horner(const V3* c, int n, double t) {
    V3 r = ZERO3;
    for (int i = n - 1; i >= 0; --i) r = r * t + c[i];
    return r;
}
inline V3 horner_v(const V3* c, int n, V3 t) {  // lane-wise variable (used by the LSQ fit where U = &f64 is splat)
    V3 r = ZERO3;
    for (int i = n - 1; i >= 0; --i) r = r * t + c[i];
    return r;
}

// Polynomial::eval_and_deriv -- trajectory.rs:368-385
inline void poly_eval_and_deriv(const Poly& p, double t, V3* ev, V3* de) {
    V3 first = p.n ? p.c[0] : ZERO3;
    V3 last = p.n ? p.c[p.n - 1] : ZERO3;
    V3 eval = last, deriv = last;
    // coeffs.iter().skip(1).rev().skip(1): c[n-2] .. c[1]
    for (int i = p.n - 2; i >= 1; --i) {
        eval = eval * t + p.c[i];
        deriv = deriv * t + eval;
    }
    eval = eval * t + first;
    *ev = eval;
    *de = deriv;
}

struct Spline {
    double start = 0.0;     // Epoch (TAI seconds)
    double interval = 0.0;  // Duration
    std::deque<Poly> polys;
    double span() const { return interval * (double)polys.size(); }  // trajectory.rs:625-627
    double end() const { return start + span(); }                    // :428-430
    // get_polynomial -- trajectory.rs:561-569 with get_index_local_exclusive :600-617
    bool get_polynomial(double at, const Poly** p, double* tau) const {
        double local = at - start;
        if (std::signbit(local) || local > span()) return false;
        double q = std::ceil(local / interval);
        size_t idx = (size_t)q;  // `as usize`: saturating, q >= 0 here
        idx = idx ? idx - 1 : 0;
        double local_poly = local - interval * (double)idx;
        double norm = local_poly / interval;
        if (idx >= polys.size()) return false;
        *p = &polys[idx];
        *tau = norm;
        return true;
    }
    bool position(double at, V3* out) const {  // trajectory.rs:449-452
        const Poly* p;
        double tau;
        if (!get_polynomial(at, &p, &tau)) return false;
        *out = horner(p->c, p->n, tau);
        return true;
    }
    bool state_vector(double at, V3* pos, V3* vel) const {  // trajectory.rs:455-470
        const Poly* p;
        double tau;
        if (!get_polynomial(at, &p, &tau)) return false;
        V3 e, d;
        poly_eval_and_deriv(*p, tau, &e, &d);
        *pos = e;
        *vel = d / interval;
        return true;
    }
};

// LeastSquaresFit::interpolate -- ephemeris_explorer/src/dynamics/celestial.rs:24-136
// (identical twin: ephemeris/tests/spacecraft_propagation.rs:19-131).  Lane-wise DVec3 arithmetic.
bool lsq_fit(int degree_req, const double* ts, const V3* xs, int data_len, Poly* out) {
    const V3 ONE = {1.0, 1.0, 1.0};
    V3 d0 = ZERO3, gamma0 = ZERO3, b0 = ZERO3;
    for (int i = 0; i < data_len; ++i) {
        d0 = d0 + xs[i];
        gamma0 = gamma0 + ONE;
        b0 = b0 + v3(ts[i], ts[i], ts[i]);
    }
    if (is_zero(gamma0)) return false;
    int degree = std::min(degree_req, data_len - 1);
    b0 = b0 / gamma0;
    d0 = d0 / gamma0;
    Poly P;
    if (degree == 0) {
        P.n = 1;
        P.c[0] = d0;
        *out = P;
        return true;
    }
    V3 p_data[9], buf_a[9], buf_b[9];
    for (int i = 0; i <= degree; ++i) p_data[i] = buf_a[i] = buf_b[i] = ZERO3;
    V3* p_km1 = buf_a;
    V3* p_k = buf_b;
    p_data[0] = d0;
    p_k[0] = ONE;
    V3 gamma_k = gamma0, b_k = b0, minus_c_k = ZERO3;
    int kp1 = 1;
    for (;;) {
        for (int i = 0; i < kp1; ++i) p_km1[i] = minus_c_k * p_km1[i] - b_k * p_k[i];
        for (int im1 = 0; im1 < kp1; ++im1) p_km1[im1 + 1] = p_km1[im1 + 1] + p_k[im1];
        V3 d_kp1 = ZERO3, gamma_kp1 = ZERO3, b_kp1 = ZERO3;
        for (int s = 0; s < data_len; ++s) {
            V3 px = horner(p_km1, kp1 + 1, ts[s]);
            V3 wipx = px;
            d_kp1 = d_kp1 + xs[s] * wipx;
            V3 wipxpx = wipx * px;
            gamma_kp1 = gamma_kp1 + wipxpx;
            b_kp1 = b_kp1 + wipxpx * ts[s];  // f64 * DVec3 (commutative per lane)
        }
        if (is_zero(gamma_kp1)) break;
        d_kp1 = d_kp1 / gamma_kp1;
        for (int i = 0; i < kp1 + 1; ++i) p_data[i] = p_data[i] + d_kp1 * p_km1[i];
        if (kp1 == degree) break;
        b_kp1 = b_kp1 / gamma_kp1;
        kp1 += 1;
        b_k = b_kp1;
        minus_c_k = -(gamma_kp1 / gamma_k);
        gamma_k = gamma_kp1;
        std::swap(p_k, p_km1);
    }
    P.n = degree + 1;
    for (int i = 0; i <= degree; ++i) P.c[i] = p_data[i];
    while (P.n > 0 && is_zero(P.c[P.n - 1])) P.n--;  // Polynomial::trim -- trajectory.rs:387-395
    *out = P;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// SplineInterpolator(s) solout -- ephemeris/src/propagators/nbody.rs:243-517
struct Interp {
    double last_sample_time = 0.0;
    double sample_period = 0.0;
    int degree = 0;
    int index = 1;  // PolyonmialInterpolator::new: index 1, samples filled with the initial position
    V3 samples[9];
    double time() const {  // nbody.rs:318-322
        int len = index;
        return last_sample_time + sample_period * (double)(len > 0 ? len - 1 : 0);
    }
};

struct MethodCoeffs {
    int order;
    const double* neg_alpha;
    const double* beta;
    double inv_beta_d;
    const double* cow_beta;
    double cow_inv_beta_d;
};
const MethodCoeffs QT12 = {12, EE_QT12_NEG_ALPHA, EE_QT12_BETA, EE_QT12_INV_BETA_D, EE_COWELL12_BETA,
                           EE_COWELL12_INV_BETA_D};
const MethodCoeffs ST13 = {13, EE_ST13_NEG_ALPHA, EE_ST13_BETA, EE_ST13_INV_BETA_D, EE_COWELL13_BETA,
                           EE_COWELL13_INV_BETA_D};

struct INBody {
    virtual ~INBody() {}
    virtual int32_t step() = 0;
    virtual void get_state(double* t, double* pos, double* vel, double* acc) = 0;
    virtual uint64_t eval_count() const = 0;
};

template <class V>
struct Slot {  // StepOrder2 -- second_order/mod.rs:23-27
    std::vector<V> y, dy, ddy;
};

template <class V>
struct NBodyT : INBody {
    // ODEProblem -- integration/src/problem.rs:1-8
    double time, bound;
    std::vector<V> y, dy;
    std::vector<double> mu;
    std::vector<V3> scratch_y, scratch_a;
    // LinearMultistepIntegrator {h, lm: ELM2, starter: Substepper<4, SRKN<BlanesMoan6B>>}
    double h;
    MethodCoeffs mc;
    uint32_t lm_i = 0;
    std::vector<V> current_ddy, sum1, sum2;
    std::vector<Slot<V>> steps;
    size_t head;
    double starter_h;          // h * (1/4) -- multistep/mod.rs:53-58
    uint32_t srkn_i = 0;
    std::vector<V> srkn_ddy;
    uint64_t evals = 0;
    // solout
    bool has_solout = false;
    bool backward = false;
    double delta = 0.0;
    std::vector<Interp> interps;
    std::vector<Spline> solution;

    // `.zero()` then NewtonianGravity::eval on the plain values (for Double<DVec3> the test's twin of eval adds
    // into `.value` only: solar_system_convergence.rs:117-140)
    void eval(std::vector<V>& out) {
        scratch_y.resize(y.size());
        scratch_a.assign(y.size(), ZERO3);
        for (size_t k = 0; k < y.size(); ++k) scratch_y[k] = value_of(y[k]);
        gravity_eval(scratch_y, mu, scratch_a);
        for (size_t k = 0; k < y.size(); ++k) out[k] = wrap<V>(scratch_a[k]);
        evals++;
    }
    Slot<V>& front() { return steps[head]; }
    // ELM2::prepare_next_step -- second_order/mod.rs:40-45; LMBuffer::rotate_right -- buffer.rs:35-37
    void prepare_next_step() {
        head = (head + steps.size() - 1) % steps.size();
        std::swap(front().ddy, current_ddy);
    }
    // SRKN<C>::advance -- runge_kutta/nystrom/symplectic.rs:70-102; C = BlanesMoan6B (the multistep starter) or
    // BlanesMoan14A (a fixed-step method in its own right, methods.rs:1730-1774); both FSAL = true
    bool srkn_main = false;  // method id 14: NBodyPropagator<.., FixedMethod<BlanesMoan14A>>
    void srkn_advance(double hs) {
        const int stages = srkn_main ? EE_BM14A_STAGES : EE_BM6B_STAGES;
        const double* CA = srkn_main ? EE_BM14A_A : EE_BM6B_A;
        const double* CB = srkn_main ? EE_BM14A_B : EE_BM6B_B;
        for (int s = 0; s < stages; ++s) {
            if (s > 0 || srkn_i == 0) eval(srkn_ddy);  // FSAL = true
            const double hb = hs * CB[s];
            const double ha = hs * CA[s];
            for (size_t k = 0; k < y.size(); ++k) {
                dy[k] = dy[k] + srkn_ddy[k] * hb;
                y[k] = y[k] + dy[k] * ha;
            }
        }
        time = time + hs;
        srkn_i += 1;
    }
    // FixedRungeKuttaIntegrator::advance -- runge_kutta/mod.rs:106-126
    int32_t frk_advance(double hs) {
        if (time >= bound) return BOUND_REACHED;
        if (time + hs == time) return STEP_SIZE_UNDERFLOW;
        srkn_advance(hs);
        return OK;
    }
    // SubstepperIntegrator<4>::advance -- multistep/mod.rs:97-108
    int32_t starter_advance() {
        for (int k = 0; k < 4; ++k) {
            int32_t st = frk_advance(starter_h);
            if (st) return st;
        }
        return OK;
    }
    uint32_t starter_step_count() const { return srkn_i / 4; }  // multistep/mod.rs:92-94
    // ELM2::advance_with -- second_order/mod.rs:134-153
    int32_t advance_with(bool run_starter) {
        prepare_next_step();
        front().y = y;
        front().dy = dy;
        if (run_starter) {
            int32_t st = starter_advance();
            if (st) return st;
        }
        eval(current_ddy);
        return OK;
    }
    // ELM2::advance -- second_order/mod.rs:91-131, Cowell::update_velocity -- cowell.rs:19-53
    void lm_advance() {
        const size_t n = y.size();
        const size_t m = steps.size();
        std::fill(sum1.begin(), sum1.end(), wrap<V>(ZERO3));
        std::fill(sum2.begin(), sum2.end(), wrap<V>(ZERO3));
        for (int j = 0; j < mc.order; ++j) {
            const std::vector<V>& yy = j == 0 ? y : steps[(head + (size_t)(j - 1)) % m].y;
            const std::vector<V>& aa = j == 0 ? current_ddy : steps[(head + (size_t)(j - 1)) % m].ddy;
            const double ca = 1.0 * mc.neg_alpha[j];
            const double cb = 1.0 * mc.beta[j];
            for (size_t k = 0; k < n; ++k) {
                sum1[k] = sum1[k] + yy[k] * ca;
                sum2[k] = sum2[k] + aa[k] * cb;
            }
        }
        prepare_next_step();
        std::swap(front().y, y);
        std::swap(front().dy, dy);
        const double f = h * h * mc.inv_beta_d;  // (h*h) * (1/beta_D)
        for (size_t k = 0; k < n; ++k) y[k] = sum1[k] + sum2[k] * f;
        time = time + h;
        eval(current_ddy);
        // Cowell velocity
        std::fill(sum1.begin(), sum1.end(), wrap<V>(ZERO3));
        for (int j = 0; j < mc.order; ++j) {
            const std::vector<V>& aa = j == 0 ? current_ddy : steps[(head + (size_t)(j - 1)) % m].ddy;
            const double c = 1.0 * mc.cow_beta[j];
            for (size_t k = 0; k < n; ++k) sum1[k] = sum1[k] + aa[k] * c;
        }
        const double g = h * mc.cow_inv_beta_d;
        const std::vector<V>& ym1 = front().y;
        for (size_t k = 0; k < n; ++k) dy[k] = (y[k] - ym1[k]) / h + sum1[k] * g;
        lm_i += 1;
    }
    // LinearMultistepIntegrator::advance -- multistep/mod.rs:194-225
    int32_t advance() {
        if (srkn_main) return frk_advance(h);  // FixedRungeKuttaIntegrator::advance -- runge_kutta/mod.rs:106-126
        if (time >= bound) return BOUND_REACHED;
        if (time + h == time) return STEP_SIZE_UNDERFLOW;
        if (starter_step_count() < (uint32_t)mc.order) {
            if (starter_step_count() == 0) {
                int32_t st = advance_with(false);
                if (st) return st;
            }
            return advance_with(true);
        }
        lm_advance();
        return OK;
    }
    // SplineBound::samples -- nbody.rs:422-442
    static void sample_ts(bool backward, double* ts) {
        for (int i = 0; i < 9; ++i) ts[i] = backward ? 1.0 - (double)i / 8.0 : (double)i / 8.0;
    }
    // Solout::new_solution -- nbody.rs:455-468
    void new_solution() {
        solution.assign(interps.size(), Spline());
        for (size_t b = 0; b < interps.size(); ++b) {
            double it = -interps[b].time();
            solution[b].start = backward ? time - it : time + it;  // D::offset(problem.time, -interp.time())
            solution[b].interval = interps[b].sample_period * 8.0;
        }
    }
    // Solout::solout -- nbody.rs:471-489 via solout_with :372-400
    bool solout() {
        double ts[9];
        sample_ts(backward, ts);
        for (size_t b = 0; b < interps.size(); ++b) {
            Interp& in = interps[b];
            in.last_sample_time = in.last_sample_time + delta;
            if (in.last_sample_time == in.sample_period) {
                in.last_sample_time = 0.0;
                // PolyonmialInterpolator::push -- nbody.rs:261-269 (assert index < LEN)
                in.samples[in.index] = value_of(y[b]);
                in.index += 1;
                if (in.index == 9) {  // is_full -> try_to_polynomial
                    Poly p;
                    if (!lsq_fit(in.degree, ts, in.samples, 9, &p)) return false;
                    Spline& sp = solution[b];
                    if (backward) {  // push_front -- trajectory.rs:496-499
                        sp.polys.push_front(p);
                        sp.start = sp.start - sp.interval;
                    } else {
                        sp.polys.push_back(p);
                    }
                    std::swap(in.samples[0], in.samples[in.index - 1]);  // finish -- nbody.rs:302-306
                    in.index = 1;
                }
            }
        }
        return true;
    }
    // Integration::advance -> Integrator::advance_solution -- integration/src/lib.rs:379-391, :497-503;
    // NBodyPropagator::step -- nbody.rs:200-207
    int32_t step() override {
        int32_t st = advance();
        if (st) return st;
        if (has_solout && !solout()) return SOLOUT_EXIT;
        return OK;
    }
    void get_state(double* t, double* pos, double* vel, double* acc) override {
        if (t) *t = time;
        for (size_t i = 0; i < y.size(); ++i) {
            if (pos) st3v(pos + 3 * i, value_of(y[i]));
            if (vel) st3v(vel + 3 * i, value_of(dy[i]));
            if (acc) st3v(acc + 3 * i, value_of(srkn_main ? srkn_ddy[i] : current_ddy[i]));  // the integrator's last evaluation
        }
    }
    uint64_t eval_count() const override { return evals; }
    void init(int64_t n, const double* pos, const double* vel, const double* mus, double t0, double h_signed, int method) {
        mc = method == 13 ? ST13 : QT12;
        srkn_main = method == 14;
        time = t0;
        bound = std::numeric_limits<double>::infinity();  // nbody.rs:112
        h = h_signed;
        starter_h = h_signed * (1.0 / 4.0);
        y.resize((size_t)n);
        dy.resize((size_t)n);
        mu.assign(mus, mus + n);
        for (int64_t i = 0; i < n; ++i) {
            y[(size_t)i] = wrap<V>(V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
            dy[(size_t)i] = wrap<V>(V3{vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]});
        }
        // ELM2::from_problem -- second_order/mod.rs:74-88
        std::vector<V> zero((size_t)n, wrap<V>(ZERO3));
        current_ddy = zero;
        sum1 = zero;
        sum2 = zero;
        steps.assign((size_t)(mc.order - 1), Slot<V>{y, dy, zero});
        head = steps.size();  // LMBuffer::from_iter: head = len -- buffer.rs:9-16
        srkn_ddy = zero;      // SRKN::from_problem -- symplectic.rs:61-67
    }
};
typedef NBodyT<V3> NBody;

// ---------------------------------------------------------------------------------------------------------
// Massless-ship propagator
struct Ephem {
    std::vector<double> mu;
    std::vector<Spline> splines;
};

struct Burn {  // (start, end, ConstantThrust{acceleration, frame}) -- spacecraft.rs:30-57
    double start, end;
    V3 acc;
    int reference;  // body index, or -1 = inertial
};
struct Segment {  // spacecraft.rs:59-70
    bool burn;
    double start, end;
    V3 acc;
    int reference;
};
const double EPOCH_MIN = -std::numeric_limits<double>::max();  // ftime Duration::MIN/MAX -- duration.rs:9-13
const double EPOCH_MAX = std::numeric_limits<double>::max();

struct SV {  // StateVector<DVec3> -- trajectory.rs:5-9; lane-wise ops :56-122
    V3 p, v;
};
inline SV operator+(SV a, SV b) { return {a.p + b.p, a.v + b.v}; }
inline SV operator-(SV a, SV b) { return {a.p - b.p, a.v - b.v}; }
inline SV operator*(SV a, double s) { return {a.p * s, a.v * s}; }

struct Knot {
    double t;
    SV sv;
};

struct Ship {
    const Ephem* eph;
    // problem
    double time, bound;
    SV state;
    // SpacecraftModel -- spacecraft.rs:224-256
    std::vector<Segment> timeline;
    size_t current_segment;
    // method params (AdaptiveRungeKutta) -- runge_kutta/mod.rs:128-156
    double h_init, h_max, tol_pos, tol_vel, fac_min, fac_max, fac;
    uint32_t n_max;
    // AdaptiveRungeKuttaIntegrator -- runge_kutta/mod.rs:286-296
    double frk_h, next_h;
    uint32_t rk_i, n;
    // the selectable method (flight_plan.rs:175-184): ERK<C,[State;S]> (explicit.rs:39-107) or, for Fine45,
    // ERKNG<C,[V;S],V> (nystrom/explicit_generalized.rs:38-138) whose dk[s] live in k[s].v
    const EeRkTableau* tab = &EE_RK_TABLE[0];
    SV k[EE_RK_MAX_STAGES];
    SV error;
    double prev_t;
    SV prev_y;
    uint32_t prev_rk_i;
    SV prev_klast;  // PreviousStep.rk carries k[STAGES-1] of an FSAL method (undo_step, explicit.rs:134-140)
    // solution: CubicHermiteSpline knots -- spacecraft.rs:645-695
    std::vector<Knot> knots;
    uint64_t rhs_evals = 0;
    // SpacecraftSolout's analytics (ephemeris_explorer/src/dynamics/spacecraft.rs:448-586), optional
    bool analytics = false;
    std::vector<double> soi_radius;
    struct Transition {
        double t;
        int body;
    };
    struct ApsisRec {
        int kind;  // 0 = periapsis, 1 = apoapsis
        double t, distance;
        int body;
    };
    std::vector<Transition> transitions;  // SoiTransitions: sorted by time
    std::vector<ApsisRec> apsides;        // Apsides: sorted by time

    // Timeline::new -- spacecraft.rs:131-160
    void build_timeline(std::vector<Burn> burns) {
        std::stable_sort(burns.begin(), burns.end(), [](const Burn& a, const Burn& b) { return a.start < b.start; });
        double cursor = EPOCH_MIN;
        timeline.clear();
        for (const Burn& b : burns) {
            if (b.start > cursor) timeline.push_back({false, cursor, b.start, ZERO3, -1});
            cursor = b.end;
            timeline.push_back({true, b.start, b.end, b.acc, b.reference});
        }
        if (cursor < EPOCH_MAX) timeline.push_back({false, cursor, EPOCH_MAX, ZERO3, -1});
    }
    // Timeline::segment_idx_at -- spacecraft.rs:167-170 (partition_point(seg.end() <= time))
    size_t segment_idx_at(double t) const {
        size_t i = 0;
        while (i < timeline.size() && timeline[i].end <= t) ++i;
        return i;
    }
    // Method::init -- runge_kutta/mod.rs:323-343
    void init_integrator() {
        frk_h = h_init;
        next_h = h_init;
        rk_i = 0;
        n = 0;
        for (auto& kk : k) kk = state;  // ERK::from_problem clones the state into every k slot
        error = state;
        prev_t = time;
        prev_y = state;
        prev_rk_i = 0;
        prev_klast = state;
    }
    // Bodies::acceleration -- ephemeris_explorer/src/dynamics/spacecraft.rs:218-229 (test twin
    // ephemeris/tests/spacecraft_propagation.rs:232-241): body order = construction (IndexMap) order.
    bool context_acceleration(double t, const SV& sv, V3* out) const {
        V3 acc = ZERO3;
        for (size_t b = 0; b < eph->splines.size(); ++b) {
            V3 bp;
            if (!eph->splines[b].position(t, &bp)) return false;
            acc = acc + accel_at(bp, eph->mu[b], sv.p);
        }
        *out = acc;
        return true;
    }
    // Segment::acceleration -> ConstantThrust::acceleration -> ReferenceFrame::transform -> TNB
    //   spacecraft.rs:46-56, :106-116; dynamics/spacecraft.rs:240-293
    bool manoeuvre_acceleration(double t, const SV& sv, V3* out) const {
        const Segment& seg = timeline[current_segment];
        if (!seg.burn) {
            *out = ZERO3;
            return true;
        }
        if (seg.reference < 0) {  // TNB::IDENTITY.mul_vec3
            V3 res = v3(1, 0, 0) * seg.acc.x;
            res = res + v3(0, 1, 0) * seg.acc.y;
            res = res + v3(0, 0, 1) * seg.acc.z;
            *out = res;
            return true;
        }
        V3 rp, rv;
        if (!eph->splines[(size_t)seg.reference].state_vector(t, &rp, &rv)) return false;
        SV rel = sv - SV{rp, rv};
        V3 x, yv;
        if (!try_normalize(rel.v, &x)) return false;
        if (!try_normalize(cross(rel.p, rel.v), &yv)) return false;
        V3 z = normalize(cross(x, yv));
        // DMat3::from_cols(x, z, y).mul_vec3(v) = x*v.x + z*v.y + y*v.z
        V3 res = x * seg.acc.x;
        res = res + z * seg.acc.y;
        res = res + yv * seg.acc.z;
        *out = res;
        return true;
    }
    // FirstOrderODE for SpacecraftModel -- spacecraft.rs:283-309
    bool rhs(double t, const SV& yv, SV* dy) {
        rhs_evals++;
        V3 ca, ma;
        if (!context_acceleration(t, yv, &ca)) return false;
        if (!manoeuvre_acceleration(t, yv, &ma)) return false;
        dy->v = ca + ma;
        dy->p = yv.v;
        return true;
    }
    // ERK<C,[_;S]>::advance -- runge_kutta/explicit.rs:73-106
    bool erk_advance(double h) {
        const EeRkTableau& T = *tab;
        const int S = T.stages;
        SV yi = state;
        for (int s = 0; s < S; ++s) {
            if (T.fsal && s == 0 && rk_i > 0) {  // :77-80
                std::swap(k[0], k[S - 1]);
                continue;
            }
            double ti = time + h * T.c[s];
            yi = state;
            for (int j = 0; j < s; ++j) yi = yi + k[j] * (h * T.a[s * (s - 1) / 2 + j]);
            k[s] = SV{ZERO3, ZERO3};
            if (!rhs(ti, yi, &k[s])) return false;
        }
        for (int i = 0; i < S; ++i) state = state + k[i] * (h * T.b[i]);
        time = time + h;
        rk_i += 1;
        return true;
    }
    // ERKNG<Fine45,[V;7],V>::advance -- runge_kutta/nystrom/explicit_generalized.rs:77-138; the ODE is
    // SecondOrderODEGeneral for SpacecraftModel -- spacecraft.rs:311-332 (same two accelerations)
    bool erkng_advance(double h) {
        const EeRkTableau& T = *tab;
        const int S = T.stages;
        for (int s = 0; s < S; ++s) {
            if (T.fsal && s == 0 && rk_i > 0) {
                std::swap(k[0], k[S - 1]);
                continue;
            }
            double ti = time + h * T.c[s];
            V3 yi = state.p;
            yi = yi + state.v * (h * T.c[s]);
            V3 dyi = state.v;
            for (int j = 0; j < s; ++j) {
                yi = yi + k[j].v * (h * h * T.a[s * (s - 1) / 2 + j]);
                dyi = dyi + k[j].v * (h * T.a2[s * (s - 1) / 2 + j]);
            }
            k[s].v = ZERO3;
            rhs_evals++;
            V3 ca, ma;
            SV sv{yi, dyi};
            if (!context_acceleration(ti, sv, &ca)) return false;
            if (!manoeuvre_acceleration(ti, sv, &ma)) return false;
            k[s].v = ca + ma;
        }
        state.p = state.p + state.v * h;
        for (int i = 0; i < S; ++i) {
            state.p = state.p + k[i].v * (h * h * T.b[i]);
            state.v = state.v + k[i].v * (h * T.b2[i]);
        }
        time = time + h;
        rk_i += 1;
        return true;
    }
    // AdaptiveRungeKuttaIntegrator::advance -- runge_kutta/mod.rs:396-440
    int32_t integrator_advance() {
        const EeRkTableau& T = *tab;
        prev_t = time;  // PreviousStep::store -- :253-268: time, state, rk.undo_step(rk) = (i, k[last] when FSAL)
        prev_y = state;
        prev_rk_i = rk_i;
        if (T.fsal) prev_klast = k[T.stages - 1];
        for (;;) {
            if (n > n_max) return MAX_ITERATIONS_REACHED;
            if (time + next_h > bound) next_h = bound - time;
            frk_h = next_h;
            // FixedRungeKuttaIntegrator::advance -- :106-126
            if (time >= bound) return BOUND_REACHED;
            if (time + frk_h == time) return STEP_SIZE_UNDERFLOW;
            if (!(T.kind ? erkng_advance(frk_h) : erk_advance(frk_h))) return EVAL_FAILED;
            n += 1;
            // RKEmbedded::error -- explicit.rs:124-132 / explicit_generalized.rs:156-175
            error = SV{ZERO3, ZERO3};
            if (T.kind) {
                for (int i = 0; i < T.stages; ++i) {
                    error.p = error.p + k[i].v * (frk_h * frk_h * T.e[i]);
                    error.v = error.v + k[i].v * (frk_h * T.e2[i]);
                }
            } else {
                for (int i = 0; i < T.stages; ++i) error = error + k[i] * (frk_h * T.e[i]);
            }
            // AbsTol::err_over_tol -- dynamics/spacecraft.rs:615-640 (same expression for both state shapes)
            V3 ep = error.p / tol_pos, ev = error.v / tol_vel;
            double a = std::fmax(std::fabs(ep.x), std::fmax(std::fabs(ep.y), std::fabs(ep.z)));
            double b = std::fmax(std::fabs(ev.x), std::fmax(std::fabs(ev.y), std::fabs(ev.z)));
            double err = std::fmax(a, b);
            // IController::step -- runge_kutta/mod.rs:225-243; order = LOWER_ORDER = min(ORDER, ORDER_EMBEDDED)
            double kk = (double)T.kord;
            double m = fac * ctrl_pow(err, -(1.0 / kk));
            double cl = m < fac_min ? fac_min : (m > fac_max ? fac_max : m);  // num_traits::clamp
            double nh = next_h * cl;
            next_h = nh > h_max ? h_max : nh;  // clamp_max
            if (err <= 1.0) break;
            time = prev_t;  // PreviousStep::restore -- :270-284
            state = prev_y;
            rk_i = prev_rk_i;
            if (T.fsal) k[T.stages - 1] = prev_klast;
        }
        return OK;
    }
    // ---- SpacecraftSolout: SOI transitions and apsides found on the Hermite segment of every accepted step
    struct Hermite {  // CubicHermite::new / eval / eval_derivative -- trajectory.rs:645-698
        double t0, t1;
        V3 a0, a1, a2, a3;
        Hermite(const Knot& k0, const Knot& k1) : t0(k0.t), t1(k1.t), a0(k0.sv.p), a1(k0.sv.v), a2(ZERO3), a3(ZERO3) {
            const V3 p0 = k0.sv.p, v0 = k0.sv.v, p1 = k1.sv.p, v1 = k1.sv.v;
            const double dt = t1 - t0;
            const bool same = p0.x == p1.x && p0.y == p1.y && p0.z == p1.z && v0.x == v1.x && v0.y == v1.y && v0.z == v1.z;
            if (!(dt == 0.0 && same)) {
                const double r = 1.0 / dt, r2 = r * r, r3 = r * r2;
                const V3 dv = p1 - p0;
                a2 = dv * r2 * 3.0 - (v0 * 2.0 + v1) * r;
                a3 = dv * r3 * -2.0 + (v0 + v1) * r2;
            }
        }
        V3 eval(double t) const {
            const double d = t - t0;
            return (((a3 * d + a2) * d) + a1) * d + a0;
        }
        V3 deriv(double t) const {
            const double d = t - t0;
            return ((a3 * d * 3.0 + a2 * 2.0) * d) + a1;
        }
    };
    static double signum(double x) { return std::isnan(x) ? x : (std::signbit(x) ? -1.0 : 1.0); }  // f64::signum
    // GravitationalBody::soi_distance_squared_at -- dynamics/spacecraft.rs:76-82
    bool f_soi(int b, const Hermite& H, double t, double* out) const {
        V3 bp;
        if (!eph->splines[(size_t)b].position(t, &bp)) return false;
        const V3 d = H.eval(t) - bp;
        *out = dot(d, d) - soi_radius[(size_t)b] * soi_radius[(size_t)b];
        return true;
    }
    // GravitationalBody::radial_velocity_at -- :84-88
    bool f_radial(int b, const Hermite& H, double t, double* out) const {
        V3 bp, bv;
        if (!eph->splines[(size_t)b].state_vector(t, &bp, &bv)) return false;
        const SV rel = SV{H.eval(t), H.deriv(t)} - SV{bp, bv};
        *out = dot(rel.p, rel.v);
        return true;
    }
    // find_zero_crossing -- :112-161.  Returns 0 = none, 1 = ascending, 2 = descending.
    template <typename F>
    static int find_zero_crossing(double t0, double t1, F f, double* when) {
        double f0, f1;
        if (!f(t0, &f0)) return 0;
        if (!f(t1, &f1)) return 0;
        if (signum(f0) == signum(f1)) return 0;
        const bool ascending = std::signbit(f0);
        double x0 = t0, x1 = t1;
        for (int it = 0; it < 100; ++it) {
            const double mid = x0 + (x1 - x0) / 2.0;
            double fm = 0.0;
            f(mid, &fm);  // the reference unwraps: mid lies between two epochs both trajectories cover
            if (signum(f0) != signum(fm)) {
                x1 = mid;
            } else {
                x0 = mid;
                f0 = fm;
            }
            if (std::fabs(x1 - x0) < 1e-3) {
                *when = x0;
                return ascending ? 1 : 2;
            }
        }
        return 0;
    }
    // Bodies::soi_at_except + find_soi -- :174-216: the closest body whose sphere contains the point (first on ties)
    int soi_at_except(double t, V3 pos, int except) const {
        int best = -1;
        double best_d = 0.0;
        for (size_t b = 0; b < eph->splines.size(); ++b) {
            if ((int)b == except) continue;
            V3 bp;
            if (!eph->splines[b].position(t, &bp)) continue;
            const V3 d = pos - bp;
            const double d2 = dot(d, d);
            if (!(d2 < soi_radius[b] * soi_radius[b])) continue;
            if (best < 0 || d2 < best_d) {  // min_by(total_cmp) keeps the first of equal minima; distances are never NaN/-0 here
                best = (int)b;
                best_d = d2;
            }
        }
        return best;
    }
    // SoiTransitions::insert -- :340-347
    void insert_transition(double t, int body) {
        size_t i = 0;
        while (i < transitions.size() && transitions[i].t < t) ++i;
        if (i < transitions.size() && transitions[i].t == t) {
            transitions[i] = {t, body};
        } else if (i > 0 && transitions[i - 1].body == body) {
        } else {
            transitions.insert(transitions.begin() + (long)i, Transition{t, body});
        }
    }
    // Apsides::insert -- :421-427
    void insert_apsis(const ApsisRec& a) {
        size_t i = 0;
        while (i < apsides.size() && apsides[i].t < a.t) ++i;
        if (i < apsides.size() && apsides[i].t == a.t)
            apsides[i] = a;
        else
            apsides.insert(apsides.begin() + (long)i, a);
    }
    // SpacecraftSolout::new_solution -- :518-533
    void new_analytics() {
        transitions.clear();
        apsides.clear();
        const int cur = soi_at_except(time, state.p, -1);
        if (cur >= 0) transitions.push_back({time, cur});
    }
    // SpacecraftSolout::solout after the knot was pushed -- :536-586
    void analyse_step() {
        const Hermite H(knots[knots.size() - 2], knots[knots.size() - 1]);
        const double t0 = H.t0, t1 = H.t1;
        for (size_t b = 0; b < eph->splines.size(); ++b) {
            double when = 0.0;
            const int dir = find_zero_crossing(t0, t1, [&](double t, double* o) { return f_soi((int)b, H, t, o); }, &when);
            if (dir == 2) {
                insert_transition(when, (int)b);
            } else if (dir == 1) {
                const int entered = soi_at_except(when, H.eval(when), (int)b);
                if (entered >= 0) insert_transition(when, entered);
            }
        }
        // SoiTransitions::starting_at(t0) -- :330-333
        size_t first = 0;
        {
            size_t i = 0;
            while (i < transitions.size() && transitions[i].t < t0) ++i;
            if (i < transitions.size() && transitions[i].t == t0)
                first = i;
            else
                first = i == 0 ? 0 : i - 1;
        }
        for (size_t i = first; i < transitions.size(); ++i) {
            const double a = std::max(transitions[i].t, t0);
            const double b = i + 1 < transitions.size() ? transitions[i + 1].t : t1;
            const int soi = transitions[i].body;
            double when = 0.0;
            const int dir = find_zero_crossing(a, b, [&](double t, double* o) { return f_radial(soi, H, t, o); }, &when);
            if (!dir) continue;
            V3 bp;
            if (!eph->splines[(size_t)soi].position(when, &bp)) continue;  // Trajectory::distance_at -- dynamics/mod.rs:141-146
            const V3 d = bp - H.eval(when);
            insert_apsis(ApsisRec{dir == 1 ? 0 : 1, when, std::sqrt(dot(d, d)), soi});
        }
    }
    // SpacecraftPropagator::step -- spacecraft.rs:599-615
    int32_t step() {
        if (time >= timeline[current_segment].end) {  // advance_timeline -- :250-256
            current_segment += 1;
            bound = timeline[current_segment].end;
            init_integrator();  // reset_integrator -- :480-485
        }
        int32_t st = integrator_advance();
        if (st) return st;
        knots.push_back({time, state});  // CubicHermiteSplineSolout::solout -- :667-679
        if (analytics) analyse_step();
        return OK;
    }
};

inline V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }
inline void st3(double* p, V3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

}  // namespace

// =========================================================================================================
// C entry points (ctypes).  AoS double[3] everywhere, matching Vec<DVec3>.
extern "C" {

void ora_set_pow_mode(int32_t mode) { g_pow_mode = mode; }
void ora_set_pair_variant(int32_t v) { g_pair_variant = v; }
double ora_pow_portable(double x, double y) { return ora_pow::pow_portable(x, y); }

// ---- stateless pieces
void ora_gravity_eval(int64_t n, const double* pos, const double* mu, double* out) {
    std::vector<V3> y((size_t)n), a((size_t)n, ZERO3);
    std::vector<double> m(mu, mu + n);
    for (int64_t i = 0; i < n; ++i) y[(size_t)i] = ld3(pos + 3 * i);
    gravity_eval(y, m, a);
    for (int64_t i = 0; i < n; ++i) st3(out + 3 * i, a[(size_t)i]);
}

// OpenMP-free multi-thread friendly variant used only as a timed CPU baseline: same pair formula, but each body
// sums all sources in index order (NOT the reference summation order; not used for parity).
void ora_gravity_eval_rows(int64_t n, const double* pos, const double* mu, int64_t i0, int64_t i1, double* out) {
    for (int64_t i = i0; i < i1; ++i) {
        V3 pi = ld3(pos + 3 * i), acc = ZERO3;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            acc = acc + accel_at(ld3(pos + 3 * j), mu[j], pi);
        }
        st3(out + 3 * i, acc);
    }
}

// TIMING SAMPLE of NewtonianGravity::eval: the reference's inner loop (nbody.rs:27-32) for rows i = i0, i0+stride, ...
// only.  `out` (3n doubles, caller-zeroed) receives those rows' contributions; returns the number of pairs evaluated.
int64_t ora_gravity_eval_row_sample(int64_t n, const double* pos, const double* mu, int64_t i0, int64_t stride, double* out) {
    int64_t pairs = 0;
    for (int64_t i = i0; i < n; i += stride) {
        V3 out_i = ZERO3;
        const V3 pi = ld3(pos + 3 * i);
        for (int64_t j = i + 1; j < n; ++j) {
            V3 ci, cj;
            pair_accel(pi, mu[i], ld3(pos + 3 * j), mu[j], &ci, &cj);
            out_i = out_i + ci;
            st3(out + 3 * j, ld3(out + 3 * j) + cj);
        }
        st3(out + 3 * i, ld3(out + 3 * i) + out_i);
        pairs += n - 1 - i;
    }
    return pairs;
}

// returns number of coefficients (after trim) or -1; coeffs_out has room for 9*3 doubles
int32_t ora_lsq_fit(int32_t degree, const double* ts, const double* xs, int32_t len, double* coeffs_out) {
    V3 x[9];
    for (int i = 0; i < len; ++i) x[i] = ld3(xs + 3 * i);
    Poly p;
    if (!lsq_fit(degree, ts, x, len, &p)) return -1;
    for (int i = 0; i < p.n; ++i) st3(coeffs_out + 3 * i, p.c[i]);
    return p.n;
}

// ---- n-body propagator
void* ora_nbody_create(int64_t n, const double* pos, const double* vel, const double* mu, double t0, double h_signed,
                       int32_t method /*12 = QuinlanTremaine12, 13 = Stormer13, 14 = BlanesMoan14A*/) {
    NBody* s = new NBody();
    s->init(n, pos, vel, mu, t0, h_signed, method);
    return s;
}
// Same integrator over the reference test's compensated Double<DVec3> state (state queries return `.value`).
// Handles from this constructor support only ora_nbody_step / ora_nbody_state / ora_nbody_evals / ora_nbody_destroy_any.
void* ora_nbody_create_compensated(int64_t n, const double* pos, const double* vel, const double* mu, double t0,
                                   double h_signed, int32_t method) {
    NBodyT<DD3>* s = new NBodyT<DD3>();
    s->init(n, pos, vel, mu, t0, h_signed, method);
    return static_cast<INBody*>(s);
}
int32_t ora_inbody_step(void* h, int64_t nsteps) {
    INBody* s = (INBody*)h;
    for (int64_t i = 0; i < nsteps; ++i) {
        int32_t st = s->step();
        if (st) return st;
    }
    return OK;
}
void ora_inbody_state(void* h, double* t, double* pos, double* vel, double* acc) { ((INBody*)h)->get_state(t, pos, vel, acc); }
void ora_inbody_destroy(void* h) { delete (INBody*)h; }
void ora_nbody_destroy(void* h) { delete (NBody*)h; }

// SplineInterpolators::new(delta, [SplineInterpolator{0, period_b, PolyonmialInterpolator::new(pos_b), LSQ{deg_b}}])
// then Integration::with_solout -> new_solution (integration/src/lib.rs:435-445).
void ora_nbody_set_solout(void* h, double delta, const double* periods, const int32_t* degrees, int32_t backward) {
    NBody* s = (NBody*)h;
    s->has_solout = true;
    s->backward = backward != 0;
    s->delta = delta;
    s->interps.assign(s->y.size(), Interp());
    for (size_t b = 0; b < s->y.size(); ++b) {
        Interp& in = s->interps[b];
        in.sample_period = periods[b];
        in.degree = degrees[b];
        in.index = 1;
        for (auto& sm : in.samples) sm = s->y[b];
    }
    s->new_solution();
}
int32_t ora_nbody_step(void* h, int64_t nsteps) {
    NBody* s = (NBody*)h;
    for (int64_t i = 0; i < nsteps; ++i) {
        int32_t st = s->step();
        if (st) return st;
    }
    return OK;
}
void ora_nbody_state(void* h, double* t, double* pos, double* vel, double* acc) { ((NBody*)h)->get_state(t, pos, vel, acc); }
uint64_t ora_nbody_evals(void* h) { return ((NBody*)h)->evals; }
int64_t ora_nbody_spline_len(void* h, int64_t b) { return (int64_t)((NBody*)h)->solution[(size_t)b].polys.size(); }
// coeffs: n_poly * 9 * 3 doubles (zero padded), ncoef: n_poly
void ora_nbody_spline_get(void* h, int64_t b, double* start, double* interval, double* coeffs, int32_t* ncoef) {
    const Spline& sp = ((NBody*)h)->solution[(size_t)b];
    *start = sp.start;
    *interval = sp.interval;
    for (size_t p = 0; p < sp.polys.size(); ++p) {
        if (ncoef) ncoef[p] = sp.polys[p].n;
        if (coeffs)
            for (int c = 0; c < 9; ++c) st3(coeffs + (p * 9 + (size_t)c) * 3, c < sp.polys[p].n ? sp.polys[p].c[c] : ZERO3);
    }
}
// Propagator::take_solution -- nbody.rs:181-189: swap in a fresh solution (caller reads the old one first)
void ora_nbody_take_solution(void* h) { ((NBody*)h)->new_solution(); }
// DirectionalSolout::solution_time / has_reached -- nbody.rs:501-516
double ora_nbody_solution_time(void* h) {
    NBody* s = (NBody*)h;
    bool first = true;
    double best = 0.0;
    for (const Spline& sp : s->solution) {
        double b = s->backward ? sp.start : sp.end();
        if (first || (s->backward ? b > best : b < best)) best = b;
        first = false;
    }
    return best;
}

// The Prediction Planner's call pattern on the oracle (bench.py's CPU arm of the C2 planner-loop figure; the GPU arm is
// tools/planner_loop.cpp through the C ABI): loop { step(); if sync.is_ready() || reached { take_solution(); clone() } }
// (ephemeris_explorer/src/prediction.rs:408-446, Synchronisation::hertz -- :314-332).  Polynomials are folded into one
// FNV-1a hash per body exactly as the tool does.  The first 12 (start-up) steps are taken before the clock starts.
static uint64_t fnv1a(uint64_t h, const void* p, size_t bytes) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < bytes; ++i) {
        h ^= c[i];
        h *= 1099511628211ull;
    }
    return h;
}
int32_t ora_planner_loop(void* hnd, double end_epoch, double tick_seconds, int64_t* steps_out, int64_t* ticks_out,
                         int64_t* polys_out, double* seconds_out, uint64_t* hash_out) {
    NBody* s = (NBody*)hnd;
    using clk = std::chrono::steady_clock;
    for (int i = 0; i < 12; ++i) {
        int32_t st = s->step();
        if (st) return st;
    }
    const size_t n = s->y.size();
    std::vector<uint64_t> hash(n, 1469598103934665603ull);
    int64_t steps = 0, ticks = 0, polys = 0;
    const auto tstart = clk::now();
    auto last = tstart;
    for (;;) {
        int32_t st = s->step();
        if (st) return st;
        ++steps;
        const bool reached = s->backward ? ora_nbody_solution_time(hnd) <= end_epoch : ora_nbody_solution_time(hnd) >= end_epoch;
        const auto now = clk::now();
        if (std::chrono::duration<double>(now - last).count() >= tick_seconds || reached) {
            std::vector<Spline> sol;
            sol.swap(s->solution);  // take_solution: std::mem::replace with a fresh solution
            s->new_solution();
            for (size_t b = 0; b < n; ++b)
                for (const Poly& p : sol[b].polys) {
                    const int32_t nc = p.n;
                    hash[b] = fnv1a(hash[b], &nc, 4);
                    for (int c = 0; c < p.n; ++c) {
                        double v[3] = {p.c[c].x, p.c[c].y, p.c[c].z};
                        hash[b] = fnv1a(hash[b], v, 24);
                    }
                    ++polys;
                }
            NBody snapshot = *s;  // propagator.clone()
            (void)snapshot;
            ++ticks;
            last = clk::now();
            if (reached) break;
        }
    }
    *seconds_out = std::chrono::duration<double>(clk::now() - tstart).count();
    uint64_t all = 1469598103934665603ull;
    for (uint64_t v : hash) all = fnv1a(all, &v, 8);
    *steps_out = steps;
    *ticks_out = ticks;
    *polys_out = polys;
    *hash_out = all;
    return OK;
}

// All-cores CONTEXT figure for bench.py (the reference's loop is serial: nbody.rs:22-38 inside one task, prediction.rs:385):
// the same symmetric pair loop with rows i0, i0+stride, .. dealt round-robin to OpenMP threads, each thread adding into its
// own output array (std::thread), followed by the reduction over threads.  Returns the pairs evaluated; *seconds is the wall time.
int64_t ora_gravity_eval_row_sample_mt(int64_t n, const double* pos, const double* mu, int64_t stride, int32_t threads,
                                       double* seconds, int32_t* threads_used) {
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    std::vector<std::vector<double>> outs((size_t)T, std::vector<double>((size_t)3 * n, 0.0));
    std::vector<int64_t> pairs((size_t)T, 0);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        double* out = outs[(size_t)tid].data();
        int64_t p = 0;
        for (int64_t i = (int64_t)tid * stride; i < n; i += (int64_t)T * stride) {
            V3 out_i = ZERO3;
            const V3 pi = ld3(pos + 3 * i);
            for (int64_t j = i + 1; j < n; ++j) {
                V3 ci, cj;
                pair_accel(pi, mu[i], ld3(pos + 3 * j), mu[j], &ci, &cj);
                out_i = out_i + ci;
                st3(out + 3 * j, ld3(out + 3 * j) + cj);
            }
            st3(out + 3 * i, ld3(out + 3 * i) + out_i);
            p += n - 1 - i;
        }
        pairs[(size_t)tid] = p;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 1; t < T; ++t)
        for (int64_t k = 0; k < 3 * n; ++k) outs[0][(size_t)k] += outs[(size_t)t][(size_t)k];
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (threads_used) *threads_used = T;
    int64_t total = 0;
    for (int64_t p : pairs) total += p;
    return total;
}

// ---- ephemeris table + spline evaluation
void* ora_ephem_create(int64_t nb, const double* mu, const double* start, const double* interval, const int64_t* npoly,
                       const double* coeffs /* sum(npoly) * 9 * 3 */, const int32_t* ncoef /* sum(npoly) */) {
    Ephem* e = new Ephem();
    e->mu.assign(mu, mu + nb);
    e->splines.resize((size_t)nb);
    size_t off = 0;
    for (int64_t b = 0; b < nb; ++b) {
        Spline& sp = e->splines[(size_t)b];
        sp.start = start[b];
        sp.interval = interval[b];
        for (int64_t p = 0; p < npoly[b]; ++p, ++off) {
            Poly q;
            q.n = ncoef[off];
            for (int c = 0; c < q.n; ++c) q.c[c] = ld3(coeffs + (off * 9 + (size_t)c) * 3);
            sp.polys.push_back(q);
        }
    }
    return e;
}
void ora_ephem_destroy(void* e) { delete (Ephem*)e; }
// returns 1 on success, 0 if `at` is outside the spline (None)
int32_t ora_ephem_state_vector(void* e, int64_t b, double at, double* pos, double* vel) {
    V3 p, v;
    if (!((Ephem*)e)->splines[(size_t)b].state_vector(at, &p, &v)) return 0;
    st3(pos, p);
    st3(vel, v);
    return 1;
}
int32_t ora_ephem_position(void* e, int64_t b, double at, double* pos) {
    V3 p;
    if (!((Ephem*)e)->splines[(size_t)b].position(at, &p)) return 0;
    st3(pos, p);
    return 1;
}

// ---- ship
// params = {h_init, h_max, tol_pos, tol_vel, fac_min, fac_max, fac}; burns: start[], end[], acc[3*], ref[]
void* ora_ship_create(void* ephem, double t0, const double* state6, const double* params, uint32_t n_max, int32_t nburns,
                      const double* bstart, const double* bend, const double* bacc, const int32_t* bref) {
    Ship* s = new Ship();
    s->eph = (const Ephem*)ephem;
    s->time = t0;
    s->state = SV{ld3(state6), ld3(state6 + 3)};
    s->h_init = params[0];
    s->h_max = params[1];
    s->tol_pos = params[2];
    s->tol_vel = params[3];
    s->fac_min = params[4];
    s->fac_max = params[5];
    s->fac = params[6];
    s->n_max = n_max;
    std::vector<Burn> burns;
    for (int i = 0; i < nburns; ++i) burns.push_back({bstart[i], bend[i], ld3(bacc + 3 * i), bref[i]});
    s->build_timeline(burns);
    // SpacecraftPropagator::new -- spacecraft.rs:453-477
    s->current_segment = s->segment_idx_at(t0);
    s->bound = s->timeline[s->current_segment].end;
    s->init_integrator();
    s->knots.push_back({s->time, s->state});  // CubicHermiteSplineSolout::new_solution -- spacecraft.rs:656-664
    return s;
}
void ora_ship_destroy(void* h) { delete (Ship*)h; }
// method ids: 0 Verner87 (default), 1 CashKarp45, 2 DormandPrince54, 3 DormandPrince87, 4 Fehlberg45, 5 Tsitouras75,
// 6 Verner98, 7 Fine45 -- before the first step only
int32_t ora_ship_set_method(void* h, int32_t method) {
    if (method < 0 || method >= EE_RK_METHODS) return -1;
    ((Ship*)h)->tab = &EE_RK_TABLE[method];
    return 0;
}
// switch the solution type to SpacecraftSolout's (trajectory + SOI transitions + apsides); soi_radius[n_bodies]
void ora_ship_enable_analytics(void* h, const double* soi_radius) {
    Ship* s = (Ship*)h;
    s->analytics = true;
    s->soi_radius.assign(soi_radius, soi_radius + s->eph->splines.size());
    s->new_analytics();
}
void ora_ship_analytics_counts(void* h, int64_t* n_transitions, int64_t* n_apsides) {
    Ship* s = (Ship*)h;
    *n_transitions = (int64_t)s->transitions.size();
    *n_apsides = (int64_t)s->apsides.size();
}
void ora_ship_analytics(void* h, double* tr_time, int32_t* tr_body, double* ap_time, double* ap_distance, int32_t* ap_body,
                        int32_t* ap_kind) {
    Ship* s = (Ship*)h;
    for (size_t i = 0; i < s->transitions.size(); ++i) {
        tr_time[i] = s->transitions[i].t;
        tr_body[i] = s->transitions[i].body;
    }
    for (size_t i = 0; i < s->apsides.size(); ++i) {
        ap_time[i] = s->apsides[i].t;
        ap_distance[i] = s->apsides[i].distance;
        ap_body[i] = s->apsides[i].body;
        ap_kind[i] = s->apsides[i].kind;
    }
}
int32_t ora_ship_step(void* h, int64_t nsteps) {
    Ship* s = (Ship*)h;
    for (int64_t i = 0; i < nsteps; ++i) {
        int32_t st = s->step();
        if (st) return st;
    }
    return OK;
}
// step until solution.end() >= t_end (IncrementalPropagator::step_to -- ephemeris/src/lib.rs:47-58), at most max_steps
int32_t ora_ship_step_to(void* h, double t_end, int64_t max_steps, int64_t* taken) {
    Ship* s = (Ship*)h;
    int64_t i = 0;
    int32_t st = OK;
    while (i < max_steps && !(s->knots.back().t >= t_end)) {
        st = s->step();
        if (st) break;
        ++i;
    }
    if (taken) *taken = i;
    return st;
}
int64_t ora_ship_knot_count(void* h) { return (int64_t)((Ship*)h)->knots.size(); }
void ora_ship_knots(void* h, double* out7 /* n * (t, px,py,pz, vx,vy,vz) */) {
    Ship* s = (Ship*)h;
    for (size_t i = 0; i < s->knots.size(); ++i) {
        out7[7 * i] = s->knots[i].t;
        st3(out7 + 7 * i + 1, s->knots[i].sv.p);
        st3(out7 + 7 * i + 4, s->knots[i].sv.v);
    }
}
void ora_ship_info(void* h, double* time, double* next_h, uint32_t* n_attempts, uint64_t* rhs_evals) {
    Ship* s = (Ship*)h;
    if (time) *time = s->time;
    if (next_h) *next_h = s->next_h;
    if (n_attempts) *n_attempts = s->n;
    if (rhs_evals) *rhs_evals = s->rhs_evals;
}

// CubicHermite::new + eval -- trajectory.rs:645-690: position on knot segment i at time t
void ora_hermite_eval(const double* k0 /*7*/, const double* k1 /*7*/, double t, double* pos, double* vel) {
    double t0 = k0[0], t1 = k1[0];
    V3 p0 = ld3(k0 + 1), v0 = ld3(k0 + 4), p1 = ld3(k1 + 1), v1 = ld3(k1 + 4);
    V3 a0 = p0, a1 = v0, a2 = ZERO3, a3 = ZERO3;
    double dt = t1 - t0;
    bool same = p0.x == p1.x && p0.y == p1.y && p0.z == p1.z && v0.x == v1.x && v0.y == v1.y && v0.z == v1.z;
    if (!(dt == 0.0 && same)) {
        double r = 1.0 / dt, r2 = r * r, r3 = r * r2;
        V3 dv = p1 - p0;
        a2 = dv * r2 * 3.0 - (v0 * 2.0 + v1) * r;
        a3 = dv * r3 * -2.0 + (v0 + v1) * r2;
    }
    double d = t - t0;
    if (pos) st3(pos, (((a3 * d + a2) * d) + a1) * d + a0);
    if (vel) st3(vel, ((a3 * d * 3.0 + a2 * 2.0) * d) + a1);
}

// CubicHermiteSpline::state_vector -- trajectory.rs:787-795 (binary search; a time equal to a knot returns the knot)
int32_t ora_hermite_spline_state_vector(const double* knots7, int64_t n, double at, double* pos, double* vel) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (knots7[7 * mid] < at)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo < n && knots7[7 * lo] == at) {
        for (int c = 0; c < 3; ++c) {
            pos[c] = knots7[7 * lo + 1 + c];
            vel[c] = knots7[7 * lo + 4 + c];
        }
        return 1;
    }
    if (lo == 0 || lo >= n) return 0;
    ora_hermite_eval(knots7 + 7 * (lo - 1), knots7 + 7 * lo, at, pos, vel);
    return 1;
}

// RelativeTrajectory::state_vector -- trajectory.rs:326-334: reference first, then the trajectory, component-wise difference.
// The trajectory is body `body` of the ephemeris, or the Hermite spline `knots7` when given; reference = body or -1.
int32_t ora_relative_state_vector(void* e, const double* knots7, int64_t n_knots, int32_t body, int32_t reference, double at,
                                  double* pos, double* vel) {
    const Ephem* E = (const Ephem*)e;
    V3 rp = ZERO3, rv = ZERO3;
    if (reference >= 0 && !E->splines[(size_t)reference].state_vector(at, &rp, &rv)) return 0;
    V3 p, v;
    if (knots7) {
        double pp[3], vv[3];
        if (!ora_hermite_spline_state_vector(knots7, n_knots, at, pp, vv)) return 0;
        p = ld3(pp);
        v = ld3(vv);
    } else if (!E->splines[(size_t)body].state_vector(at, &p, &v)) {
        return 0;
    }
    st3(pos, p - rp);
    st3(vel, v - rv);
    return 1;
}

}  // extern "C"
