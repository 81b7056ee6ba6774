// ee_solout.cu -- dense output of the n-body propagator and the device-resident spline ephemeris.
//
// Reference functions restated here (file:line under the reference tree):
//   SplineInterpolators::solout / solout_with / new_solution   ephemeris/src/propagators/nbody.rs:372-489   (a10)
//   PolyonmialInterpolator<9, V>                                ephemeris/src/propagators/nbody.rs:243-307
//   LeastSquaresFit::interpolate                                ephemeris_explorer/src/dynamics/celestial.rs:24-136 (a11)
//   UniformSpline::{position, state_vector, get_polynomial}     ephemeris/src/trajectory.rs:449-471, :552-617 (a12)
//   Polynomial::{eval, eval_and_deriv, trim}                    ephemeris/src/trajectory.rs:360-410
//
// The sampling schedule is pure f64 bookkeeping and runs on the host exactly as the reference does it
// (`last_sample_time += delta; if last_sample_time == sample_period`); sampled positions never leave the device:
// a sample kernel appends them to per-body buffers, a fit kernel turns every complete group of 9 samples into a
// polynomial (one thread per fit, reference operation order, no FMA), and the polynomials accumulate in a device
// pool until take_solution.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

#include "ee_engine.h"

namespace ee {

// ---------------------------------------------------------------------------------------------------------
__global__ void k_sample(int64_t n, int64_t step, const int64_t* __restrict__ stride, const int64_t* __restrict__ off,
                         const int64_t* __restrict__ qbase, const double4* __restrict__ y, double* __restrict__ samples) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int64_t s = stride[b];
    if (s <= 0 || step % s != 0) return;
    const int64_t idx = off[b] + (step / s - qbase[b]);
    const double4 p = y[b];
    samples[3 * idx] = p.x;
    samples[3 * idx + 1] = p.y;
    samples[3 * idx + 2] = p.z;
}

struct Ts9 {
    double t[9];
};

// LeastSquaresFit::interpolate, lane-wise on DVec3.  The orthogonal polynomials p_k depend on the abscissae only,
// so their three lanes are identical in the reference; they are kept as scalars here (same bits).
__device__ int lsq_fit_dev(int degree_req, const double* ts, const D3* xs, D3* out) {
    D3 d0 = {0.0, 0.0, 0.0};
    double gamma0 = 0.0, b0 = 0.0;
    for (int i = 0; i < 9; ++i) {
        d0 = xadd3(d0, xs[i]);
        gamma0 = xadd(gamma0, 1.0);
        b0 = xadd(b0, ts[i]);
    }
    if (gamma0 == 0.0) return -1;
    const int degree = min(degree_req, 8);
    b0 = xdiv(b0, gamma0);
    d0 = xdiv3(d0, gamma0);
    if (degree == 0) {
        out[0] = d0;
        return 1;
    }
    double P[2][9];
    for (int i = 0; i <= degree; ++i) {
        out[i] = d3(0.0, 0.0, 0.0);
        P[0][i] = 0.0;
        P[1][i] = 0.0;
    }
    int km1 = 0, kk = 1;  // P[km1] = p_{k-1}, P[kk] = p_k
    out[0] = d0;
    P[kk][0] = 1.0;
    double gamma_k = gamma0, b_k = b0, minus_c_k = 0.0;
    int kp1 = 1;
    for (;;) {
        for (int i = 0; i < kp1; ++i) P[km1][i] = xsub(xmul(minus_c_k, P[km1][i]), xmul(b_k, P[kk][i]));
        for (int im1 = 0; im1 < kp1; ++im1) P[km1][im1 + 1] = xadd(P[km1][im1 + 1], P[kk][im1]);
        D3 d = {0.0, 0.0, 0.0};
        double gamma = 0.0, b = 0.0;
        for (int s = 0; s < 9; ++s) {
            double px = 0.0;
            for (int i = kp1; i >= 0; --i) px = xadd(xmul(px, ts[s]), P[km1][i]);
            d = xadd3(d, xmul3(xs[s], px));
            const double w2 = xmul(px, px);
            gamma = xadd(gamma, w2);
            b = xadd(b, xmul(ts[s], w2));
        }
        if (gamma == 0.0) break;
        d = xdiv3(d, gamma);
        for (int i = 0; i <= kp1; ++i) out[i] = xadd3(out[i], xmul3(d, P[km1][i]));
        if (kp1 == degree) break;
        b = xdiv(b, gamma);
        kp1 += 1;
        b_k = b;
        minus_c_k = -xdiv(gamma, gamma_k);
        gamma_k = gamma;
        const int tmp = km1;
        km1 = kk;
        kk = tmp;
    }
    int nc = degree + 1;
    while (nc > 0 && out[nc - 1].x == 0.0 && out[nc - 1].y == 0.0 && out[nc - 1].z == 0.0) nc--;  // Polynomial::trim
    return nc;
}

// one thread per fit; src[f] = index (in DVec3 units) of the first of 9 consecutive samples
__global__ void k_fit(int64_t nfits, const int64_t* __restrict__ src, const int32_t* __restrict__ deg, Ts9 ts,
                      const double* __restrict__ samples, double* __restrict__ coef, int32_t* __restrict__ ncoef) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfits) return;
    D3 xs[9], out[9];
    const double* sp = samples + 3 * src[f];
    for (int i = 0; i < 9; ++i) xs[i] = d3(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]);
    for (int i = 0; i < 9; ++i) out[i] = d3(0.0, 0.0, 0.0);
    const int nc = lsq_fit_dev(deg[f], ts.t, xs, out);
    double* c = coef + 27 * f;
    for (int i = 0; i < 9; ++i) {
        const bool live = nc > 0 && i < nc;
        c[3 * i] = live ? out[i].x : 0.0;
        c[3 * i + 1] = live ? out[i].y : 0.0;
        c[3 * i + 2] = live ? out[i].z : 0.0;
    }
    ncoef[f] = nc;
}

// after fitting g groups, samples [8g, held) slide to the front of the body's buffer (PolyonmialInterpolator::finish)
__global__ void k_compact(int64_t n, const int64_t* __restrict__ off, const int64_t* __restrict__ shift,
                          const int64_t* __restrict__ rem, double* __restrict__ samples) {
    const int64_t b = blockIdx.x;
    if (b >= n) return;
    const int64_t sh = shift[b];
    if (sh == 0) return;
    // rem <= 8 <= sh: source and destination never overlap
    for (int64_t i = threadIdx.x; i < 3 * rem[b]; i += blockDim.x)
        samples[3 * off[b] + i] = samples[3 * (off[b] + sh) + i];
}

// ---------------------------------------------------------------------------------------------------------
int64_t sampling_stride(double delta, double period) {
    // exact emulation of `last_sample_time += delta; last_sample_time == sample_period` (nbody.rs:389-391)
    double acc = 0.0;
    for (int64_t c = 1; c <= (int64_t)1 << 24; ++c) {
        acc = acc + delta;
        if (acc == period) return c;
        if (std::fabs(acc) > std::fabs(period) || std::isnan(acc)) return 0;
    }
    return 0;
}

Solout::Solout(NBodyEngine& e, double delta_, const double* periods, const int32_t* degrees) {
    n = e.n;
    delta = delta_;
    backward = e.h < 0.0;
    period.assign(periods, periods + n);
    degree.assign(degrees, degrees + n);
    last_sample_time.assign((size_t)n, 0.0);
    stride.resize((size_t)n);
    since.assign((size_t)n, 0);
    double density = 0.0;
    for (int64_t b = 0; b < n; ++b) {
        EE_REQUIRE(degree[(size_t)b] >= 0, "negative polynomial degree");
        stride[(size_t)b] = sampling_stride(delta, period[(size_t)b]);
        if (stride[(size_t)b] > 0) density += 1.0 / (double)stride[(size_t)b];
    }
    int64_t batch = 8192;
    if (density > 0.0) batch = (int64_t)std::max(64.0, std::min(8192.0, 4.0e6 / density));
    off.resize((size_t)n);
    cap.resize((size_t)n);
    held.assign((size_t)n, 1);  // PolyonmialInterpolator::new: index = 1, sample 0 = initial position
    qbase.assign((size_t)n, 0);
    int64_t total = 0;
    for (int64_t b = 0; b < n; ++b) {
        off[(size_t)b] = total;
        cap[(size_t)b] = stride[(size_t)b] > 0 ? batch / stride[(size_t)b] + 10 : 1;
        total += cap[(size_t)b];
    }
    samples.alloc((size_t)total * 3);
    d_stride.alloc((size_t)n);
    d_off.alloc((size_t)n);
    d_qbase.alloc((size_t)n);
    segs.assign((size_t)n, {});
    done.assign((size_t)n, 0);
    // sample 0 of every body = current position: run the sample kernel with "every body samples now"
    {
        std::vector<int64_t> ones((size_t)n, 1), zeros((size_t)n, 0);
        EE_CUDA(cudaMemcpyAsync(d_stride.p, ones.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
        EE_CUDA(cudaMemcpyAsync(d_off.p, off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
        EE_CUDA(cudaMemcpyAsync(d_qbase.p, zeros.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
        const int B = 128;
        k_sample<<<(unsigned)((n + B - 1) / B), B, 0, e.stream>>>(n, 0, d_stride.p, d_off.p, d_qbase.p, e.positions_dev(),
                                                                   samples.p);
        EE_CUDA(cudaGetLastError());
        count_launch();
        EE_CUDA(cudaStreamSynchronize(e.stream));
    }
    dirty_meta = true;
    new_solution(e);
}

void Solout::upload_meta(NBodyEngine& e) {
    if (!dirty_meta) return;
    EE_CUDA(cudaMemcpyAsync(d_stride.p, stride.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
    EE_CUDA(cudaMemcpyAsync(d_off.p, off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
    EE_CUDA(cudaMemcpyAsync(d_qbase.p, qbase.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
    dirty_meta = false;
}

double Solout::interp_time(int64_t b) const {
    double lst = 0.0;
    for (int64_t i = 0; i < since[(size_t)b]; ++i) lst = lst + delta;
    const int64_t len = held[(size_t)b] - 8 * ((held[(size_t)b] - 1) / 8);  // samples the interpolator holds (1..8)
    return lst + period[(size_t)b] * (double)(len > 0 ? len - 1 : 0);
}

void Solout::new_solution(const NBodyEngine& e) {
    sol_start.resize((size_t)n);
    sol_interval.resize((size_t)n);
    for (int64_t b = 0; b < n; ++b) {
        const double it = -interp_time(b);
        sol_start[(size_t)b] = backward ? e.t - it : e.t + it;  // D::offset(problem.time, -interp.time())
        sol_interval[(size_t)b] = period[(size_t)b] * 8.0;
    }
}

double Solout::bound(int64_t b) const {
    const int64_t np = n_poly(b);
    if (!backward) return sol_start[(size_t)b] + sol_interval[(size_t)b] * (double)np;  // end() = start + span()
    double s = sol_start[(size_t)b];
    for (int64_t i = 0; i < np; ++i) s = s - sol_interval[(size_t)b];                    // push_front: start -= interval
    return s;
}

double Solout::solution_time() const {
    double best = bound(0);
    for (int64_t b = 1; b < n; ++b) {
        const double v = bound(b);
        if (backward ? v > best : v < best) best = v;
    }
    return best;
}

bool Solout::has_reached(double epoch) const {
    for (int64_t b = 0; b < n; ++b) {
        const double v = bound(b);
        if (backward ? !(v <= epoch) : !(v >= epoch)) return false;
    }
    return true;
}

// Each body's bound only moves when it completes a polynomial (every 8*stride steps), and once past `epoch` it
// stays past it, so step_to's stopping step is the maximum over bodies of the step at which that body arrives.
int64_t Solout::steps_until(double epoch) const {
    int64_t worst = 0;
    for (int64_t b = 0; b < n; ++b) {
        const int64_t s = stride[(size_t)b];
        const int64_t have = n_poly(b);
        auto reached = [&](int64_t np) {
            double v;
            if (!backward) {
                v = sol_start[(size_t)b] + sol_interval[(size_t)b] * (double)np;
                return v >= epoch;
            }
            v = sol_start[(size_t)b];
            for (int64_t i = 0; i < np; ++i) v = v - sol_interval[(size_t)b];
            return v <= epoch;
        };
        if (reached(have)) continue;
        if (s <= 0) return (int64_t)1 << 62;  // this body never samples: the reference would loop forever too
        const double span = std::fabs(epoch - sol_start[(size_t)b]) / sol_interval[(size_t)b];
        int64_t np = std::max<int64_t>(have + 1, (int64_t)span - 2);
        while (!reached(np)) ++np;
        while (np - 1 > have && reached(np - 1)) --np;
        // progress inside the current polynomial: (held-1) % 8 samples done, `since` steps towards the next one
        const int64_t inside = ((held[(size_t)b] - 1) % 8) * s + since[(size_t)b];
        worst = std::max(worst, (np - have) * 8 * s - inside);
    }
    return worst;
}

int32_t Solout::after_step(NBodyEngine& e) {
    steps_done += 1;
    bool any = false, full = false;
    for (int64_t b = 0; b < n; ++b) {
        const int64_t s = stride[(size_t)b];
        if (s <= 0) continue;
        since[(size_t)b] += 1;
        if (since[(size_t)b] == s) {
            since[(size_t)b] = 0;
            held[(size_t)b] += 1;
            any = true;
            if (held[(size_t)b] >= cap[(size_t)b]) full = true;
        }
    }
    if (any) {
        upload_meta(e);
        const int B = 128;
        k_sample<<<(unsigned)((n + B - 1) / B), B, 0, e.stream>>>(n, steps_done, d_stride.p, d_off.p, d_qbase.p,
                                                                   e.positions_dev(), samples.p);
        EE_CUDA(cudaGetLastError());
        count_launch();
    }
    if (full) flush(e);
    return EE_OK;
}

int64_t Solout::room() const {
    int64_t r = (int64_t)1 << 40;
    for (int64_t b = 0; b < n; ++b) {
        const int64_t s = stride[(size_t)b];
        if (s <= 0) continue;
        r = std::min(r, (cap[(size_t)b] - held[(size_t)b]) * s - since[(size_t)b]);
    }
    return std::max<int64_t>(0, r);
}

void Solout::begin_batch(NBodyEngine& e) { upload_meta(e); }

void Solout::advance_host(int64_t k) {
    steps_done += k;
    for (int64_t b = 0; b < n; ++b) {
        const int64_t s = stride[(size_t)b];
        if (s <= 0) continue;
        const int64_t tot = since[(size_t)b] + k;
        held[(size_t)b] += tot / s;
        since[(size_t)b] = tot % s;
    }
}

void Solout::grow_pool(NBodyEngine& e, int64_t need) {
    if (pool_len + need <= pool_cap) return;
    int64_t ncap = std::max<int64_t>(1024, pool_cap * 2);
    while (ncap < pool_len + need) ncap *= 2;
    DBuf<double> np((size_t)ncap * 27);
    DBuf<int32_t> nn((size_t)ncap);
    if (pool_len) {
        EE_CUDA(cudaMemcpyAsync(np.p, pool.p, (size_t)pool_len * 27 * 8, cudaMemcpyDeviceToDevice, e.stream));
        EE_CUDA(cudaMemcpyAsync(nn.p, pool_nc.p, (size_t)pool_len * 4, cudaMemcpyDeviceToDevice, e.stream));
        EE_CUDA(cudaStreamSynchronize(e.stream));
    }
    pool = std::move(np);
    pool_nc = std::move(nn);
    pool_cap = ncap;
}

void Solout::flush(NBodyEngine& e) {
    e.flush_pending();  // samples of run-ahead steps must be on the device before they are fitted
    std::vector<int64_t> src, shift((size_t)n, 0), rem((size_t)n, 0);
    std::vector<int32_t> deg;
    bool any = false;
    for (int64_t b = 0; b < n; ++b) {
        const int64_t g = (held[(size_t)b] - 1) / 8;
        if (g == 0) continue;
        any = true;
        segs[(size_t)b].push_back({pool_len + (int64_t)src.size(), g});
        for (int64_t k = 0; k < g; ++k) {
            src.push_back(off[(size_t)b] + 8 * k);
            deg.push_back(degree[(size_t)b]);
        }
        shift[(size_t)b] = 8 * g;
        rem[(size_t)b] = held[(size_t)b] - 8 * g;
        held[(size_t)b] -= 8 * g;
        qbase[(size_t)b] += 8 * g;
        done[(size_t)b] += g;
    }
    if (!any) return;
    const int64_t nf = (int64_t)src.size();
    upload_meta(e);  // k_compact reads d_off: a fresh clone that is flushed before it ever stepped has not uploaded it yet
    grow_pool(e, nf);
    DBuf<int64_t> d_src((size_t)nf), d_shift((size_t)n), d_rem((size_t)n);
    DBuf<int32_t> d_deg((size_t)nf);
    EE_CUDA(cudaMemcpyAsync(d_src.p, src.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, e.stream));
    EE_CUDA(cudaMemcpyAsync(d_deg.p, deg.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, e.stream));
    EE_CUDA(cudaMemcpyAsync(d_shift.p, shift.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
    EE_CUDA(cudaMemcpyAsync(d_rem.p, rem.data(), (size_t)n * 8, cudaMemcpyHostToDevice, e.stream));
    Ts9 ts;
    for (int i = 0; i < 9; ++i) ts.t[i] = backward ? 1.0 - (double)i / 8.0 : (double)i / 8.0;  // nbody.rs:422-442
    const int B = 64;
    k_fit<<<(unsigned)((nf + B - 1) / B), B, 0, e.stream>>>(nf, d_src.p, d_deg.p, ts, samples.p, pool.p + (size_t)pool_len * 27,
                                                            pool_nc.p + pool_len);
    EE_CUDA(cudaGetLastError());
    k_compact<<<(unsigned)n, 32, 0, e.stream>>>(n, d_off.p, d_shift.p, d_rem.p, samples.p);
    EE_CUDA(cudaGetLastError());
    count_launch(2);
    pool_len += nf;
    dirty_meta = true;
    EE_CUDA(cudaStreamSynchronize(e.stream));  // temporaries die here
}

void Solout::take(NBodyEngine& e, HostSolution& out) {
    flush(e);
    e.sync();
    std::vector<double> hc((size_t)pool_len * 27);
    std::vector<int32_t> hn((size_t)pool_len);
    if (pool_len) {
        EE_CUDA(cudaMemcpy(hc.data(), pool.p, hc.size() * 8, cudaMemcpyDeviceToHost));
        EE_CUDA(cudaMemcpy(hn.data(), pool_nc.p, hn.size() * 4, cudaMemcpyDeviceToHost));
    }
    out.start.resize((size_t)n);
    out.interval.resize((size_t)n);
    out.n_poly.resize((size_t)n);
    out.coeffs.clear();
    out.n_coef.clear();
    for (int64_t b = 0; b < n; ++b) {
        std::vector<int64_t> idx;
        for (auto& sg : segs[(size_t)b])
            for (int64_t k = 0; k < sg.second; ++k) idx.push_back(sg.first + k);
        if (backward) std::reverse(idx.begin(), idx.end());  // push_front: newest polynomial first
        out.n_poly[(size_t)b] = (int64_t)idx.size();
        out.interval[(size_t)b] = sol_interval[(size_t)b];
        out.start[(size_t)b] = backward ? bound(b) : sol_start[(size_t)b];  // push_front moved the start back
        for (int64_t i : idx) {
            out.coeffs.insert(out.coeffs.end(), hc.begin() + (size_t)i * 27, hc.begin() + (size_t)(i + 1) * 27);
            out.n_coef.push_back(hn[(size_t)i]);
        }
    }
    pool_len = 0;
    for (auto& s : segs) s.clear();
    std::fill(done.begin(), done.end(), 0);
    new_solution(e);
}

Solout* Solout::clone(NBodyEngine& owner) const {
    EE_REQUIRE(owner.pending == 0, "internal: clone with run-ahead steps pending");
    Solout* c = new Solout();
    c->n = n;
    c->delta = delta;
    c->backward = backward;
    c->period = period;
    c->degree = degree;
    c->last_sample_time = last_sample_time;
    c->stride = stride;
    c->since = since;
    c->steps_done = steps_done;
    c->off = off;
    c->cap = cap;
    c->held = held;
    c->qbase = qbase;
    c->segs = segs;
    c->done = done;
    c->sol_start = sol_start;
    c->sol_interval = sol_interval;
    c->pool_cap = pool_cap;
    c->pool_len = pool_len;
    c->samples.alloc(samples.n);
    c->d_stride.alloc((size_t)n);
    c->d_off.alloc((size_t)n);
    c->d_qbase.alloc((size_t)n);
    EE_CUDA(cudaMemcpyAsync(c->samples.p, samples.p, samples.bytes(), cudaMemcpyDeviceToDevice, owner.stream));
    if (pool_cap) {
        c->pool.alloc(pool.n);
        c->pool_nc.alloc(pool_nc.n);
        EE_CUDA(cudaMemcpyAsync(c->pool.p, pool.p, (size_t)pool_len * 27 * 8, cudaMemcpyDeviceToDevice, owner.stream));
        EE_CUDA(cudaMemcpyAsync(c->pool_nc.p, pool_nc.p, (size_t)pool_len * 4, cudaMemcpyDeviceToDevice, owner.stream));
    }
    c->dirty_meta = true;
    return c;
}

// ---- serialisation (ee_nbody_snapshot / ee_nbody_restore with a solout attached) -----------------------------------
namespace {
struct BlobWriter {
    unsigned char* p;
    template <class T>
    void pod(const T& v) {
        std::memcpy(p, &v, sizeof(T));
        p += sizeof(T);
    }
    template <class T>
    void vec(const std::vector<T>& v) {
        pod<int64_t>((int64_t)v.size());
        if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
        p += v.size() * sizeof(T);
    }
};
struct BlobReader {
    const unsigned char* p;
    const unsigned char* end;
    void need(size_t k) const {
        if ((size_t)(end - p) < k) throw Error(EE_ERR_INVALID, "snapshot blob is truncated");
    }
    template <class T>
    T pod() {
        need(sizeof(T));
        T v;
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    template <class T>
    void vec(std::vector<T>& v) {
        const int64_t k = pod<int64_t>();
        if (k < 0) throw Error(EE_ERR_INVALID, "corrupt snapshot blob");
        need((size_t)k * sizeof(T));
        v.resize((size_t)k);
        if (k) std::memcpy(v.data(), p, (size_t)k * sizeof(T));
        p += (size_t)k * sizeof(T);
    }
};
template <class T>
int64_t vec_bytes(const std::vector<T>& v) {
    return 8 + (int64_t)(v.size() * sizeof(T));
}
}  // namespace

int64_t Solout::blob_bytes() const {
    int64_t b = 8 + 8 + 8 + 8 + 8 + 8;  // n, delta, backward, steps_done, pool_len, samples count
    b += vec_bytes(period) + vec_bytes(degree) + vec_bytes(last_sample_time) + vec_bytes(stride) + vec_bytes(since);
    b += vec_bytes(off) + vec_bytes(cap) + vec_bytes(held) + vec_bytes(qbase) + vec_bytes(done);
    b += vec_bytes(sol_start) + vec_bytes(sol_interval);
    for (const auto& sg : segs) b += 8 + (int64_t)sg.size() * 16;
    b += (int64_t)samples.bytes() + pool_len * 27 * 8 + pool_len * 4;
    return b;
}

void Solout::save(NBodyEngine& e, unsigned char* out) {
    e.flush_pending();
    BlobWriter w{out};
    w.pod<int64_t>(n);
    w.pod<double>(delta);
    w.pod<int64_t>(backward ? 1 : 0);
    w.pod<int64_t>(steps_done);
    w.pod<int64_t>(pool_len);
    w.pod<int64_t>((int64_t)samples.n);
    w.vec(period);
    w.vec(degree);
    w.vec(last_sample_time);
    w.vec(stride);
    w.vec(since);
    w.vec(off);
    w.vec(cap);
    w.vec(held);
    w.vec(qbase);
    w.vec(done);
    w.vec(sol_start);
    w.vec(sol_interval);
    for (const auto& sg : segs) {
        w.pod<int64_t>((int64_t)sg.size());
        for (const auto& pr : sg) {
            w.pod<int64_t>(pr.first);
            w.pod<int64_t>(pr.second);
        }
    }
    EE_CUDA(cudaMemcpyAsync(w.p, samples.p, samples.bytes(), cudaMemcpyDeviceToHost, e.stream));
    w.p += samples.bytes();
    if (pool_len) {
        EE_CUDA(cudaMemcpyAsync(w.p, pool.p, (size_t)pool_len * 27 * 8, cudaMemcpyDeviceToHost, e.stream));
        w.p += (size_t)pool_len * 27 * 8;
        EE_CUDA(cudaMemcpyAsync(w.p, pool_nc.p, (size_t)pool_len * 4, cudaMemcpyDeviceToHost, e.stream));
        w.p += (size_t)pool_len * 4;
    }
    EE_CUDA(cudaStreamSynchronize(e.stream));
}

Solout* Solout::load(NBodyEngine& e, const unsigned char* in, int64_t bytes) {
    std::unique_ptr<Solout> c(new Solout());
    BlobReader r{in, in + bytes};
    c->n = r.pod<int64_t>();
    EE_REQUIRE(c->n == e.n, "snapshot solout does not match this handle");
    c->delta = r.pod<double>();
    c->backward = r.pod<int64_t>() != 0;
    c->steps_done = r.pod<int64_t>();
    c->pool_len = r.pod<int64_t>();
    const int64_t nsamp = r.pod<int64_t>();
    EE_REQUIRE(c->pool_len >= 0 && nsamp >= 0, "corrupt snapshot blob");
    r.vec(c->period);
    r.vec(c->degree);
    r.vec(c->last_sample_time);
    r.vec(c->stride);
    r.vec(c->since);
    r.vec(c->off);
    r.vec(c->cap);
    r.vec(c->held);
    r.vec(c->qbase);
    r.vec(c->done);
    r.vec(c->sol_start);
    r.vec(c->sol_interval);
    const size_t nn = (size_t)c->n;
    EE_REQUIRE(c->period.size() == nn && c->degree.size() == nn && c->stride.size() == nn && c->since.size() == nn &&
                   c->off.size() == nn && c->cap.size() == nn && c->held.size() == nn && c->qbase.size() == nn &&
                   c->done.size() == nn && c->sol_start.size() == nn && c->sol_interval.size() == nn,
               "corrupt snapshot blob");
    c->segs.assign(nn, {});
    for (auto& sg : c->segs) {
        const int64_t k = r.pod<int64_t>();
        EE_REQUIRE(k >= 0 && k <= c->pool_len, "corrupt snapshot blob");
        for (int64_t i = 0; i < k; ++i) {
            const int64_t a = r.pod<int64_t>(), b = r.pod<int64_t>();
            sg.push_back({a, b});
        }
    }
    c->samples.alloc((size_t)nsamp);
    r.need(c->samples.bytes());
    EE_CUDA(cudaMemcpyAsync(c->samples.p, r.p, c->samples.bytes(), cudaMemcpyHostToDevice, e.stream));
    r.p += c->samples.bytes();
    c->pool_cap = std::max<int64_t>(1024, c->pool_len);
    c->pool.alloc((size_t)c->pool_cap * 27);
    c->pool_nc.alloc((size_t)c->pool_cap);
    if (c->pool_len) {
        r.need((size_t)c->pool_len * (27 * 8 + 4));
        EE_CUDA(cudaMemcpyAsync(c->pool.p, r.p, (size_t)c->pool_len * 27 * 8, cudaMemcpyHostToDevice, e.stream));
        r.p += (size_t)c->pool_len * 27 * 8;
        EE_CUDA(cudaMemcpyAsync(c->pool_nc.p, r.p, (size_t)c->pool_len * 4, cudaMemcpyHostToDevice, e.stream));
        r.p += (size_t)c->pool_len * 4;
    }
    c->d_stride.alloc(nn);
    c->d_off.alloc(nn);
    c->d_qbase.alloc(nn);
    c->dirty_meta = true;
    EE_CUDA(cudaStreamSynchronize(e.stream));
    return c.release();
}

// stand-alone batched LeastSquaresFit::interpolate
void lsq_fit_batch(int64_t n_fits, const int32_t* degrees, const double* ts9, const double* samples, int device,
                   double* coeffs, int32_t* n_coef) {
    EE_REQUIRE(n_fits >= 0, "negative count");
    if (n_fits == 0) return;
    int ndev = 0;
    EE_CUDA(cudaGetDeviceCount(&ndev));
    EE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU path)");
    EE_CUDA(cudaSetDevice(device));
    DBuf<double> d_s((size_t)n_fits * 27), d_c((size_t)n_fits * 27);
    DBuf<int64_t> d_src((size_t)n_fits);
    DBuf<int32_t> d_deg((size_t)n_fits), d_nc((size_t)n_fits);
    std::vector<int64_t> src((size_t)n_fits);
    for (int64_t f = 0; f < n_fits; ++f) src[(size_t)f] = 9 * f;
    EE_CUDA(cudaMemcpy(d_s.p, samples, d_s.bytes(), cudaMemcpyHostToDevice));
    EE_CUDA(cudaMemcpy(d_src.p, src.data(), d_src.bytes(), cudaMemcpyHostToDevice));
    EE_CUDA(cudaMemcpy(d_deg.p, degrees, d_deg.bytes(), cudaMemcpyHostToDevice));
    Ts9 ts;
    for (int i = 0; i < 9; ++i) ts.t[i] = ts9[i];
    const int B = 64;
    k_fit<<<(unsigned)((n_fits + B - 1) / B), B>>>(n_fits, d_src.p, d_deg.p, ts, d_s.p, d_c.p, d_nc.p);
    EE_CUDA(cudaGetLastError());
    count_launch();
    EE_CUDA(cudaMemcpy(coeffs, d_c.p, d_c.bytes(), cudaMemcpyDeviceToHost));
    EE_CUDA(cudaMemcpy(n_coef, d_nc.p, d_nc.bytes(), cudaMemcpyDeviceToHost));
}

// ---------------------------------------------------------------------------------------------------------
// Ephemeris table
Ephem::Ephem(int64_t nb_, const double* mus, const double* start_, const double* interval_, const int64_t* n_poly_,
             const double* coeffs, const int32_t* n_coef, int device_)
    : device(device_), nb(nb_) {
    EE_REQUIRE(nb >= 1, "ephemeris needs at least one body");
    int ndev = 0;
    EE_CUDA(cudaGetDeviceCount(&ndev));
    EE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU path)");
    EE_CUDA(cudaSetDevice(device));
    mu.assign(mus, mus + nb);
    start.assign(start_, start_ + nb);
    interval.assign(interval_, interval_ + nb);
    n_poly.assign(n_poly_, n_poly_ + nb);
    first.resize((size_t)nb);
    total = 0;
    for (int64_t b = 0; b < nb; ++b) {
        first[(size_t)b] = total;
        total += n_poly[(size_t)b];
    }
    coef.alloc((size_t)std::max<int64_t>(1, total) * 27);
    ncoef.alloc((size_t)std::max<int64_t>(1, total));
    d_mu.alloc((size_t)nb);
    d_start.alloc((size_t)nb);
    d_interval.alloc((size_t)nb);
    d_npoly.alloc((size_t)nb);
    d_first.alloc((size_t)nb);
    if (total) {
        EE_CUDA(cudaMemcpy(coef.p, coeffs, (size_t)total * 27 * 8, cudaMemcpyHostToDevice));
        EE_CUDA(cudaMemcpy(ncoef.p, n_coef, (size_t)total * 4, cudaMemcpyHostToDevice));
    }
    EE_CUDA(cudaMemcpy(d_mu.p, mu.data(), (size_t)nb * 8, cudaMemcpyHostToDevice));
    EE_CUDA(cudaMemcpy(d_start.p, start.data(), (size_t)nb * 8, cudaMemcpyHostToDevice));
    EE_CUDA(cudaMemcpy(d_interval.p, interval.data(), (size_t)nb * 8, cudaMemcpyHostToDevice));
    EE_CUDA(cudaMemcpy(d_npoly.p, n_poly.data(), (size_t)nb * 8, cudaMemcpyHostToDevice));
    EE_CUDA(cudaMemcpy(d_first.p, first.data(), (size_t)nb * 8, cudaMemcpyHostToDevice));
}

}  // namespace ee
