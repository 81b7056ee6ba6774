// ee_kernels.cuh -- device code of the n-body path: all-pairs acceleration (parity + throughput flavours) and the
// integrator epilogues fused behind it.
//
// Reference functions restated here (file:line under the reference tree):
//   NewtonianGravity::eval                 ephemeris/src/propagators/nbody.rs:16-39            (a1/a2)
//   ELM2::advance (QT12 / Stormer13)       integration/src/multistep/second_order/mod.rs:91-131 (a3)
//   Cowell::update_velocity                integration/src/multistep/second_order/cowell.rs:19-53 (a4)
//   SRKN<BlanesMoan6B>::advance            integration/src/runge_kutta/nystrom/symplectic.rs:70-102 (a8)
#pragma once
#include "ee_common.cuh"

namespace ee {

constexpr int kMaxOrder = 13;

// Arguments of one linear-multistep step.  Steps are numbered absolutely; ring slot of step s is s % R.
// slot[j] = ring slot holding (y, a) of step (m+1-j), j = 0..order-1, where m+1 is the step being completed.
struct QtArgs {
    int order;
    int slot[kMaxOrder];
    int slot_next;  // ring slot receiving the predicted y_{m+2}
    double nalpha[kMaxOrder];
    double beta[kMaxOrder];
    double cow[kMaxOrder];
    double f;  // (h*h) * (1/beta_D)          second_order/mod.rs:120
    double g;  // h * (1/cowell_beta_D)       cowell.rs:51
    double h;
};

struct KdArgs {  // one SRKN kick-drift stage
    double hb;   // h_sub * B[s]
    double ha;   // h_sub * A[s]
};

enum EpKind : int { EP_STORE = 0, EP_KD = 1, EP_QT = 2 };

struct EpArgs {
    int kind;
    int64_t n;          // bodies (array stride)
    double* a_out;      // [3][n]   acceleration of the evaluated positions
    // EP_KD
    KdArgs kd;
    const double4* y_in;  // positions the acceleration was evaluated at (x,y,z,mu)
    double4* y_out;       // drifted positions
    double* dy;           // [3][n] velocities, updated in place
    // EP_QT
    QtArgs qt;
    double4* ry;  // ring of positions  [R][n]
    double* ra;   // ring of accelerations [R][3][n]
};

__device__ __forceinline__ D3 ld_a(const double* a, int64_t n, int64_t k) { return {a[k], a[n + k], a[2 * n + k]}; }
__device__ __forceinline__ void st_a(double* a, int64_t n, int64_t k, D3 v) {
    a[k] = v.x;
    a[n + k] = v.y;
    a[2 * n + k] = v.z;
}

template <bool EXACT>
__device__ __forceinline__ double madd(double acc, double v, double c) {  // acc + v*c
    if (EXACT) return xadd(acc, xmul(v, c));
    return fma(v, c, acc);
}
template <bool EXACT>
__device__ __forceinline__ D3 madd3(D3 acc, D3 v, double c) {
    return {madd<EXACT>(acc.x, v.x, c), madd<EXACT>(acc.y, v.y, c), madd<EXACT>(acc.z, v.z, c)};
}

// y_{s+1} = S1 + S2 * f with S1 = sum_j (-alpha_{j+1}) y_{s-j}, S2 = sum_j beta_{j+1} a_{s-j}, j ascending from zero
// sums (second_order/mod.rs:93-121).  Zero coefficients are skipped: sum + v*0 == sum bit for bit for finite v
// (the running sums are never -0).  `a0`/`y0` are the newest (j = 0) values, already in registers.
template <bool EXACT>
__device__ __forceinline__ D3 qt_predict(const QtArgs& q, const double4* ry, const double* ra, int64_t n, int64_t k, D3 y0,
                                         D3 a0) {
    D3 s1 = {0.0, 0.0, 0.0}, s2 = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < kMaxOrder; ++j) {
        if (j < q.order) {
            const double ca = q.nalpha[j], cb = q.beta[j];
            if (ca != 0.0) {
                D3 yy;
                if (j == 0) {
                    yy = y0;
                } else {
                    double4 t = ry[(int64_t)q.slot[j] * n + k];
                    yy = {t.x, t.y, t.z};
                }
                s1 = madd3<EXACT>(s1, yy, ca);
            }
            if (cb != 0.0) {
                D3 aa = j == 0 ? a0 : ld_a(ra + (int64_t)q.slot[j] * 3 * n, n, k);
                s2 = madd3<EXACT>(s2, aa, cb);
            }
        }
    }
    return madd3<EXACT>(s1, s2, q.f);
}

// dy = (y_{s} - y_{s-1}) / h + W * g,  W = sum_j c_j a_{s-j}   (cowell.rs:34-52)
template <bool EXACT>
__device__ __forceinline__ D3 qt_velocity(const QtArgs& q, const double4* ry, const double* ra, int64_t n, int64_t k, D3 y0,
                                          D3 a0) {
    D3 w = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < kMaxOrder; ++j) {
        if (j < q.order) {
            const double c = q.cow[j];
            D3 aa = j == 0 ? a0 : ld_a(ra + (int64_t)q.slot[j] * 3 * n, n, k);
            w = madd3<EXACT>(w, aa, c);
        }
    }
    double4 t = ry[(int64_t)q.slot[1] * n + k];
    D3 ym1 = {t.x, t.y, t.z};
    D3 d;
    if (EXACT) {
        d = xdiv3(xsub3(y0, ym1), q.h);
        return xadd3(d, xmul3(w, q.g));
    }
    d = {(y0.x - ym1.x) / q.h, (y0.y - ym1.y) / q.h, (y0.z - ym1.z) / q.h};
    return {fma(w.x, q.g, d.x), fma(w.y, q.g, d.y), fma(w.z, q.g, d.z)};
}

// Everything that follows an acceleration evaluation for body k, fused behind the pair loop.
template <bool EXACT>
__device__ __forceinline__ void apply_epilogue(const EpArgs& ep, int64_t k, D3 a) {
    const int64_t n = ep.n;
    if (ep.kind == EP_STORE) {
        st_a(ep.a_out, n, k, a);
    } else if (ep.kind == EP_KD) {
        // dy += a * (h*B[s]);  y += dy * (h*A[s])     symplectic.rs:93-94
        st_a(ep.a_out, n, k, a);
        D3 v = ld_a(ep.dy, n, k);
        v = madd3<EXACT>(v, a, ep.kd.hb);
        st_a(ep.dy, n, k, v);
        double4 p = ep.y_in[k];
        D3 y = madd3<EXACT>(D3{p.x, p.y, p.z}, v, ep.kd.ha);
        ep.y_out[k] = make_double4(y.x, y.y, y.z, p.w);
    } else {
        // one full multistep step: record a_{s}, reconstruct the velocity, predict y_{s+1}
        const QtArgs& q = ep.qt;
        st_a(ep.ra + (int64_t)q.slot[0] * 3 * n, n, k, a);
        double4 p = ep.ry[(int64_t)q.slot[0] * n + k];
        D3 y0 = {p.x, p.y, p.z};
        D3 v = qt_velocity<EXACT>(q, ep.ry, ep.ra, n, k, y0, a);
        st_a(ep.dy, n, k, v);
        D3 yn = qt_predict<EXACT>(q, ep.ry, ep.ra, n, k, y0, a);
        ep.ry[(int64_t)q.slot_next * n + k] = make_double4(yn.x, yn.y, yn.z, p.w);
    }
}

// ---------------------------------------------------------------------------------------------------------
// PARITY acceleration: thread k walks every other body in index order and reproduces the reference's
// summation exactly:  ddy[k] = (((0 + c_{0->k}) + c_{1->k}) + ...) + (((0 + c_{k+1->k}) + c_{k+2->k}) + ...)
// with each pair evaluated in the reference's (i<j) orientation (nbody.rs:23-35).
// Pair formula (particular): dir = p_j - p_i; n = dir.dir; mag = n*sqrt(n);
//   a_i = dir*(mu_j/mag);  a_j = -(dir*(mu_i/mag)).
// `variant` selects the reading of particular's scalar factor (its source is not in the reference tree): 0 = mu / mag
// (the published form, default), 1 = mu * (1 / mag) -- one reciprocal, two products.  Both are kept bit-exact against the
// oracle's twin switch so that learning the true form costs a flag, not new kernels.
__device__ __forceinline__ double pair_scale(double mu, double mag, int variant) {
    return variant == 1 ? xmul(mu, xdiv(1.0, mag)) : xdiv(mu, mag);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_accel_parity(int64_t n, int64_t i0, int64_t i1, const double4* __restrict__ pm,
                                                        EpArgs ep, int variant) {
    __shared__ double4 tile[BLOCK];
    const int64_t k = i0 + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool active = k < i1;
    const double4 pk = active ? pm[k] : make_double4(0, 0, 0, 0);
    D3 acc = {0.0, 0.0, 0.0}, out = {0.0, 0.0, 0.0};
    for (int64_t base = 0; base < n; base += BLOCK) {
        const int64_t j = base + threadIdx.x;
        tile[threadIdx.x] = j < n ? pm[j] : make_double4(0, 0, 0, 0);
        __syncthreads();
        const int lim = (int)min((int64_t)BLOCK, n - base);
        if (active) {
            for (int t = 0; t < lim; ++t) {
                const int64_t jj = base + t;
                const double4 pj = tile[t];
                if (jj < k) {  // pair (jj, k): self = jj, other = k; we receive computed.1
                    D3 dir = {xsub(pk.x, pj.x), xsub(pk.y, pj.y), xsub(pk.z, pj.z)};
                    double nn = xdot3(dir, dir);
                    double mag = xmul(nn, xsqrt(nn));
                    double s = pair_scale(pj.w, mag, variant);
                    acc = xadd3(acc, xneg3(xmul3(dir, s)));
                } else if (jj > k) {  // pair (k, jj): self = k, other = jj; we receive computed.0
                    D3 dir = {xsub(pj.x, pk.x), xsub(pj.y, pk.y), xsub(pj.z, pk.z)};
                    double nn = xdot3(dir, dir);
                    double mag = xmul(nn, xsqrt(nn));
                    double s = pair_scale(pj.w, mag, variant);
                    out = xadd3(out, xmul3(dir, s));
                }
            }
        }
        __syncthreads();
    }
    if (active) apply_epilogue<true>(ep, k, xadd3(acc, out));
}

// ---------------------------------------------------------------------------------------------------------
// THROUGHPUT acceleration.
//   grid.x = target tiles of BLOCK bodies, grid.y = S source splits.  Sources are staged through shared memory
//   as double4 (x,y,z,mu) with coalesced 32-byte loads; each thread owns one target and keeps its partial
//   acceleration in registers.  With S > 1 every block writes its partial sum to part[s][3][n]; the block that
//   finishes a tile last (ticket counter) adds the S partials in split order -- deterministic -- and runs the
//   integrator epilogue for that tile, so a whole step is ONE launch.
//
//   Per interaction (FP64 pipe ops): 3 sub, 3 for r^2, 7 for mu*r^-3 (MUFU.RSQ64H seed on the XU pipe + one
//   cubic-convergent correction), 3 FMA accumulate = 16.
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

template <bool CHECK>
__device__ __forceinline__ void interact_fast(const double4 pi, const double4 pj, double& ax, double& ay, double& az) {
    const double dx = pj.x - pi.x;
    const double dy = pj.y - pi.y;
    const double dz = pj.z - pi.z;
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const double y0 = rsqrt_seed(r2);     // |1 - r2*y0^2| <~ 2^-19
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);   // exact residual (one rounding)
    const double p = fma(1.875, e, 1.5);  // (1-e)^(-3/2) = 1 + e*(3/2 + 15/8 e) + O(e^3), e^3 ~ 2^-57
    const double q = e * p;
    const double c = y2 * y0;
    const double m = pj.w * c;
    double s = fma(m, q, m);
    if (CHECK) s = r2 > 0.0 ? s : 0.0;    // self-interaction only (tile on the diagonal)
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
    az = fma(s, dz, az);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_accel_fast(int64_t n, int64_t i0, int64_t i1, int64_t j0, int64_t j1,
                                                      int64_t chunk, int splits, const double4* __restrict__ pm,
                                                      double* __restrict__ part, unsigned* __restrict__ tickets,
                                                      EpArgs ep) {
    __shared__ double4 tile[BLOCK];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int64_t tb = i0 + (int64_t)blockIdx.x * BLOCK;  // first target of this tile
    const int64_t i = tb + tid;
    const bool active = i < i1;
    const double4 pi = pm[active ? i : (i1 - 1)];
    const int64_t js = j0 + (int64_t)blockIdx.y * chunk;
    const int64_t je = min(j1, js + chunk);
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (int64_t base = js; base < je; base += BLOCK) {
        const int64_t j = base + tid;
        if (j < je) tile[tid] = pm[j];
        __syncthreads();
        const int lim = (int)min((int64_t)BLOCK, je - base);
        const bool diag = (base < tb + BLOCK) && (base + lim > tb);
        if (diag) {
#pragma unroll 4
            for (int t = 0; t < lim; ++t) interact_fast<true>(pi, tile[t], ax, ay, az);
        } else if (lim == BLOCK) {
#pragma unroll 8
            for (int t = 0; t < BLOCK; ++t) interact_fast<false>(pi, tile[t], ax, ay, az);
        } else {
#pragma unroll 4
            for (int t = 0; t < lim; ++t) interact_fast<false>(pi, tile[t], ax, ay, az);
        }
        __syncthreads();
    }
    if (splits == 1) {
        if (active) apply_epilogue<false>(ep, i, D3{ax, ay, az});
        return;
    }
    if (active) {
        double* p = part + (int64_t)blockIdx.y * 3 * n;
        __stcg(p + i, ax);
        __stcg(p + n + i, ay);
        __stcg(p + 2 * n + i, az);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&tickets[blockIdx.x], 1u);
        const int last = prev == (unsigned)(splits - 1);
        if (last) tickets[blockIdx.x] = 0u;  // self-resetting for the next launch
        s_last = last;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (active) {
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int s = 0; s < splits; ++s) {
            const double* p = part + (int64_t)s * 3 * n;
            sx += __ldcg(p + i);
            sy += __ldcg(p + n + i);
            sz += __ldcg(p + 2 * n + i);
        }
        apply_epilogue<false>(ep, i, D3{sx, sy, sz});
    }
}

// ---------------------------------------------------------------------------------------------------------
// Element-wise forms of the same epilogues (used when the acceleration arrives from a collective, for the FSAL
// stage of the starter, and for the first prediction after start-up).
template <bool EXACT>
__global__ void k_epilogue(int64_t i0, int64_t i1, const double* __restrict__ a_in, EpArgs ep) {
    const int64_t k = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= i1) return;
    apply_epilogue<EXACT>(ep, k, ld_a(a_in, ep.n, k));
}

// Prediction only: y_{s+1} from steps s, s-1, ...   Here slot[j] = slot of step (s-j), slot_next = slot of s+1.
template <bool EXACT>
__global__ void k_predict(int64_t i0, int64_t i1, int64_t n, QtArgs q, double4* ry, const double* ra) {
    const int64_t k = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= i1) return;
    double4 p = ry[(int64_t)q.slot[0] * n + k];
    D3 a0 = ld_a(ra + (int64_t)q.slot[0] * 3 * n, n, k);
    D3 yn = qt_predict<EXACT>(q, ry, ra, n, k, D3{p.x, p.y, p.z}, a0);
    ry[(int64_t)q.slot_next * n + k] = make_double4(yn.x, yn.y, yn.z, p.w);
}

}  // namespace ee
