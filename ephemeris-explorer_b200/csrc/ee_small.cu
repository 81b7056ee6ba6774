// ee_small.cu -- persistent single-CTA stepper for small systems (n <= 64: the reference's own solar-system configs).
//
// At 3..64 bodies a step is a short dependency chain, not a throughput problem: 496 pair evaluations and three
// 12-term linear combinations.  Launching kernels per step would cost more than the arithmetic, so one CTA keeps the
// whole multistep history in shared memory and runs K steps per launch with __syncthreads() as the only
// synchronisation -- no host round trip, no global memory traffic inside the loop except the sampled positions.
//
// Bit-exact parity with the reference order (same rules as k_accel_parity / apply_epilogue<true>):
//   pairs     one thread per unordered pair (i<j), evaluated in the reference's orientation        nbody.rs:23-35
//   row sums  thread (k, c) adds the contributions of bodies i<k in ascending order, then the
//             pre-summed contributions of bodies j>k, exactly like `ddy[j] += ..` / `ddy[i] += output_i`
//   S1/S2/W   left-to-right from zero, separate multiply and add                                   second_order/mod.rs:93-121, cowell.rs:34-52
#include "ee_engine.h"
#include "ee_kernels.cuh"

namespace ee {

constexpr int kSmallMaxN = 64;
constexpr int kSmallThreads = 512;
constexpr int kSmallMaxR = kMaxOrder + 1;

struct SmallArgs {
    int n, R, order;
    long long m0;      // index of the current state (steps completed)
    long long k_steps;  // steps to run in this launch
    double nalpha[kMaxOrder], beta[kMaxOrder], cow[kMaxOrder];
    double f, g, h;
    double4* ry;  // [R][n]
    double* ra;   // [R][3][n]
    double* dy;   // [3][n]
    // sampling (may be null)
    const int64_t* stride;
    const int64_t* off;
    const int64_t* qbase;
    double* samples;
    long long steps_done0;  // solout step counter at launch
};

struct SmallSmem {
    double y[kSmallMaxR][3][kSmallMaxN];
    double a[kSmallMaxR][3][kSmallMaxN];
    double c[3][kSmallMaxN][kSmallMaxN];  // c[comp][partner][target]
    double s1[3][kSmallMaxN], s2[3][kSmallMaxN];
    double mu[kSmallMaxN];
    unsigned short pi[kSmallMaxN * (kSmallMaxN - 1) / 2], pj[kSmallMaxN * (kSmallMaxN - 1) / 2];
};

__global__ void __launch_bounds__(kSmallThreads, 1) k_small_steps(SmallArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem& S = *reinterpret_cast<SmallSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int n = A.n, R = A.R, order = A.order;
    const int npairs = n * (n - 1) / 2;

    // ---- load the ring
    for (int idx = tid; idx < R * n; idx += kSmallThreads) {
        const int r = idx / n, k = idx % n;
        const double4 p = A.ry[(size_t)r * n + k];
        S.y[r][0][k] = p.x;
        S.y[r][1][k] = p.y;
        S.y[r][2][k] = p.z;
        if (r == 0) S.mu[k] = p.w;
        for (int c = 0; c < 3; ++c) S.a[r][c][k] = A.ra[((size_t)r * 3 + c) * n + k];
    }
    for (int p = tid; p < npairs; p += kSmallThreads) {  // enumerate pairs (i<j) row by row
        int i = 0, rem = p;
        while (rem >= n - 1 - i) {
            rem -= n - 1 - i;
            ++i;
        }
        S.pi[p] = (unsigned short)i;
        S.pj[p] = (unsigned short)(i + 1 + rem);
    }
    __syncthreads();

    // Ring slots and sampling counters are tracked incrementally: 64-bit divisions inside the step loop would cost more
    // than the physics.
    int cur = (int)(A.m0 % R);  // slot of the current state (step m)
    // role of this thread in phase 1 / 3
    const bool p1_active = tid < 6 * n;
    const bool p1_second = tid >= 3 * n;
    const int p1_idx = p1_second ? tid - 3 * n : tid;
    const int p1_c = p1_idx / n, p1_k = p1_idx % n;
    const bool p3_active = tid < 3 * n;
    const int p3_c = tid / n, p3_k = tid % n;
    const bool smp_active = A.stride && tid >= 3 * n && tid < 4 * n;
    long long smp_stride = 0, smp_rem = 0, smp_next = 0;
    int smp_b = 0;
    if (smp_active) {
        smp_b = tid - 3 * n;
        smp_stride = A.stride[smp_b];
        if (smp_stride > 0) {
            smp_rem = A.steps_done0 % smp_stride;                                   // steps since the last sample
            smp_next = A.off[smp_b] + (A.steps_done0 / smp_stride + 1 - A.qbase[smp_b]);  // buffer slot of the next sample
        }
    }
    long long m = A.m0;
    for (long long step = 0; step < A.k_steps; ++step, ++m) {
        const int snew = cur + 1 == R ? 0 : cur + 1;
        // ---- phase 1: S1 (threads [0,3n)) and S2 (threads [3n,6n)) over steps m, m-1, ...
        if (p1_active) {
            double s = 0.0;
            int sl = cur;
#pragma unroll
            for (int j = 0; j < kMaxOrder; ++j) {
                if (j < order) {
                    const double coef = p1_second ? A.beta[j] : A.nalpha[j];
                    if (coef != 0.0) s = xadd(s, xmul(p1_second ? S.a[sl][p1_c][p1_k] : S.y[sl][p1_c][p1_k], coef));
                    sl = sl == 0 ? R - 1 : sl - 1;
                }
            }
            if (p1_second)
                S.s2[p1_c][p1_k] = s;
            else
                S.s1[p1_c][p1_k] = s;
        }
        __syncthreads();
        // ---- phase 2: new positions (every consumer recomputes S1 + S2*f; threads [0,3n) also store them) and pairs
        if (p3_active) S.y[snew][p3_c][p3_k] = xadd(S.s1[p3_c][p3_k], xmul(S.s2[p3_c][p3_k], A.f));
        for (int p = tid; p < npairs; p += kSmallThreads) {
            const int i = S.pi[p], j = S.pj[p];
            D3 yi, yj;
            yi.x = xadd(S.s1[0][i], xmul(S.s2[0][i], A.f));
            yi.y = xadd(S.s1[1][i], xmul(S.s2[1][i], A.f));
            yi.z = xadd(S.s1[2][i], xmul(S.s2[2][i], A.f));
            yj.x = xadd(S.s1[0][j], xmul(S.s2[0][j], A.f));
            yj.y = xadd(S.s1[1][j], xmul(S.s2[1][j], A.f));
            yj.z = xadd(S.s1[2][j], xmul(S.s2[2][j], A.f));
            const D3 dir = xsub3(yj, yi);
            const double nn = xdot3(dir, dir);
            const double mag = xmul(nn, xsqrt(nn));
            const D3 ci = xmul3(dir, xdiv(S.mu[j], mag));         // computed.0: acceleration of i
            const D3 cj = xneg3(xmul3(dir, xdiv(S.mu[i], mag)));  // computed.1: acceleration of j
            S.c[0][j][i] = ci.x;
            S.c[1][j][i] = ci.y;
            S.c[2][j][i] = ci.z;
            S.c[0][i][j] = cj.x;
            S.c[1][i][j] = cj.y;
            S.c[2][i][j] = cj.z;
        }
        __syncthreads();
        // ---- phase 3: ordered row sums -> a_{m+1}; sampled positions go straight to HBM
        if (p3_active) {
            double acc = 0.0, out = 0.0;
            for (int i = 0; i < p3_k; ++i) acc = xadd(acc, S.c[p3_c][i][p3_k]);
            for (int j = p3_k + 1; j < n; ++j) out = xadd(out, S.c[p3_c][j][p3_k]);
            S.a[snew][p3_c][p3_k] = xadd(acc, out);
        } else if (smp_active && smp_stride > 0) {
            smp_rem += 1;
            if (smp_rem == smp_stride) {
                smp_rem = 0;
                A.samples[3 * smp_next] = S.y[snew][0][smp_b];
                A.samples[3 * smp_next + 1] = S.y[snew][1][smp_b];
                A.samples[3 * smp_next + 2] = S.y[snew][2][smp_b];
                smp_next += 1;
            }
        }
        __syncthreads();
        cur = snew;
    }

    // ---- velocity of the final state only (Cowell; an output, never fed back) and write-back
    if (A.k_steps > 0 && tid < 3 * n) {
        const int c = tid / n, k = tid % n;
        double w = 0.0;
        for (int j = 0; j < order; ++j) {
            const int sl = (int)(((m - j) % R + R) % R);
            w = xadd(w, xmul(S.a[sl][c][k], A.cow[j]));
        }
        const int s0 = (int)(m % R), s1 = (int)(((m - 1) % R + R) % R);
        const double d = xdiv(xsub(S.y[s0][c][k], S.y[s1][c][k]), A.h);
        A.dy[(size_t)c * n + k] = xadd(d, xmul(w, A.g));
    }
    for (int idx = tid; idx < R * n; idx += kSmallThreads) {
        const int r = idx / n, k = idx % n;
        A.ry[(size_t)r * n + k] = make_double4(S.y[r][0][k], S.y[r][1][k], S.y[r][2][k], S.mu[k]);
        for (int c = 0; c < 3; ++c) A.ra[((size_t)r * 3 + c) * n + k] = S.a[r][c][k];
    }
}

bool small_path_available(const NBodyEngine& e) {
    return e.n <= kSmallMaxN && e.n >= 2 && e.mode == EE_MODE_PARITY && e.world == 1;
}

// Run k steady-state steps in one launch.  Preconditions: e.m >= order (start-up done), solout capacity reserved.
void small_steps(NBodyEngine& e, int64_t k) {
    static bool attr_set = false;
    if (!attr_set) {
        EE_CUDA(cudaFuncSetAttribute(k_small_steps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)));
        attr_set = true;
    }
    QtArgs q = e.qt_args(e.m, e.m + 1);
    SmallArgs A{};
    A.n = (int)e.n;
    A.R = e.R;
    A.order = e.order;
    A.m0 = e.m;
    A.k_steps = k;
    for (int j = 0; j < kMaxOrder; ++j) {
        A.nalpha[j] = q.nalpha[j];
        A.beta[j] = q.beta[j];
        A.cow[j] = q.cow[j];
    }
    A.f = q.f;
    A.g = q.g;
    A.h = q.h;
    A.ry = e.ry.p;
    A.ra = e.ra.p;
    A.dy = e.dy.p;
    if (e.solout) {
        A.stride = e.solout->d_stride.p;
        A.off = e.solout->d_off.p;
        A.qbase = e.solout->d_qbase.p;
        A.samples = e.solout->samples.p;
        A.steps_done0 = e.solout->steps_done;
    }
    k_small_steps<<<1, kSmallThreads, sizeof(SmallSmem), e.stream>>>(A);
    EE_CUDA(cudaGetLastError());
    count_launch();
    e.accel_launches++;
}

}  // namespace ee
