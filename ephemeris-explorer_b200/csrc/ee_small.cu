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
#include <cstdlib>

#include "ee_engine.h"
#include "ee_kernels.cuh"

namespace ee {

constexpr int kSmallMaxN = 64;
constexpr int kSmallThreads = 512;
constexpr int kSmallMaxR = kMaxOrder + 1;

struct SmallArgs {
    int n, R, order, variant;
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
    // c[comp][partner][target]; rows padded to 65 doubles so that the pair phase's transposed store (consecutive
    // partners, fixed target) does not land all 32 lanes on one bank
    double c[3][kSmallMaxN][kSmallMaxN + 1];
    double racc[3][kSmallMaxN], rout[3][kSmallMaxN];  // row sums of the partners below / above the target (combined in phase 1)
    double mu[kSmallMaxN];
    unsigned short pi[kSmallMaxN * (kSmallMaxN - 1) / 2], pj[kSmallMaxN * (kSmallMaxN - 1) / 2];
};

// Sequential (reference-order) sum of col[t*stride] for t in [begin, end), software-pipelined: the loads of the next
// eight terms are in flight while the current eight are added, so the chain costs one DADD latency per term.
__device__ __forceinline__ double ordered_sum(const double* col, int stride, int begin, int end) {
    double s = 0.0;
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = begin + u < end ? col[(begin + u) * stride] : 0.0;
    for (int t = begin; t < end; t += 8) {
        double w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = t + 8 + u < end ? col[(t + 8 + u) * stride] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t + u < end) s = xadd(s, v[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = w[u];
    }
    return s;
}

template <int ORDER, bool PROF>
__global__ void __launch_bounds__(kSmallThreads, 1) k_small_steps(SmallArgs A, long long* prof) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem& S = *reinterpret_cast<SmallSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int n = A.n, R = A.R;
    constexpr int order = ORDER;
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
    // role of this thread: (body, component) owner for the linear combinations and the row sums
    const bool p3_active = tid < 3 * n;
    const int p3_c = tid / n, p3_k = tid % n;
    const bool p3b_active = tid >= 3 * n && tid < 6 * n;   // second half of the row sums: partners above the target
    const int p3b_c = (tid - 3 * n) / n, p3b_k = (tid - 3 * n) % n;
    const bool smp_active = A.stride && tid >= 6 * n && tid < 7 * n;
    bool pending = false;  // racc/rout hold the acceleration of the current state, not yet combined into S.a[cur]
    // this thread's first pair (all of them when n <= 32): shared-memory offsets are loop invariants
    const bool pair0 = tid < npairs;
    const int pr_i = pair0 ? S.pi[tid] : 0, pr_j = pair0 ? S.pj[tid] : 1;
    long long smp_stride = 0, smp_rem = 0, smp_next = 0;
    int smp_b = 0;
    if (smp_active) {
        smp_b = tid - 6 * n;
        smp_stride = A.stride[smp_b];
        if (smp_stride > 0) {
            smp_rem = A.steps_done0 % smp_stride;                                   // steps since the last sample
            smp_next = A.off[smp_b] + (A.steps_done0 / smp_stride + 1 - A.qbase[smp_b]);  // buffer slot of the next sample
        }
    }
    long long m = A.m0;
    long long pc[6] = {0, 0, 0, 0, 0, 0};  // PROF: cycles of phase 1 | barrier | phase 2 | barrier | phase 3 | barrier (this thread)
    for (long long step = 0; step < A.k_steps; ++step, ++m) {
        const int snew = cur + 1 == R ? 0 : cur + 1;
        long long tc0 = 0, tc1 = 0;
        if (PROF) tc0 = clock64();
        // ---- phase 1: thread (k, c) forms S1 and S2 over steps m, m-1, ... (two interleaved chains, every term kept --
        //      zero coefficients included, as the reference does) and the new position y = S1 + S2 * f
        if (p3_active) {
            if (pending) S.a[cur][p3_c][p3_k] = xadd(S.racc[p3_c][p3_k], S.rout[p3_c][p3_k]);  // ddy[k] += output_k
            double yv[ORDER], av[ORDER];
            int sl = cur;
#pragma unroll
            for (int j = 0; j < ORDER; ++j) {
                yv[j] = S.y[sl][p3_c][p3_k];
                av[j] = S.a[sl][p3_c][p3_k];
                sl = sl == 0 ? R - 1 : sl - 1;
            }
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int j = 0; j < ORDER; ++j) {
                s1 = xadd(s1, xmul(yv[j], A.nalpha[j]));
                s2 = xadd(s2, xmul(av[j], A.beta[j]));
            }
            S.y[snew][p3_c][p3_k] = xadd(s1, xmul(s2, A.f));
        }
        if (PROF) { tc1 = clock64(); pc[0] += tc1 - tc0; tc0 = tc1; }
        __syncthreads();
        if (PROF) { tc1 = clock64(); pc[1] += tc1 - tc0; tc0 = tc1; }
        // ---- phase 2: one thread per unordered pair
        for (int p = tid; p < npairs; p += kSmallThreads) {
            const int i = p == tid ? pr_i : S.pi[p], j = p == tid ? pr_j : S.pj[p];
            const D3 yi = {S.y[snew][0][i], S.y[snew][1][i], S.y[snew][2][i]};
            const D3 yj = {S.y[snew][0][j], S.y[snew][1][j], S.y[snew][2][j]};
            const D3 dir = xsub3(yj, yi);
            const double nn = xdot3(dir, dir);
            const double mag = xmul(nn, xsqrt(nn));
            const D3 ci = xmul3(dir, pair_scale(S.mu[j], mag, A.variant));         // computed.0: acceleration of i
            const D3 cj = xneg3(xmul3(dir, pair_scale(S.mu[i], mag, A.variant)));  // computed.1: acceleration of j
            S.c[0][j][i] = ci.x;
            S.c[1][j][i] = ci.y;
            S.c[2][j][i] = ci.z;
            S.c[0][i][j] = cj.x;
            S.c[1][i][j] = cj.y;
            S.c[2][i][j] = cj.z;
        }
        if (PROF) { tc1 = clock64(); pc[2] += tc1 - tc0; tc0 = tc1; }
        __syncthreads();
        if (PROF) { tc1 = clock64(); pc[3] += tc1 - tc0; tc0 = tc1; }
        // ---- phase 3: ordered row sums -> a_{m+1}; sampled positions go straight to HBM
        // Two threads per (body, component): one adds the partners below the target in ascending order, the other the
        // partners above it.  Both walk all n partners with the unwanted ones replaced by +0.0 -- the select sits on the
        // loaded value, off the dependent chain, and x + (+0.0) == x bit for bit because a running sum that starts at
        // +0.0 can never become -0.0.  The two halves meet in phase 1 of the next step.
        if (p3_active || p3b_active) {
            const bool lower = p3_active;
            const int c = lower ? p3_c : p3b_c, k = lower ? p3_k : p3b_k;
            const double* col = &S.c[c][0][k];
            double sum = 0.0;
            for (int t0 = 0; t0 < n; t0 += 32) {  // all loads of a batch are issued before its (dependent) adds
                double v[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const int t = t0 + u;  // rows up to kSmallMaxN-1 exist, so the load is always in bounds
                    const double x = col[t * (kSmallMaxN + 1)];
                    const bool take = lower ? t < k : (t > k && t < n);
                    v[u] = take ? x : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 32; ++u) sum = xadd(sum, v[u]);
            }
            if (lower)
                S.racc[c][k] = sum;
            else
                S.rout[c][k] = sum;
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
        if (PROF) { tc1 = clock64(); pc[4] += tc1 - tc0; tc0 = tc1; }
        __syncthreads();
        if (PROF) { tc1 = clock64(); pc[5] += tc1 - tc0; }
        cur = snew;
        pending = true;
    }
    if (pending && p3_active) S.a[cur][p3_c][p3_k] = xadd(S.racc[p3_c][p3_k], S.rout[p3_c][p3_k]);
    __syncthreads();
    if (PROF && prof) {
        for (int q = 0; q < 6; ++q) prof[(size_t)tid * 6 + q] = pc[q];
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
    return e.n <= kSmallMaxN && e.n >= 2 && e.mode == EE_MODE_PARITY && e.world == 1 && !e.srkn_main;
}

// Run k steady-state steps in one launch, starting from device state m0 (= steps completed on the device; the host's
// e.m may already be ahead: run-ahead) and solout step counter steps_done0.  Preconditions: m0 >= order (start-up done),
// solout capacity reserved.
void small_steps(NBodyEngine& e, int64_t m0, int64_t steps_done0, int64_t k) {
    // the opt-in shared-memory size is a per-device attribute of the function: track it per device (a process may run
    // small systems on several GPUs, DESIGN.md section 9 "replicas only"), thread-safe
    static std::atomic<uint64_t> attr_mask{0};
    const uint64_t bit = 1ull << (e.device & 63);
    if (!(attr_mask.load(std::memory_order_acquire) & bit)) {
        EE_CUDA(cudaFuncSetAttribute(k_small_steps<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)));
        EE_CUDA(cudaFuncSetAttribute(k_small_steps<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)));
        EE_CUDA(cudaFuncSetAttribute(k_small_steps<13, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)));
        EE_CUDA(cudaFuncSetAttribute(k_small_steps<13, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)));
        attr_mask.fetch_or(bit, std::memory_order_release);
    }
    const char* penv = getenv("EE_SMALL_PROFILE");  // developer aid: per-phase cycle counts to stderr
    QtArgs q = e.qt_args(m0, m0 + 1);
    SmallArgs A{};
    A.n = (int)e.n;
    A.R = e.R;
    A.order = e.order;
    A.variant = e.pair_variant;
    A.m0 = m0;
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
        A.steps_done0 = steps_done0;
    }
    if (penv && penv[0] == '1') {
        DBuf<long long> d((size_t)kSmallThreads * 6);
        if (e.order == 12)
            k_small_steps<12, true><<<1, kSmallThreads, sizeof(SmallSmem), e.stream>>>(A, d.p);
        else
            k_small_steps<13, true><<<1, kSmallThreads, sizeof(SmallSmem), e.stream>>>(A, d.p);
        EE_CUDA(cudaGetLastError());
        std::vector<long long> hp((size_t)kSmallThreads * 6);
        EE_CUDA(cudaMemcpyAsync(hp.data(), d.p, hp.size() * 8, cudaMemcpyDeviceToHost, e.stream));
        EE_CUDA(cudaStreamSynchronize(e.stream));
        const int probes[] = {0, (int)e.n - 1, 3 * (int)e.n - 1, 3 * (int)e.n, kSmallThreads - 1};
        for (int t : probes)
            fprintf(stderr, "[small-profile] k=%lld tid=%d cycles/step: p1 %.0f bar %.0f p2 %.0f bar %.0f p3 %.0f bar %.0f\n", (long long)k, t,
                    (double)hp[(size_t)t * 6] / k, (double)hp[(size_t)t * 6 + 1] / k, (double)hp[(size_t)t * 6 + 2] / k,
                    (double)hp[(size_t)t * 6 + 3] / k, (double)hp[(size_t)t * 6 + 4] / k, (double)hp[(size_t)t * 6 + 5] / k);
    } else {
        if (e.order == 12)
            k_small_steps<12, false><<<1, kSmallThreads, sizeof(SmallSmem), e.stream>>>(A, nullptr);
        else
            k_small_steps<13, false><<<1, kSmallThreads, sizeof(SmallSmem), e.stream>>>(A, nullptr);
    }
    EE_CUDA(cudaGetLastError());
    count_launch();
    e.accel_launches++;
}

}  // namespace ee
