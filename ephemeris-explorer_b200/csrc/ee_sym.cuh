// ee_sym.cuh -- throughput acceleration kernel that uses Newton's third law: each unordered pair is evaluated ONCE
// and applied to both bodies (as the reference's loop does, nbody.rs:23-35), which cuts the FP64-pipe work from 16 to
// ~10.4 instructions per directed interaction.
//
// Decomposition.  Bodies are cut into I-tiles of 1024 (256 threads x 4 targets held in registers) and J-superchunks
// of 512.  A work item is (I-tile ti, superchunk sj) with some j > i in it (sj >= 2*ti).  Persistent CTAs (2 per SM)
// pull items from an atomic counter.  Inside an item every warp walks the superchunk in chunks of 32 bodies: lane l
// pairs its four targets with body (l + k) mod 32 of the chunk at rotation k, so at any instant the 32 lanes touch 32
// different j -- the j-side partial sums live in a warp-private shared-memory array and are updated with plain,
// conflict-free read-modify-writes (no atomics).  At the end of the item the eight warps' arrays are added in warp order
// and written to part_j[ti][j]; the register accumulators go to part_i[sj][i].  A second, tiny kernel adds each body's
// partials in a fixed order and runs the integrator epilogue.  Every sum has a fixed order => deterministic.
//
// Per rotation (4 pairs): 4 x (3 sub + 3 r^2 + 6 r^-3 + 1 mu_j*t + 3 FMA_i + 1 mu_i*t + 3 FMA_j) + 3 adds into shared
// = 83 FP64-pipe instructions for 8 directed interactions; shared memory: 4 loads (x, y, z, mu of j) + 3 loads + 3 stores.
#pragma once
#include "ee_kernels.cuh"

namespace ee {

constexpr int kSymThreads = 256;
constexpr int kSymWarps = kSymThreads / 32;
constexpr int kSymTI = 4;
constexpr int kSymTile = kSymThreads * kSymTI;  // 1024 bodies per I-tile
// bodies per J-superchunk: a template parameter (512 / 256 / 128) so that a sharded run still has enough work items
// per GPU to balance 2 x 148 persistent CTAs; `ratio` = superchunks per I-tile

template <int JS>
struct SymSmem {
    double wacc[kSymWarps][3][JS];
    double sx[kSymWarps][32], sy[kSymWarps][32], sz[kSymWarps][32], sm[kSymWarps][32];
};

// canonical item numbering: items of tile ti are (ti, sj) for sj = ratio*ti .. ns-1
__host__ __device__ inline long long sym_item_prefix(long long ti, long long ns, long long ratio) {
    return ti * ns - ratio * (ti * (ti - 1) / 2);
}

__device__ __forceinline__ double sym_rcube(double r2) {  // r^-3 from r^2 (see interact_fast)
    const double y0 = rsqrt_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    const double c = y2 * y0;
    return fma(c, q, c);
}

template <int JS>
__global__ void __launch_bounds__(kSymThreads, 2) k_accel_sym(int64_t n, const double4* __restrict__ pm, long long item_lo,
                                                              long long item_hi, unsigned long long* __restrict__ counter,
                                                              double* __restrict__ part_i, double* __restrict__ part_j) {
    extern __shared__ __align__(16) unsigned char sym_raw[];
    SymSmem<JS>& S = *reinterpret_cast<SymSmem<JS>*>(sym_raw);
    constexpr int kSymJS = JS;
    constexpr long long kSymRatio = kSymTile / JS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ns = n / kSymJS, nt = n / kSymTile;
    for (;;) {
        __shared__ long long s_item;
        if (tid == 0) {
            const long long t = item_lo + (long long)atomicAdd(counter, 1ull);
            s_item = t < item_hi ? t : -1;
        }
        __syncthreads();
        const long long item = s_item;
        if (item < 0) break;
        // decode (ti, sj) from the canonical index
        long long ti = 0;
        {
            long long lo = 0, hi = nt - 1;
            while (lo < hi) {  // largest ti with prefix(ti) <= item
                const long long mid = (lo + hi + 1) >> 1;
                if (sym_item_prefix(mid, ns, kSymRatio) <= item) lo = mid; else hi = mid - 1;
            }
            ti = lo;
        }
        const long long sj = (long long)kSymRatio * ti + (item - sym_item_prefix(ti, ns, kSymRatio));
        const long long ibase = ti * kSymTile + warp * (32 * kSymTI) + lane;
        const long long jbase = sj * kSymJS;
        const bool diag = jbase < (ti + 1) * kSymTile;  // some j <= some i: pairs must be masked to j > i

        double xi[kSymTI], yi[kSymTI], zi[kSymTI], mi[kSymTI], ax[kSymTI], ay[kSymTI], az[kSymTI];
#pragma unroll
        for (int t = 0; t < kSymTI; ++t) {
            const double4 p = pm[ibase + 32 * t];
            xi[t] = p.x;
            yi[t] = p.y;
            zi[t] = p.z;
            mi[t] = p.w;
            ax[t] = ay[t] = az[t] = 0.0;
        }
        for (int k = lane; k < kSymJS; k += 32) {
            S.wacc[warp][0][k] = 0.0;
            S.wacc[warp][1][k] = 0.0;
            S.wacc[warp][2][k] = 0.0;
        }
        for (int chunk = 0; chunk < kSymJS / 32; ++chunk) {
            const long long j0 = jbase + chunk * 32;
            {
                const double4 p = pm[j0 + lane];
                S.sx[warp][lane] = p.x;
                S.sy[warp][lane] = p.y;
                S.sz[warp][lane] = p.z;
                S.sm[warp][lane] = p.w;
            }
            __syncwarp();
            double* wx = &S.wacc[warp][0][chunk * 32];
            double* wy = &S.wacc[warp][1][chunk * 32];
            double* wz = &S.wacc[warp][2][chunk * 32];
#pragma unroll 2
            for (int k = 0; k < 32; ++k) {
                const int jj = (lane + k) & 31;
                const double xj = S.sx[warp][jj], yj = S.sy[warp][jj], zj = S.sz[warp][jj], mj = S.sm[warp][jj];
                double bx = 0.0, by = 0.0, bz = 0.0;
#pragma unroll
                for (int t = 0; t < kSymTI; ++t) {
                    const double dx = xj - xi[t];
                    const double dy = yj - yi[t];
                    const double dz = zj - zi[t];
                    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    double rc = sym_rcube(r2);
                    if (diag) rc = (j0 + jj) > (ibase + 32 * t) ? rc : 0.0;
                    const double si = mj * rc;
                    ax[t] = fma(si, dx, ax[t]);
                    ay[t] = fma(si, dy, ay[t]);
                    az[t] = fma(si, dz, az[t]);
                    const double sjv = mi[t] * rc;
                    bx = fma(-sjv, dx, bx);
                    by = fma(-sjv, dy, by);
                    bz = fma(-sjv, dz, bz);
                }
                wx[jj] += bx;
                wy[jj] += by;
                wz[jj] += bz;
                __syncwarp();
            }
        }
        // i side: registers -> part_i[sj][c][i]
        {
            double* pi = part_i + (size_t)sj * 3 * n;
#pragma unroll
            for (int t = 0; t < kSymTI; ++t) {
                pi[ibase + 32 * t] = ax[t];
                pi[n + ibase + 32 * t] = ay[t];
                pi[2 * n + ibase + 32 * t] = az[t];
            }
        }
        __syncthreads();
        // j side: add the eight warps' arrays in warp order -> part_j[ti][c][j]
        {
            double* pj = part_j + (size_t)ti * 3 * n;
            for (int idx = tid; idx < 3 * kSymJS; idx += kSymThreads) {
                const int c = idx / kSymJS, k = idx % kSymJS;
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kSymWarps; ++w) s += S.wacc[w][c][k];
                pj[(size_t)c * n + jbase + k] = s;
            }
        }
        __syncthreads();
    }
}

// One 32-body chunk against this lane's four targets (32 rotations), masked to j > i when DIAG.  Written as a template so
// that the unmasked version carries no compare/select at all (the static kernel below picks one per item).
template <bool DIAG>
__device__ __forceinline__ void sym_chunk(const double* sx, const double* sy, const double* sz, const double* sm, double* wx,
                                          double* wy, double* wz, int lane, long long j0, long long ibase,
                                          const double (&xi)[kSymTI], const double (&yi)[kSymTI], const double (&zi)[kSymTI],
                                          const double (&mi)[kSymTI], double (&ax)[kSymTI], double (&ay)[kSymTI],
                                          double (&az)[kSymTI]) {
#pragma unroll 2
    for (int k = 0; k < 32; ++k) {
        const int jj = (lane + k) & 31;
        const double xj = sx[jj], yj = sy[jj], zj = sz[jj], mj = sm[jj];
        double bx = 0.0, by = 0.0, bz = 0.0;
#pragma unroll
        for (int t = 0; t < kSymTI; ++t) {
            const double dx = xj - xi[t];
            const double dy = yj - yi[t];
            const double dz = zj - zi[t];
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            double rc = sym_rcube(r2);
            if (DIAG) rc = (j0 + jj) > (ibase + 32 * t) ? rc : 0.0;
            const double si = mj * rc;
            ax[t] = fma(si, dx, ax[t]);
            ay[t] = fma(si, dy, ay[t]);
            az[t] = fma(si, dz, az[t]);
            const double sjv = mi[t] * rc;
            bx = fma(-sjv, dx, bx);
            by = fma(-sjv, dy, by);
            bz = fma(-sjv, dz, bz);
        }
        wx[jj] += bx;
        wy[jj] += by;
        wz[jj] += bz;
        __syncwarp();
    }
}

template <int JS>
__global__ void __launch_bounds__(kSymThreads, 2) k_accel_sym_static(int64_t n, const double4* __restrict__ pm, long long item_lo,
                                                                     long long item_hi, double* __restrict__ part_i,
                                                                     double* __restrict__ part_j) {
    extern __shared__ __align__(16) unsigned char sym_raw[];
    SymSmem<JS>& S = *reinterpret_cast<SymSmem<JS>*>(sym_raw);
    constexpr int kSymJS = JS;
    constexpr long long kSymRatio = kSymTile / JS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ns = n / kSymJS, nt = n / kSymTile;
    // Static chunk-granular split ("stream-K"): the rank's items are a list of 32-body chunk units; CTA g owns units
    // [ubegin(g), ubegin(g+1)) -- an equal share to within one unit -- so its range may start and end inside an item.
    constexpr int kChunks = JS / 32;
    const long long u_lo = item_lo * kChunks, u_len = (item_hi - item_lo) * kChunks;
    long long u = u_lo + u_len * (long long)blockIdx.x / (long long)gridDim.x;
    const long long u_end = u_lo + u_len * ((long long)blockIdx.x + 1) / (long long)gridDim.x;
    while (u < u_end) {
        const long long item = u / kChunks;
        const int c0 = (int)(u % kChunks);
        const int c1 = (int)min((long long)kChunks, c0 + (u_end - u));
        u += c1 - c0;
        // decode (ti, sj) from the canonical index
        long long ti = 0;
        {
            long long lo = 0, hi = nt - 1;
            while (lo < hi) {  // largest ti with prefix(ti) <= item
                const long long mid = (lo + hi + 1) >> 1;
                if (sym_item_prefix(mid, ns, kSymRatio) <= item) lo = mid; else hi = mid - 1;
            }
            ti = lo;
        }
        const long long sj = (long long)kSymRatio * ti + (item - sym_item_prefix(ti, ns, kSymRatio));
        const long long ibase = ti * kSymTile + warp * (32 * kSymTI) + lane;
        const long long jbase = sj * kSymJS;
        const bool diag = jbase < (ti + 1) * kSymTile;  // some j <= some i: pairs must be masked to j > i

        double xi[kSymTI], yi[kSymTI], zi[kSymTI], mi[kSymTI], ax[kSymTI], ay[kSymTI], az[kSymTI];
#pragma unroll
        for (int t = 0; t < kSymTI; ++t) {
            const double4 p = pm[ibase + 32 * t];
            xi[t] = p.x;
            yi[t] = p.y;
            zi[t] = p.z;
            mi[t] = p.w;
            ax[t] = ay[t] = az[t] = 0.0;
        }
        for (int k = c0 * 32 + lane; k < c1 * 32; k += 32) {
            S.wacc[warp][0][k] = 0.0;
            S.wacc[warp][1][k] = 0.0;
            S.wacc[warp][2][k] = 0.0;
        }
        for (int chunk = c0; chunk < c1; ++chunk) {
            const long long j0 = jbase + chunk * 32;
            {
                const double4 p = pm[j0 + lane];
                S.sx[warp][lane] = p.x;
                S.sy[warp][lane] = p.y;
                S.sz[warp][lane] = p.z;
                S.sm[warp][lane] = p.w;
            }
            __syncwarp();
            double* wx = &S.wacc[warp][0][chunk * 32];
            double* wy = &S.wacc[warp][1][chunk * 32];
            double* wz = &S.wacc[warp][2][chunk * 32];
            if (diag)
                sym_chunk<true>(S.sx[warp], S.sy[warp], S.sz[warp], S.sm[warp], wx, wy, wz, lane, j0, ibase, xi, yi, zi, mi, ax, ay, az);
            else
                sym_chunk<false>(S.sx[warp], S.sy[warp], S.sz[warp], S.sm[warp], wx, wy, wz, lane, j0, ibase, xi, yi, zi, mi, ax, ay, az);
        }
        // i side: registers -> part_i[2*sj + slot][c][i]; slot 0 belongs to the CTA that starts the item, slot 1 to the one
        // that finishes it when a range boundary falls inside the item (at most one does: ranges are >= one item long)
        {
            double* pi = part_i + (size_t)(2 * sj + (c0 == 0 ? 0 : 1)) * 3 * n;
#pragma unroll
            for (int t = 0; t < kSymTI; ++t) {
                pi[ibase + 32 * t] = ax[t];
                pi[n + ibase + 32 * t] = ay[t];
                pi[2 * n + ibase + 32 * t] = az[t];
            }
        }
        __syncthreads();
        // j side: add the eight warps' arrays in warp order -> part_j[ti][c][j]
        {
            double* pj = part_j + (size_t)ti * 3 * n;
            const int span = (c1 - c0) * 32;
            for (int idx = tid; idx < 3 * span; idx += kSymThreads) {
                const int c = idx / span, k = c0 * 32 + idx % span;
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kSymWarps; ++w) s += S.wacc[w][c][k];
                pj[(size_t)c * n + jbase + k] = s;
            }
        }
        __syncthreads();
    }
}


// Adds body b's partials in a fixed order (i-side superchunks ascending, then j-side tiles ascending), keeping only the
// items this rank owns, and runs the epilogue.
template <int JS>
__global__ void k_sym_reduce(int64_t n, long long item_lo, long long item_hi, const double* __restrict__ part_i,
                             const double* __restrict__ part_j, EpArgs ep) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    constexpr int kSymJS = JS;
    constexpr long long kSymRatio = kSymTile / JS;
    const long long ns = n / kSymJS;
    const long long tb = b / kSymTile, sb = b / kSymJS;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    {   // i side: this body's items (tb, sj) have consecutive canonical indices, so the owned ones are one sj range
        const long long base = sym_item_prefix(tb, ns, kSymRatio);
        long long s0 = kSymRatio * tb, s1 = ns;
        if (item_lo > base) s0 += item_lo - base;
        if (item_hi - base < ns - kSymRatio * tb) s1 = kSymRatio * tb + (item_hi - base);
        for (long long sj = s0; sj < s1; ++sj) {
            const double* p = part_i + (size_t)sj * 3 * n;
            sx += p[b];
            sy += p[n + b];
            sz += p[2 * n + b];
        }
    }
    for (long long ti = 0; ti <= sb / kSymRatio; ++ti) {  // j side: at most n/1024 candidates
        const long long idx = sym_item_prefix(ti, ns, kSymRatio) + (sb - (long long)kSymRatio * ti);
        if (idx < item_lo || idx >= item_hi) continue;
        const double* p = part_j + (size_t)ti * 3 * n;
        sx += p[b];
        sy += p[n + b];
        sz += p[2 * n + b];
    }
    apply_epilogue<false>(ep, b, D3{sx, sy, sz});
}


// Same reduction for the static split: an item's i-side sum may come in two slots (split[item - item_lo] != 0).
template <int JS>
__global__ void k_sym_reduce_static(int64_t n, long long item_lo, long long item_hi, const double* __restrict__ part_i,
                                    const double* __restrict__ part_j, const unsigned char* __restrict__ split, EpArgs ep) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    constexpr int kSymJS = JS;
    constexpr long long kSymRatio = kSymTile / JS;
    const long long ns = n / kSymJS;
    const long long tb = b / kSymTile, sb = b / kSymJS;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    {   // i side: this body's items (tb, sj) have consecutive canonical indices, so the owned ones are one sj range
        const long long base = sym_item_prefix(tb, ns, kSymRatio);
        long long s0 = kSymRatio * tb, s1 = ns;
        if (item_lo > base) s0 += item_lo - base;
        if (item_hi - base < ns - kSymRatio * tb) s1 = kSymRatio * tb + (item_hi - base);
        for (long long sj = s0; sj < s1; ++sj) {
            const double* p = part_i + (size_t)(2 * sj) * 3 * n;
            sx += p[b];
            sy += p[n + b];
            sz += p[2 * n + b];
            if (split[base + (sj - kSymRatio * tb) - item_lo]) {  // a second CTA finished this item
                const double* p1 = p + (size_t)3 * n;
                sx += p1[b];
                sy += p1[n + b];
                sz += p1[2 * n + b];
            }
        }
    }
    for (long long ti = 0; ti <= sb / kSymRatio; ++ti) {  // j side: at most n/1024 candidates
        const long long idx = sym_item_prefix(ti, ns, kSymRatio) + (sb - (long long)kSymRatio * ti);
        if (idx < item_lo || idx >= item_hi) continue;
        const double* p = part_j + (size_t)ti * 3 * n;
        sx += p[b];
        sy += p[n + b];
        sz += p[2 * n + b];
    }
    apply_epilogue<false>(ep, b, D3{sx, sy, sz});
}



// ---------------------------------------------------------------------------------------------------------
// Multi-GPU without NCCL on the data path: every rank evaluates its share of the pair items and reduces them locally
// to one partial acceleration per body (k_sym_reduce, 1.5 MB).  After a cross-GPU flag barrier each rank finishes the
// bodies of ITS slice: it adds the G partial accelerations in rank order, reading the peers' over NVLink (peer loads
// through CUDA-IPC mappings), runs the integrator epilogue for the slice and stores the new positions straight into
// every peer's ring (peer stores) -- reduce-scatter, epilogue and all-gather in one kernel, no collective library.
constexpr int kMaxPeers = 8;
struct PeerTable {
    int world, rank;
    const double* a_part[kMaxPeers];       // [3][n] partial accelerations of rank q
    double4* ry[kMaxPeers];                // ring of positions of rank q
    double* ra[kMaxPeers];                 // ring of accelerations of rank q
    double* dy[kMaxPeers];                 // velocities of rank q
    unsigned long long* flags[kMaxPeers];  // flags[q][r]: written by rank r, lives on rank q
};

__global__ void k_peer_finish(int64_t n, int64_t b0, int64_t b1, PeerTable T, EpArgs ep) {
    const int64_t b = b0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= b1) return;
    double px[kMaxPeers], py[kMaxPeers], pz[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {  // all (remote) loads first, then the ordered sum
        if (q < T.world) {
            const double* p = T.a_part[q];
            px[q] = p[b];
            py[q] = p[n + b];
            pz[q] = p[2 * n + b];
        } else {
            px[q] = py[q] = pz[q] = 0.0;
        }
    }
    double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {
        if (q < T.world) {
            sx += px[q];
            sy += py[q];
            sz += pz[q];
        }
    }
    apply_epilogue<false>(ep, b, D3{sx, sy, sz});
    if (ep.kind == EP_QT) {  // all-gather by peer stores: everything the epilogue produced for this body goes to every rank
        const size_t at = (size_t)ep.qt.slot_next * n + b;
        const double4 v = ep.ry[at];
        const size_t aoff = (size_t)ep.qt.slot[0] * 3 * n;
        const D3 vel = ld_a(ep.dy, n, b);
        for (int q = 0; q < T.world; ++q) {
            if (q == T.rank) continue;
            T.ry[q][at] = v;
            st_a(T.ra[q] + aoff, n, b, D3{sx, sy, sz});
            st_a(T.dy[q], n, b, vel);
        }
    }
}

// Flag barrier across the GPUs of one node: rank r stores `epoch` into slot r of every peer's flag array, then waits
// until all slots of its own array reach `epoch`.  Epochs only grow, so nothing is ever reset.  A spin that exceeds
// ~4e9 cycles records an error instead of hanging the GPU.
__global__ void k_peer_barrier(PeerTable T, unsigned long long epoch, int* err) {
    const int q = threadIdx.x;
    if (q >= T.world) return;
    __threadfence_system();
    volatile unsigned long long* out = T.flags[q] + T.rank;
    *out = epoch;
    volatile unsigned long long* in = T.flags[T.rank] + q;
    const long long t0 = clock64();
    while (*in < epoch) {
        if (clock64() - t0 > 4000000000LL) {
            *err = 1;
            break;
        }
    }
    __threadfence_system();
}

}  // namespace ee
