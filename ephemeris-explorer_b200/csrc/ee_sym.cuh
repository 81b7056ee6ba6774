// ee_sym.cuh -- throughput acceleration kernel that uses Newton's third law: each unordered pair is evaluated ONCE
// and applied to both bodies (as the reference's loop does, nbody.rs:23-35), which cuts the FP64-pipe work from 16 to
// 10 instructions per directed interaction.
//
// Decomposition.  Bodies are cut into I-tiles of 256*TI (256 threads x TI targets held in registers); the j axis is cut
// into chunks of 32 bodies.  A UNIT is (tile row ti, chunk c) with some j > i in it (c >= ti*tile/32); units are numbered
// canonically (rows ascending, chunks ascending inside a row) and a rank owns a contiguous unit range.  A WORK ITEM is a
// run of consecutive units of one row; the host builds the item table (ee_nbody.cu: build_sym_schedule) with GUIDED sizes:
// long runs first, halving towards the end of the rank's range down to single units, so that the dynamic queue (one
// atomic counter, persistent CTAs) drains with a tail of one 32-body chunk instead of one full item.
//
// Inside an item every warp walks the run chunk by chunk: lane l pairs its TI targets with body (l + k) mod 32 of the
// chunk at rotation k, so at any instant the 32 lanes touch 32 different j -- the j-side partial sums live in a
// warp-private shared-memory array and are updated with plain, conflict-free read-modify-writes (no atomics).  Every
// SBC chunks (one sub-block) the eight warps' arrays are added in warp order and written to part_j[ti][j]; the register
// accumulators go to part_i[slot][i] once per item.  A second, small kernel adds each body's partials in canonical
// order and runs the integrator epilogue.  Every sum has a fixed order => deterministic, whatever the queue did.
//
// Per rotation (TI pairs): TI x (3 sub + 3 r^2 + 6 r^-3 + 1 mu_j*t + 3 FMA_i + 1 mu_i*t + 3 FMA_j) = 20*TI FP64-pipe
// instructions for 2*TI directed interactions (the j-side chain starts from the loaded shared-memory value, so the
// read-modify-write costs no extra add); shared memory: 4 loads (x, y, z, mu of j) + 3 loads + 3 stores.
#pragma once
#include "ee_kernels.cuh"
#include "ee_sym_types.h"

namespace ee {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may be
// made resident while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor has
// completed and its writes are visible, pdl_trigger() lets the successor's CTAs be scheduled as soon as resources free
// up.  Both kernels of a step wait at the very top (everything they read is the predecessor's output), so the only effect
// is that the launch latency of the next kernel is hidden behind the tail of the current one -- which is a third of a
// 4 096-body step.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int NW, int SBC>
struct SymSmem {
    double wacc[NW][3][SBC * 32];
    double sx[NW][32], sy[NW][32], sz[NW][32], sm[NW][32];
};

__device__ __forceinline__ double sym_rcube(double r2) {  // r^-3 from r^2 (see interact_fast)
    const double y0 = rsqrt_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    const double c = y2 * y0;
    return fma(c, q, c);
}

// One 32-body chunk against this lane's TI targets (32 rotations), masked to j > i when DIAG.  Targets are processed G
// at a time with the Newton refinement written stage by stage across the group, so that every instruction has G - 1
// independent neighbours: with 4 warps per scheduler the FP64 pipe (one warp instruction every 2 cycles, dependent-issue
// latency several times that) needs instruction-level parallelism inside a warp, not a serial chain per target.
// Operand bandwidth (measured, tools/fp64_mix_bench.cu, profiles/README.md): an FP64 instruction with three DISTINCT
// register sources occupies the issue path for 3 cycles instead of 2 (DFMA a*b+c with three different registers runs at
// 66 % of the DFMA peak, two-source forms at 99 %).  The six accumulate FMAs of a pair are the only three-source
// instructions here; together with the rest of the mix that caps this loop near 80 % "pipe active" -- which is where
// every schedule tried (4 warps x serial chains, 3 warps x 4-way interleave, asm-pinned triples with operand reuse)
// ends up.
template <int TI, bool DIAG>
__device__ __forceinline__ void sym_chunk(const double* __restrict__ sx, const double* __restrict__ sy,
                                          const double* __restrict__ sz, const double* __restrict__ sm, double* wx, double* wy,
                                          double* wz, int lane, int j0, int ibase, const double (&xi)[TI],
                                          const double (&yi)[TI], const double (&zi)[TI], const double (&mi)[TI],
                                          double (&ax)[TI], double (&ay)[TI], double (&az)[TI]) {
    constexpr int G = TI >= 2 ? 2 : 1;
    int jj = lane;
#pragma unroll 2
    for (int k = 0; k < 32; ++k) {
        const double xj = sx[jj], yj = sy[jj], zj = sz[jj], mj = sm[jj];
        double bx = wx[jj], by = wy[jj], bz = wz[jj];
#pragma unroll
        for (int g = 0; g < TI; g += G) {
            double dx[G], dy[G], dz[G], r2[G], y0[G], y2[G], e[G], pp[G], c[G], rc[G], si[G], sj[G];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                dx[u] = xj - xi[g + u];
                dy[u] = yj - yi[g + u];
                dz[u] = zj - zi[g + u];
            }
#pragma unroll
            for (int u = 0; u < G; ++u) r2[u] = dx[u] * dx[u];
#pragma unroll
            for (int u = 0; u < G; ++u) r2[u] = fma(dy[u], dy[u], r2[u]);
#pragma unroll
            for (int u = 0; u < G; ++u) r2[u] = fma(dz[u], dz[u], r2[u]);
#pragma unroll
            for (int u = 0; u < G; ++u) y0[u] = rsqrt_seed(r2[u]);
#pragma unroll
            for (int u = 0; u < G; ++u) y2[u] = y0[u] * y0[u];
#pragma unroll
            for (int u = 0; u < G; ++u) e[u] = fma(-r2[u], y2[u], 1.0);
#pragma unroll
            for (int u = 0; u < G; ++u) c[u] = y2[u] * y0[u];
#pragma unroll
            for (int u = 0; u < G; ++u) pp[u] = fma(1.875, e[u], 1.5);
#pragma unroll
            for (int u = 0; u < G; ++u) pp[u] = e[u] * pp[u];
#pragma unroll
            for (int u = 0; u < G; ++u) rc[u] = fma(c[u], pp[u], c[u]);
            if (DIAG) {
#pragma unroll
                for (int u = 0; u < G; ++u) rc[u] = (j0 + jj) > (ibase + 32 * (g + u)) ? rc[u] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < G; ++u) {
                si[u] = mj * rc[u];
                sj[u] = mi[g + u] * rc[u];
            }
#pragma unroll
            for (int u = 0; u < G; ++u) {
                ax[g + u] = fma(si[u], dx[u], ax[g + u]);
                ay[g + u] = fma(si[u], dy[u], ay[g + u]);
                az[g + u] = fma(si[u], dz[u], az[g + u]);
            }
#pragma unroll
            for (int u = 0; u < G; ++u) {
                bx = fma(-sj[u], dx[u], bx);
                by = fma(-sj[u], dy[u], by);
                bz = fma(-sj[u], dz[u], bz);
            }
        }
        wx[jj] = bx;
        wy[jj] = by;
        wz[jj] = bz;
        __syncwarp();
        jj = (jj + 1) & 31;
    }
}

// One work item (a run of chunks of one tile row), executed by the whole CTA.  Reserves the CTA's next item from the queue
// when the LAST chunk of this one starts and leaves it in *s_next.
// PROF (developer aid): thread 0 accumulates clock64() spent in item prologue / chunk loop / sub-block merge / i-side store.
template <int TI, int NT, int SBC, bool PROF>
__device__ __forceinline__ void sym_item(SymSmem<NT / 32, SBC>& S, int* s_next, int item, int n, const double4* __restrict__ pm,
                                         const SymItem* __restrict__ items, unsigned* __restrict__ counter,
                                         double* __restrict__ part_i, double* __restrict__ part_j, long long (&pc)[9], long long& tc) {
    constexpr int kSymThreads = NT, kSymWarps = NT / 32;
    constexpr int kTile = kSymThreads * TI;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (PROF) tc = clock64();
    __syncthreads();  // every thread has read s_next
    int fetched = 0;
    const SymItem it = items[item];
    const int ibase = it.ti * kTile + warp * (32 * TI) + lane;
    double xi[TI], yi[TI], zi[TI], mi[TI], ax[TI], ay[TI], az[TI];
#pragma unroll
    for (int t = 0; t < TI; ++t) {
        const double4 p = pm[ibase + 32 * t];
        xi[t] = p.x;
        yi[t] = p.y;
        zi[t] = p.z;
        mi[t] = p.w;
        ax[t] = ay[t] = az[t] = 0.0;
    }
    const int w_lo = it.ti * kTile + warp * (32 * TI);  // this warp's targets are [w_lo, w_lo + 32*TI)
    double4 nxt = pm[it.c0 * 32 + lane];
    if (PROF) {
        asm volatile("" ::"d"(xi[0]), "d"(nxt.x));  // the prologue ends when the loads have landed
        const long long t1 = clock64();
        pc[0] += t1 - tc;
        tc = t1;
        pc[4] += 1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pc[7]));
        pc[8] = it.nc;
    }
    for (int sb0 = 0; sb0 < it.nc; sb0 += SBC) {
        const int sbn = min(SBC, it.nc - sb0);
        for (int k = lane; k < sbn * 32; k += 32) {
            S.wacc[warp][0][k] = 0.0;
            S.wacc[warp][1][k] = 0.0;
            S.wacc[warp][2][k] = 0.0;
        }
        for (int cc = 0; cc < sbn; ++cc) {
            const int c = it.c0 + sb0 + cc;
            const int j0 = c * 32;
            const double4 cur = nxt;
            if (sb0 + cc + 1 < it.nc) nxt = pm[j0 + 32 + lane];  // next chunk's bodies: in flight during this chunk
            // Reserve the next item when the LAST chunk of this one starts: late enough that items are still executed
            // in queue order (reserving at item start let a CTA sit on a long item for the whole duration of its
            // current one -- measured: 102 of 296 CTAs began an 8-chunk item when the rest were drawing single chunks,
            // an 80 us tail on a 400 us share), early enough that the atomic's round trip hides behind a chunk.
            else if (tid == 0) fetched = (int)atomicAdd(counter, 1u);
            if (j0 + 32 <= w_lo) continue;  // diagonal tile: every j of this chunk is below every i of this warp
            S.sx[warp][lane] = cur.x;
            S.sy[warp][lane] = cur.y;
            S.sz[warp][lane] = cur.z;
            S.sm[warp][lane] = cur.w;
            __syncwarp();
            double* wx = &S.wacc[warp][0][cc * 32];
            double* wy = &S.wacc[warp][1][cc * 32];
            double* wz = &S.wacc[warp][2][cc * 32];
            if (j0 < w_lo + 32 * TI)  // some j <= some i of this warp: mask to j > i
                sym_chunk<TI, true>(S.sx[warp], S.sy[warp], S.sz[warp], S.sm[warp], wx, wy, wz, lane, j0, ibase, xi, yi, zi, mi,
                                    ax, ay, az);
            else
                sym_chunk<TI, false>(S.sx[warp], S.sy[warp], S.sz[warp], S.sm[warp], wx, wy, wz, lane, j0, ibase, xi, yi, zi, mi,
                                     ax, ay, az);
        }
        if (PROF) {
            const long long t1 = clock64();
            pc[1] += t1 - tc;
            tc = t1;
        }
        __syncthreads();
        // j side: add the eight warps' arrays in warp order -> part_j[ti][c][j]
        {
            double* pj = part_j + (size_t)it.ti * 3 * n + (size_t)(it.c0 + sb0) * 32;
            const int span = sbn * 32;
            for (int idx = tid; idx < 3 * span; idx += kSymThreads) {
                const int c = idx / span, k = idx - c * span;
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kSymWarps; ++w) s += S.wacc[w][c][k];
                pj[(size_t)c * n + k] = s;
            }
        }
        if (tid == 0 && sb0 + SBC >= it.nc) *s_next = fetched;
        __syncthreads();
        if (PROF) {
            const long long t1 = clock64();
            pc[2] += t1 - tc;
            tc = t1;
        }
    }
    // i side: registers -> part_i[slot][c][local i]
    {
        double* pi = part_i + (size_t)it.slot * 3 * kTile + warp * (32 * TI) + lane;
#pragma unroll
        for (int t = 0; t < TI; ++t) {
            pi[32 * t] = ax[t];
            pi[kTile + 32 * t] = ay[t];
            pi[2 * kTile + 32 * t] = az[t];
        }
    }
    if (PROF) pc[3] += clock64() - tc;
}

// TI targets per lane, NT threads per CTA (tile = NT*TI bodies), MINB resident CTAs per SM, SBC chunks per shared-memory
// sub-block.
// PROF (developer aid): thread 0 of every CTA writes the four phase totals + item count, then (globaltimer ns) loop start,
// loop end, start and size of its last item to prof[blockIdx.x][9].
template <int TI, int NT, int MINB, int SBC, bool PROF = false>
__global__ void __launch_bounds__(NT, MINB) k_accel_sym(int n, const double4* __restrict__ pm,
                                                                 const SymItem* __restrict__ items, int n_items,
                                                                 unsigned* __restrict__ counter, double* __restrict__ part_i,
                                                                 double* __restrict__ part_j, long long* __restrict__ prof = nullptr) {
    extern __shared__ __align__(16) unsigned char sym_raw[];
    SymSmem<NT / 32, SBC>& S = *reinterpret_cast<SymSmem<NT / 32, SBC>*>(sym_raw);
    __shared__ int s_next;
    long long pc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tc = 0;
    pdl_wait();
    pdl_trigger();
    if (PROF) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pc[5]));
    // The first item of a CTA is its own index (the queue counter starts at gridDim.x, see k_sym_reduce): a launch of
    // thousands of one-warp CTAs would otherwise begin with thousands of atomics on one address.
    int item = (int)blockIdx.x;
    while (item < n_items) {
        sym_item<TI, NT, SBC, PROF>(S, &s_next, item, n, pm, items, counter, part_i, part_j, pc, tc);
        item = s_next;
    }
    if (PROF && threadIdx.x == 0 && prof) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pc[6]));
        for (int q = 0; q < 9; ++q) prof[blockIdx.x * 9 + q] = pc[q];
    }
}

// Adds body b's partials and runs the epilogue; also re-arms the item queue for the next launch.  A body's partials are
// the i-side slots of its tile row (this rank's items there, j ascending) and one j-side entry per tile row at or above
// its own whose unit this rank owns.  Eight threads share a body: thread w adds the slots / rows congruent to w modulo
// 8 in ascending order, then the sub-sums are added in the order w = 0..7 -- a fixed order, so the result does not
// depend on how the queue was drained, and a row made of hundreds of single-chunk items (the guided tail, or a rank's
// share of a sharded run) is not a serial chain of hundreds of dependent loads.
// KL threads share a body, KB bodies per CTA.  Large systems: 8 x 32 (one coalesced 256-byte segment per load; 16 lanes
// measured slower there).  Mid-size systems (warp-sized tiles, a few thousand bodies): 32 x 8 -- the grid would otherwise
// be a fraction of a wave and the longest row (hundreds of single-unit items) a serial chain of dependent loads.
// The reduce of KB consecutive bodies (block vb) by KL * KB threads.
template <int kTile, int KL, int KB>
__device__ __forceinline__ void sym_reduce_block(int vb, int n, const SymShare& sh, const int* __restrict__ row_slot,
                                                 const double* __restrict__ part_i, const double* __restrict__ part_j,
                                                 const EpArgs& ep, double (&red)[3][KL][KB]) {
    const int l = threadIdx.x % KB, w = threadIdx.x / KB;
    const int b = vb * KB + l;
    if (w == 0 && b < n && ep.kind == EP_QT) {
        // the epilogue's history reads do not depend on the sums: start them now, they land in L1 while the partials are added
        const QtArgs& q = ep.qt;
        for (int j = 0; j < q.order; ++j) {
            const double* a = ep.ra + (size_t)q.slot[j] * 3 * n + b;
            if (j > 0) {
                prefetch_l1(a);
                prefetch_l1(a + n);
                prefetch_l1(a + 2 * (size_t)n);
            }
            if (j <= 1 || q.nalpha[j] != 0.0) prefetch_l1(ep.ry + (size_t)q.slot[j] * n + b);
        }
    }
    auto ld = [](const double* p) { return *p; };
    double sx = 0.0, sy = 0.0, sz = 0.0;
    if (b < n) {
        const int tb = b / kTile, lb = b - tb * kTile, cb = b >> 5;
        {   // i side: this rank's items of row tb occupy slots [row_slot[tb], row_slot[tb+1])
            const int s0 = row_slot[tb], s1 = row_slot[tb + 1];
            const double* p = part_i + (size_t)(s0 + w) * 3 * kTile + lb;
#pragma unroll 4
            for (int s = s0 + w; s < s1; s += KL, p += (size_t)KL * 3 * kTile) {  // loads run ahead of the adds
                sx += ld(p);
                sy += ld(p + kTile);
                sz += ld(p + 2 * kTile);
            }
        }
        const long long cpt = kTile / 32;
        for (int ti = w; ti <= tb; ti += KL) {  // j side: one candidate unit per tile row at or above this body's row
            const long long u = sym_row_unit(ti, sh.nch, cpt) + (cb - (long long)ti * cpt);
            if (u < sh.u_lo || u >= sh.u_hi) continue;
            const double* p = part_j + (size_t)ti * 3 * n + b;
            sx += ld(p);
            sy += ld(p + n);
            sz += ld(p + 2 * (size_t)n);
        }
    }
    red[0][w][l] = sx;
    red[1][w][l] = sy;
    red[2][w][l] = sz;
    __syncthreads();
    if (KL > 8) {  // first level: lane group g adds sub-sums 8g .. 8g+7 in order
        if (w < KL / 8) {
            sx = sy = sz = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                sx += red[0][w * 8 + q][l];
                sy += red[1][w * 8 + q][l];
                sz += red[2][w * 8 + q][l];
            }
        }
        __syncthreads();
        if (w < KL / 8) {
            red[0][w][l] = sx;
            red[1][w][l] = sy;
            red[2][w][l] = sz;
        }
        __syncthreads();
    }
    if (w == 0 && b < n) {
        constexpr int kTop = KL > 8 ? KL / 8 : KL;
        sx = sy = sz = 0.0;
#pragma unroll
        for (int q = 0; q < kTop; ++q) {
            sx += red[0][q][l];
            sy += red[1][q][l];
            sz += red[2][q][l];
        }
        apply_epilogue<false>(ep, b, D3{sx, sy, sz});
    }
}

template <int kTile, int KL, int KB>
__global__ void __launch_bounds__(KL * KB) k_sym_reduce(int n, SymShare sh, const int* __restrict__ row_slot,
                                                       const double* __restrict__ part_i, const double* __restrict__ part_j,
                                                       unsigned* __restrict__ counter, unsigned queue_start, EpArgs ep) {
    __shared__ double red[3][KL][KB];
    pdl_wait();
    pdl_trigger();
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = queue_start;  // items below it are taken by CTA index
    sym_reduce_block<kTile, KL, KB>((int)blockIdx.x, n, sh, row_slot, part_i, part_j, ep, red);
}

// ---------------------------------------------------------------------------------------------------------
// Multi-GPU without NCCL on the data path: every rank evaluates its share of the pair units and reduces them locally
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

// Flag barrier across the GPUs of one node, callable by one warp: lane q < world stores `epoch` into slot `rank` of peer
// q's flag array, then waits until slot q of its own array reaches `epoch`.  Epochs only grow, so nothing is ever reset.
// A spin that exceeds `timeout` cycles records a sticky error instead of hanging the GPU: in device memory (what the
// kernels of later steps look at -- a read of host memory per thread would cost a PCIe round trip each) and in mapped
// host memory (what the host looks at without synchronising).
__device__ __forceinline__ bool peer_barrier_warp(const PeerTable& T, unsigned long long epoch, volatile int* err,
                                                  volatile int* err_host, long long timeout) {
    const int q = threadIdx.x & 31;
    bool ok = true;
    if (q < T.world) {
        __threadfence_system();
        volatile unsigned long long* out = T.flags[q] + T.rank;
        *out = epoch;
        volatile unsigned long long* in = T.flags[T.rank] + q;
        const long long t0 = clock64();
        while (*in < epoch) {
            if (*err || clock64() - t0 > timeout) {
                *err = 1;
                *err_host = 1;
                ok = false;
                break;
            }
        }
        __threadfence_system();
    }
    return __all_sync(0xffffffffu, ok);
}

__global__ void k_peer_barrier(PeerTable T, unsigned long long epoch, int* err, int* err_host, long long timeout) {
    peer_barrier_warp(T, epoch, err, err_host, timeout);
}

__global__ void k_peer_finish(int n, int b0, int b1, PeerTable T, EpArgs ep, const int* err) {
    const int b = b0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= b1) return;
    if (*(volatile const int*)err) return;  // a barrier timed out: peer data is stale, leave the state untouched
    double px[kMaxPeers], py[kMaxPeers], pz[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {  // all (remote) loads first, then the ordered sum
        if (q < T.world) {
            const double* p = T.a_part[q];
            px[q] = p[b];
            py[q] = p[n + b];
            pz[q] = p[2 * (size_t)n + b];
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

}  // namespace ee
