// ee_nbody.cu -- host side of the n-body propagator (generic multi-CTA back end) + launch logic.
//
// Mirrors, in order of the reference call stack:
//   NBodyPropagator::{new, step}                     ephemeris/src/propagators/nbody.rs:93-121, :200-207
//   LinearMultistepIntegrator::advance               integration/src/multistep/mod.rs:194-225
//   ELM2::{from_problem, advance, advance_with}      integration/src/multistep/second_order/mod.rs:74-153
//   SubstepperIntegrator<4>::advance                 integration/src/multistep/mod.rs:97-108
//   FixedRungeKuttaIntegrator::advance               integration/src/runge_kutta/mod.rs:106-126
//   SRKN<BlanesMoan6B>::advance                      integration/src/runge_kutta/nystrom/symplectic.rs:70-102
//   SplineInterpolators::{new_solution, solout}      ephemeris/src/propagators/nbody.rs:372-489
//
// Device layout (HBM): positions live as double4 (x, y, z, mu) so one coalesced 32-byte load feeds the pair loop;
// the multistep history is a ring of R = order+1 slots indexed by absolute step number (slot = step % R):
//   ry[R][n] double4, ra[R][3][n] double (SoA), dy[3][n].  A steady-state step is one fused launch:
//   a_{s} = accel(y_s)  ->  dy_s (Cowell)  ->  y_{s+1} (predictor), written to the next ring slot.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>

#include "ee_coeffs.h"
#include "ee_engine.h"
#include "ee_kernels.cuh"
#include "ee_sym.cuh"

namespace ee {

constexpr int64_t kSmallBatch = 4096;  // run-ahead batch of the persistent small-system kernel (~5 ms of device work at 32 bodies)
constexpr long long kPeerTimeoutCycles = 60000000000LL;  // ~30 s at 1.97 GHz: a rank that is merely slow is not an error

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
std::atomic<int> g_pair_variant{0};  // ee_set_pair_variant: reading of particular's pair kernel used by handles created afterwards

namespace {
struct DeviceInfo {
    std::once_flag once;
    cudaStream_t pool = nullptr;
    int sm_count = 0;
    int occ_fast256 = 0, occ_fast128 = 0;  // resident CTAs per SM of the plain throughput kernel (filled on demand)
};
DeviceInfo g_dev[64];

DeviceInfo& device_info(int device) {
    if (device < 0 || device >= 64) throw Error(EE_ERR_INVALID, "device ordinal out of range");
    DeviceInfo& d = g_dev[device];
    std::call_once(d.once, [&] {
        int cur = -1;
        EE_CUDA(cudaGetDevice(&cur));
        EE_CUDA(cudaSetDevice(device));
        EE_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
        EE_CUDA(cudaStreamCreateWithFlags(&d.pool, cudaStreamNonBlocking));
        cudaMemPool_t mp;
        EE_CUDA(cudaDeviceGetDefaultMemPool(&mp, device));
        uint64_t keep = ~0ull;  // never hand freed blocks back to the driver: allocations become pointer bumps
        EE_CUDA(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep));
        if (cur >= 0 && cur != device) EE_CUDA(cudaSetDevice(cur));
    });
    return d;
}
}  // namespace

cudaStream_t pool_stream(int device) { return device_info(device).pool; }

// ---- NCCL through dlopen: single-GPU use never touches it; under torchrun the already-loaded libnccl is reused
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) throw Error(EE_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
#define EE_SYM(field, name)                                                   \
    api.field = (decltype(api.field))dlsym(api.lib, name);                    \
    if (!api.field) throw Error(EE_ERR_NCCL, std::string("missing symbol ") + name)
    EE_SYM(GetUniqueId, "ncclGetUniqueId");
    EE_SYM(CommInitRank, "ncclCommInitRank");
    EE_SYM(CommDestroy, "ncclCommDestroy");
    EE_SYM(AllReduce, "ncclAllReduce");
    EE_SYM(AllGather, "ncclAllGather");
    EE_SYM(GetErrorString, "ncclGetErrorString");
#undef EE_SYM
    return api;
}
#define EE_NCCL(expr)                                                                                       \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess)                                                                              \
            throw Error(EE_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(_r));                \
    } while (0)
}  // namespace

void nccl_unique_id(void* out128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    EE_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, 128);
}

// ------------------------------------------------------------------------------------------------------------
struct MethodTable {
    int order;
    const double *nalpha, *beta, *cow;
    double inv_beta_d, cow_inv_beta_d;
};
static MethodTable method_table(int method) {
    if (method == EE_QUINLAN_TREMAINE_12)
        return {12, EE_QT12_NEG_ALPHA, EE_QT12_BETA, EE_COWELL12_BETA, EE_QT12_INV_BETA_D, EE_COWELL12_INV_BETA_D};
    if (method == EE_STORMER_13)
        return {13, EE_ST13_NEG_ALPHA, EE_ST13_BETA, EE_COWELL13_BETA, EE_ST13_INV_BETA_D, EE_COWELL13_INV_BETA_D};
    if (method == EE_BLANES_MOAN_14A)  // a symplectic Runge-Kutta-Nystrom method: no multistep tables, one "history" slot
        return {1, nullptr, nullptr, nullptr, 0.0, 0.0};
    throw Error(EE_ERR_INVALID, "unknown method id (12 = QuinlanTremaine12, 13 = Stormer13, 14 = BlanesMoan14A)");
}

NBodyEngine::NBodyEngine(int64_t n_, const double* pos, const double* vel, const double* mus, double t0, double h_signed,
                         int method_, int mode_, int device_, int rank_, int world_, const void* uid, int exchange_)
    : n(n_), method(method_), mode(mode_), device(device_), rank(rank_), world(world_), exchange(exchange_) {
    try {
        init(pos, vel, mus, t0, h_signed, uid);
    } catch (...) {  // the destructor does not run for a half-built object: give back the stream, events and communicator
        release_all();
        throw;
    }
}

void NBodyEngine::init(const double* pos, const double* vel, const double* mus, double t0, double h_signed, const void* uid) {
    EE_REQUIRE(n >= 1, "n must be >= 1");
    const bool blank = !pos && !vel && !mus;  // clone(): the state arrives by device-to-device copies
    EE_REQUIRE(blank || (pos && vel && mus), "null input array");
    EE_REQUIRE(mode == EE_MODE_PARITY || mode == EE_MODE_THROUGHPUT, "unknown mode");
    EE_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    MethodTable mt = method_table(method);
    order = mt.order;
    R = order + 1;
    srkn_main = method == EE_BLANES_MOAN_14A;
    pair_variant = g_pair_variant.load();
    h = h_signed;
    hs = h_signed * (1.0 / 4.0);  // Substepper::new -- multistep/mod.rs:53-58
    t = t0;
    bound = std::numeric_limits<double>::infinity();  // nbody.rs:112
    int ndev = 0;
    EE_CUDA(cudaGetDeviceCount(&ndev));
    EE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU path)");
    EE_CUDA(cudaSetDevice(device));
    sm_count = device_info(device).sm_count;  // cudaGetDeviceProperties costs milliseconds: cached per device
    EE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    EE_CUDA(cudaEventCreate(&ev0));
    EE_CUDA(cudaEventCreate(&ev1));
    if (world > 1) {
        EE_REQUIRE(uid, "sharded create needs a NCCL unique id");
        EE_REQUIRE(n % world == 0, "n must be divisible by the number of ranks");
        if (exchange == EE_EXCHANGE_ALLREDUCE && mode == EE_MODE_PARITY)
            throw Error(EE_ERR_UNSUPPORTED, "parity mode cannot be source-sharded (summation order changes); use allgather");
        ncclUniqueId id;
        std::memcpy(&id, uid, 128);
        ncclComm_t c = nullptr;
        EE_NCCL(nccl().CommInitRank(&c, world, id, rank));
        comm = (void*)c;
    }
    const int64_t per = n / world;
    if (world > 1 && exchange == EE_EXCHANGE_ALLGATHER) {
        i0 = rank * per;
        i1 = i0 + per;
        j0 = 0;
        j1 = n;
    } else if (world > 1) {
        i0 = 0;
        i1 = n;
        j0 = rank * per;
        j1 = j0 + per;
    } else {
        i0 = 0;
        i1 = n;
        j0 = 0;
        j1 = n;
    }
    const bool ipc = world > 1;  // the NVLink peer path exports these buffers as CUDA-IPC handles
    ry.alloc((size_t)R * n, ipc);
    ra.alloc((size_t)R * 3 * n, ipc);
    dy.alloc((size_t)3 * n, ipc);
    EE_CUDA(cudaMemsetAsync(ry.p, 0, ry.bytes(), stream));
    EE_CUDA(cudaMemsetAsync(ra.p, 0, ra.bytes(), stream));
    if (!blank) {
        std::vector<double4> hp((size_t)n);
        std::vector<double> hv((size_t)3 * n);
        for (int64_t k = 0; k < n; ++k) {
            hp[(size_t)k] = make_double4(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], mus[k]);
            for (int c = 0; c < 3; ++c) hv[(size_t)(c * n + k)] = vel[3 * k + c];
        }
        EE_CUDA(cudaMemcpyAsync(ry.p, hp.data(), (size_t)n * sizeof(double4), cudaMemcpyHostToDevice, stream));
        EE_CUDA(cudaMemcpyAsync(dy.p, hv.data(), hv.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    EE_CUDA(cudaStreamSynchronize(stream));
    plan_launch();
}

NBodyEngine::~NBodyEngine() {
    // pooled buffers are freed stream-ordered on the pool's stream: nothing of ours may still be running on them
    if (stream) cudaStreamSynchronize(stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    release_all();
}

void NBodyEngine::release_all() {
    for (auto& e : p2p_ev) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    for (int k = 0; k < 2; ++k) {
        if (stage_packed[k]) cudaEventDestroy(stage_packed[k]);
        if (stage_copied[k]) cudaEventDestroy(stage_copied[k]);
        stage_packed[k] = stage_copied[k] = nullptr;
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    copy_stream = nullptr;
    for (cudaEvent_t& ev : batch_ev) {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
    }
    for (void* p : p2p_opened) cudaIpcCloseMemHandle(p);
    p2p_opened.clear();
    delete (PeerTable*)p2p_table;
    p2p_table = nullptr;
    if (p2p_err_h) cudaFreeHost(p2p_err_h);
    p2p_err_h = nullptr;
    if (comm) nccl().CommDestroy((ncclComm_t)comm);
    comm = nullptr;
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    ev0 = ev1 = nullptr;
    stream = nullptr;
}

// Choose the throughput-kernel decomposition: tiles x splits blocks, with the block count a whole number of waves
// (sm_count x resident CTAs per SM) whenever the problem is big enough to allow it.
void NBodyEngine::plan_launch() {
    // Large systems take the pair-symmetric kernel (ee_sym.cuh): single GPU, or pair units sharded + allreduce / peer path.
    const bool dev_aids = getenv("EE_DEV_AIDS") && getenv("EE_DEV_AIDS")[0] == '1';  // developer switches below need this
    const char* env = dev_aids ? getenv("EE_SYM") : nullptr;
    const bool sym_allowed = !(env && env[0] == '0');
    sym_ti = 4;
    sym_nt = 256;
    sym_minb = 2;
    sym_sbc = 16;
    if (dev_aids) {
        if (const char* v = getenv("EE_SYM_VARIANT")) {  // "TI,NT,MINB,SBC"
            int a = 0, b = 0, c = 0, d = 0;
            EE_REQUIRE(sscanf(v, "%d,%d,%d,%d", &a, &b, &c, &d) == 4, "EE_SYM_VARIANT must be TI,NT,MINB,SBC");
            sym_ti = a;
            sym_nt = b;
            sym_minb = c;
            sym_sbc = d;
        }
    }
    // Mid-size systems (a few thousand bodies: BASELINE.json configs[2]) hold too few pairs for 1 024-body tiles -- 4 096
    // bodies are 320 (tile, chunk) units for 296 CTAs, and every warp's share is about ONE 128 x 32 block of pairs, so
    // the step is launch ramp + one block + reduce.  Measured (profiles/r02/mid_probe_*.jsonl): 4 CTAs of 4 warps per SM
    // with 512-body tiles is the fastest shape from 4 096 to 16 384 bodies (fewer CTAs to launch than one-warp CTAs, finer
    // than 1 024-body tiles); sizes that are a multiple of 128 but not of 512 take one-warp CTAs with 128-body tiles.
    const bool mid = n < 32768 || n % 1024 != 0;
    if (mid && !(dev_aids && getenv("EE_SYM_VARIANT"))) {
        if (n % 512 == 0) {
            sym_nt = 128;
            sym_minb = 4;
        } else {
            sym_nt = 32;
            sym_minb = 16;
        }
        sym_sbc = 4;
    }
    const int tile = sym_nt * sym_ti;
    use_sym = sym_allowed && mode == EE_MODE_THROUGHPUT && n % tile == 0 && n >= 2048 && n < (1ll << 30) &&
              (tile >= 1024 || n <= 65536) &&  // j-side partials are (n / tile) x 3n doubles: small tiles only for mid sizes
              (world == 1 || exchange == EE_EXCHANGE_ALLREDUCE);
    if (use_sym) {
        int share_rank = rank, share_world = world;
        if (dev_aids && world == 1) {
            // developer aid "a/b": behave like rank a of b on one GPU (partial sums of that share only) -- lets a one-GPU box
            // time and check one rank's share of a sharded run
            if (const char* rg = getenv("EE_SYM_RANGE")) {
                int a = 0, b = 1;
                if (sscanf(rg, "%d/%d", &a, &b) == 2 && b >= 1 && a >= 0 && a < b) {
                    share_rank = a;
                    share_world = b;
                }
            }
        }
        int guided_max = sym_sbc;  // longest item = one shared-memory sub-block (measured: longer items are not faster)
        if (dev_aids)
            if (const char* g = getenv("EE_SYM_MAXC")) guided_max = std::max(1, atoi(g));
        int guided = 1;  // items per CTA the remaining work is dealt into (measured: 1 beats 2 -- fewer, larger tail items)
        if (dev_aids)
            if (const char* g = getenv("EE_SYM_GUIDED")) guided = std::max(1, atoi(g));
        SymSchedule sc = build_sym_schedule(n, tile, guided * sym_minb * sm_count, share_world, share_rank, guided_max);
        sym_n_items = (int)sc.items.size();
        sym_share = SymShare{sc.u_lo, sc.u_hi, tile, (int)(n / tile), (int)(n / 32)};
        sym_items.alloc(std::max<size_t>(1, sc.items.size()));
        sym_row_slot.alloc(sc.row_slot.size());
        EE_CUDA(cudaMemcpy(sym_items.p, sc.items.data(), sc.items.size() * sizeof(SymItem), cudaMemcpyHostToDevice));
        EE_CUDA(cudaMemcpy(sym_row_slot.p, sc.row_slot.data(), sc.row_slot.size() * sizeof(int), cudaMemcpyHostToDevice));
        sym_part_i_count = std::max<size_t>(1, sc.items.size()) * 3 * tile;  // allocated by ensure_scratch()
        sym_part_j_count = (size_t)(n / tile) * 3 * n;
        sym_counter.alloc(1);
        const unsigned q0 = (unsigned)(sym_minb * sm_count);  // the first gridDim.x items are taken by CTA index (k_accel_sym)
        EE_CUDA(cudaMemcpy(sym_counter.p, &q0, sizeof(unsigned), cudaMemcpyHostToDevice));
    }
    block = n >= 16384 ? 256 : 128;
    const int64_t targets = i1 - i0, sources = j1 - j0;
    tiles = (int)((targets + block - 1) / block);
    DeviceInfo& di = device_info(device);
    int& occ_ref = block == 256 ? di.occ_fast256 : di.occ_fast128;
    if (occ_ref == 0) {  // benign race: every thread computes the same value
        int q = 0;
        if (block == 256)
            EE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_accel_fast<256>, 256, 0));
        else
            EE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_accel_fast<128>, 128, 0));
        occ_ref = std::max(1, q);
    }
    const int occ = occ_ref;
    const int64_t slots = (int64_t)sm_count * occ;
    const int64_t max_splits = std::max<int64_t>(1, std::min<int64_t>(256, sources / (2 * block)));
    double best = 1e300;
    int best_s = 1;
    for (int64_t s = 1; s <= max_splits; ++s) {
        const int64_t blocks = (int64_t)tiles * s;
        const int64_t waves = (blocks + slots - 1) / slots;
        const double cost = (double)waves / (double)s * (1.0 + 0.002 * (double)s);  // mild penalty on partial traffic
        if (cost < best - 1e-12) {
            best = cost;
            best_s = (int)s;
        }
    }
    chunk = (sources + best_s - 1) / best_s;
    splits = (int)((sources + chunk - 1) / chunk);
}

// Scratch of the acceleration kernels and the starter: allocated when the handle first needs it, so that a clone kept
// only as a snapshot (prediction.rs:224-229) costs no more than the state it holds.
void NBodyEngine::ensure_scratch() {
    if (scratch_ready) return;
    ytmp[0].alloc((size_t)n);
    ytmp[1].alloc((size_t)n);
    a_scr.alloc((size_t)3 * n, world > 1);
    if (use_sym) {
        sym_part_i.alloc(sym_part_i_count);
        sym_part_j.alloc(sym_part_j_count);
    } else if (splits > 1) {
        part.alloc((size_t)splits * 3 * n);
        tickets.alloc((size_t)tiles);
        EE_CUDA(cudaMemsetAsync(tickets.p, 0, tickets.bytes(), stream));
    }
    scratch_ready = true;
}

void NBodyEngine::exchange_y(double4* buf) {
    if (world == 1 || exchange != EE_EXCHANGE_ALLGATHER) return;
    const int64_t per = n / world;
    EE_NCCL(nccl().AllGather((const void*)(buf + i0), (void*)buf, (size_t)per * 4, ncclDouble, (ncclComm_t)comm, stream));
}

// One acceleration evaluation of the positions in `y_in`, followed by epilogue `ep` for every local target.
void NBodyEngine::accel(const double4* y_in, EpArgs ep) {
    ensure_scratch();
    ep.n = n;
    const bool reduce = world > 1 && exchange == EE_EXCHANGE_ALLREDUCE;
    EpArgs kep = ep;
    if (reduce) {  // partial accelerations only; the epilogue runs after the collective
        kep = EpArgs{};
        kep.kind = EP_STORE;
        kep.n = n;
        kep.a_out = a_scr.p;
    }
    if (use_sym) {
        launch_sym(y_in, kep);
    } else {
        if (mode == EE_MODE_PARITY) {
            constexpr int B = 128;
            const int grid = (int)((i1 - i0 + B - 1) / B);
            k_accel_parity<B><<<grid, B, 0, stream>>>(n, i0, i1, y_in, kep, pair_variant);
        } else {
            dim3 grid((unsigned)tiles, (unsigned)splits);
            if (block == 256)
                k_accel_fast<256><<<grid, 256, 0, stream>>>(n, i0, i1, j0, j1, chunk, splits, y_in, part.p, tickets.p, kep);
            else
                k_accel_fast<128><<<grid, 128, 0, stream>>>(n, i0, i1, j0, j1, chunk, splits, y_in, part.p, tickets.p, kep);
        }
        EE_CUDA(cudaGetLastError());
        count_launch();
        accel_launches++;
    }
    if (reduce) {
        EE_NCCL(nccl().AllReduce(a_scr.p, a_scr.p, (size_t)3 * n, ncclDouble, ncclSum, (ncclComm_t)comm, stream));
        epilogue_from(a_scr.p, ep);
    }
}

void NBodyEngine::epilogue_from(const double* a_in, EpArgs ep) {
    ep.n = n;
    const int B = 128;
    const int grid = (int)((i1 - i0 + B - 1) / B);
    if (mode == EE_MODE_PARITY)
        k_epilogue<true><<<grid, B, 0, stream>>>(i0, i1, a_in, ep);
    else
        k_epilogue<false><<<grid, B, 0, stream>>>(i0, i1, a_in, ep);
    EE_CUDA(cudaGetLastError());
    count_launch();
}

QtArgs NBodyEngine::qt_args(int64_t newest, int64_t next) const {
    MethodTable mt = method_table(method);
    QtArgs q{};
    q.order = order;
    for (int j = 0; j < order; ++j) {
        q.slot[j] = slot_of(newest - j);
        q.nalpha[j] = 1.0 * mt.nalpha[j];  // P::Time::one() * Ratio::from_int(..) -- second_order/mod.rs:106-107
        q.beta[j] = 1.0 * mt.beta[j];
        q.cow[j] = 1.0 * mt.cow[j];        // cowell.rs:40
    }
    q.slot_next = slot_of(next);
    q.f = h * h * mt.inv_beta_d;           // second_order/mod.rs:120
    q.g = h * mt.cow_inv_beta_d;           // cowell.rs:51
    q.h = h;
    return q;
}

void NBodyEngine::ensure_a0() {
    if (have_a0) return;
    EpArgs ep{};
    ep.kind = EP_STORE;
    ep.a_out = ra.p + (size_t)slot_of(0) * 3 * n;
    exchange_y(ry.p + (size_t)slot_of(0) * n);  // no-op unless sharded by targets (every rank already has all of y_0)
    accel(ry.p + (size_t)slot_of(0) * n, ep);
    have_a0 = true;
}

// `substeps` x SRKN<C>::advance(h_sub) (runge_kutta/nystrom/symplectic.rs:70-102) behind FixedRungeKuttaIntegrator's guards
// (runge_kutta/mod.rs:106-126), every stage one fused launch: acceleration + kick + drift (EP_KD).  Used two ways:
//   * C = BlanesMoan6B, 4 substeps of h/4: one call of LinearMultistepIntegrator::advance while the multistep start-up is
//     active (multistep/mod.rs:97-108); the acceleration at the new state is the starter's own last-stage evaluation
//     (A[6] = 0 leaves y untouched after it);
//   * C = BlanesMoan14A, 1 substep of h: the fixed-step method itself (methods.rs:1730-1774).  FSAL with B[0] = 0: stage 0
//     re-uses the previous step's last evaluation, multiplied by an exact zero.
int32_t NBodyEngine::srkn_step(int stages, const double* CA, const double* CB, int substeps, double h_sub) {
    ensure_scratch();
    ensure_a0();
    const double4* y_in = ry.p + (size_t)slot_of(m) * n;
    const double* a_in = ra.p + (size_t)slot_of(m) * 3 * n;
    int pp = 0;
    for (int sub = 0; sub < substeps; ++sub) {
        if (t >= bound) return EE_BOUND_REACHED;             // runge_kutta/mod.rs:113-115
        if (t + h_sub == t) return EE_STEP_SIZE_UNDERFLOW;   // :117-119
        for (int s = 0; s < stages; ++s) {
            const bool final_stage = sub == substeps - 1 && s == stages - 1;
            double4* y_out = final_stage ? ry.p + (size_t)slot_of(m + 1) * n : ytmp[pp].p;
            EpArgs ep{};
            ep.kind = EP_KD;
            ep.kd.hb = h_sub * CB[s];
            ep.kd.ha = h_sub * CA[s];
            ep.y_in = y_in;
            ep.y_out = y_out;
            ep.dy = dy.p;
            if (s == 0) {  // FSAL stage: reuse the previous evaluation (symplectic.rs:75)
                ep.a_out = const_cast<double*>(a_in);
                epilogue_from(a_in, ep);
            } else {
                ep.a_out = final_stage ? ra.p + (size_t)slot_of(m + 1) * 3 * n : a_scr.p;
                if (world > 1 && exchange == EE_EXCHANGE_ALLREDUCE && !final_stage) ep.a_out = a_scr.p;
                accel(y_in, ep);
                a_in = ep.a_out;
            }
            exchange_y(y_out);
            y_in = y_out;
            pp ^= 1;
        }
        t = t + h_sub;  // symplectic.rs:98
    }
    m += 1;
    predicted = false;
    return EE_OK;
}

int32_t NBodyEngine::starter_step() { return srkn_step(EE_BM6B_STAGES, EE_BM6B_A, EE_BM6B_B, 4, hs); }

struct P2PBlob {
    cudaIpcMemHandle_t a_part, ry, flags, ra, dy, unused[3];
};
static_assert(sizeof(P2PBlob) == 512, "8 IPC handle slots of 64 bytes");

void NBodyEngine::p2p_export(void* blob256 /* 512 bytes */) {
    EE_REQUIRE(world > 1 && use_sym && exchange == EE_EXCHANGE_ALLREDUCE,
               "the peer path needs a sharded handle on the pair-symmetric kernel (throughput mode, allreduce layout, n >= 32768)");
    EE_REQUIRE(world <= kMaxPeers, "at most 8 peers");
    EE_CUDA(cudaSetDevice(device));
    ensure_scratch();
    if (!p2p_flags.p) {
        p2p_flags.alloc(kMaxPeers, true);
        EE_CUDA(cudaMemset(p2p_flags.p, 0, p2p_flags.bytes()));
        p2p_err_dev.alloc(1);
        EE_CUDA(cudaMemset(p2p_err_dev.p, 0, sizeof(int)));
        EE_CUDA(cudaHostAlloc((void**)&p2p_err_h, sizeof(int), cudaHostAllocMapped));
        *p2p_err_h = 0;
        EE_CUDA(cudaHostGetDevicePointer((void**)&p2p_err_d, p2p_err_h, 0));
    }
    P2PBlob b;
    std::memset(&b, 0, sizeof(b));
    EE_CUDA(cudaIpcGetMemHandle(&b.a_part, a_scr.p));
    EE_CUDA(cudaIpcGetMemHandle(&b.ry, ry.p));
    EE_CUDA(cudaIpcGetMemHandle(&b.flags, p2p_flags.p));
    EE_CUDA(cudaIpcGetMemHandle(&b.ra, ra.p));
    EE_CUDA(cudaIpcGetMemHandle(&b.dy, dy.p));
    std::memcpy(blob256, &b, sizeof(b));
}

void NBodyEngine::p2p_connect(const void* all_blobs) {
    EE_REQUIRE(p2p_flags.p, "call p2p_export on every rank first");
    EE_CUDA(cudaSetDevice(device));
    std::unique_ptr<PeerTable> owner(new PeerTable());  // released into p2p_table only once every mapping is open
    PeerTable* T = owner.get();
    T->world = world;
    T->rank = rank;
    for (int q = 0; q < world; ++q) {
        if (q == rank) {
            T->a_part[q] = a_scr.p;
            T->ry[q] = ry.p;
            T->ra[q] = ra.p;
            T->dy[q] = dy.p;
            T->flags[q] = p2p_flags.p;
            continue;
        }
        const P2PBlob* b = (const P2PBlob*)all_blobs + q;
        void *pa, *pr, *pf, *pra, *pdy;
        EE_CUDA(cudaIpcOpenMemHandle(&pa, b->a_part, cudaIpcMemLazyEnablePeerAccess));
        EE_CUDA(cudaIpcOpenMemHandle(&pr, b->ry, cudaIpcMemLazyEnablePeerAccess));
        EE_CUDA(cudaIpcOpenMemHandle(&pf, b->flags, cudaIpcMemLazyEnablePeerAccess));
        EE_CUDA(cudaIpcOpenMemHandle(&pra, b->ra, cudaIpcMemLazyEnablePeerAccess));
        EE_CUDA(cudaIpcOpenMemHandle(&pdy, b->dy, cudaIpcMemLazyEnablePeerAccess));
        p2p_opened.insert(p2p_opened.end(), {pa, pr, pf, pra, pdy});
        T->ra[q] = (double*)pra;
        T->dy[q] = (double*)pdy;
        T->a_part[q] = (const double*)pa;
        T->ry[q] = (double4*)pr;
        T->flags[q] = (unsigned long long*)pf;
    }
    p2p_table = owner.release();
    p2p_ready = true;
}

// The peer barriers record a timeout in pinned host memory (sticky): every observer and every step call looks at it,
// so a rank that did not arrive turns into an error instead of a silently stale trajectory.
void NBodyEngine::check_async_error() {
    if (!p2p_err_h) return;
    if (*(volatile int*)p2p_err_h) throw Error(EE_ERR_CUDA, "peer barrier timed out (a rank did not arrive); the handle is dead");
}

// one steady-state step over NVLink peer memory (no NCCL):
//   pair items -> local reduce to a_part -> barrier -> slice finish (peer loads + epilogue + peer stores) -> barrier
void NBodyEngine::p2p_step(const EpArgs& ep_in) {
    const PeerTable& T = *(const PeerTable*)p2p_table;
    EpArgs ep = ep_in;
    ep.n = n;
    const double4* y_in = ry.p + (size_t)ep.qt.slot[0] * n;
    EpArgs store{};
    store.kind = EP_STORE;
    store.n = n;
    store.a_out = a_scr.p;
    const int64_t per = n / world, b0 = rank * per, b1 = b0 + per;
    const unsigned fg = (unsigned)((per + 127) / 128);
    // optional event trace of the five launches (ee_nbody_p2p_trace): where a sharded step spends its time
    const bool tr = p2p_trace_on;
    if (tr && !p2p_ev[0])
        for (auto& e : p2p_ev) EE_CUDA(cudaEventCreate(&e));
    if (tr) EE_CUDA(cudaEventRecord(p2p_ev[0], stream));
    launch_sym(y_in, store);  // k_accel_sym + k_sym_reduce
    if (tr) EE_CUDA(cudaEventRecord(p2p_ev[1], stream));
    k_peer_barrier<<<1, 32, 0, stream>>>(T, ++p2p_epoch, p2p_err_dev.p, p2p_err_d, kPeerTimeoutCycles);
    if (tr) EE_CUDA(cudaEventRecord(p2p_ev[2], stream));
    k_peer_finish<<<fg, 128, 0, stream>>>((int)n, (int)b0, (int)b1, T, ep, p2p_err_dev.p);
    if (tr) EE_CUDA(cudaEventRecord(p2p_ev[3], stream));
    k_peer_barrier<<<1, 32, 0, stream>>>(T, ++p2p_epoch, p2p_err_dev.p, p2p_err_d, kPeerTimeoutCycles);
    if (tr) EE_CUDA(cudaEventRecord(p2p_ev[4], stream));
    EE_CUDA(cudaGetLastError());
    count_launch(3);
    p2p_used = true;
    if (tr) {  // tracing synchronises every step: it is a diagnosis mode, not the timed path
        EE_CUDA(cudaEventSynchronize(p2p_ev[4]));
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            EE_CUDA(cudaEventElapsedTime(&ms, p2p_ev[k], p2p_ev[k + 1]));
            p2p_trace_ms[k] += (double)ms;
        }
        p2p_trace_steps += 1;
    }
}

int32_t NBodyEngine::steady_step() {
    if (!predicted) {
        QtArgs q = qt_args(m, m + 1);
        const int B = 128;
        const int grid = (int)((i1 - i0 + B - 1) / B);
        if (mode == EE_MODE_PARITY)
            k_predict<true><<<grid, B, 0, stream>>>(i0, i1, n, q, ry.p, ra.p);
        else
            k_predict<false><<<grid, B, 0, stream>>>(i0, i1, n, q, ry.p, ra.p);
        EE_CUDA(cudaGetLastError());
        count_launch();
        exchange_y(ry.p + (size_t)slot_of(m + 1) * n);
        predicted = true;
    }
    EpArgs ep{};
    ep.kind = EP_QT;
    ep.qt = qt_args(m + 1, m + 2);
    ep.ry = ry.p;
    ep.ra = ra.p;
    ep.dy = dy.p;
    if (p2p_ready) {
        p2p_step(ep);
    } else {
        accel(ry.p + (size_t)slot_of(m + 1) * n, ep);
        exchange_y(ry.p + (size_t)slot_of(m + 2) * n);
    }
    m += 1;
    t = t + h;  // second_order/mod.rs:122
    return EE_OK;
}

int32_t NBodyEngine::step_once() {
    if (t >= bound) return EE_BOUND_REACHED;           // multistep/mod.rs:203-205
    if (t + h == t) return EE_STEP_SIZE_UNDERFLOW;     // :207-209
    int32_t st = srkn_main ? srkn_step(EE_BM14A_STAGES, EE_BM14A_A, EE_BM14A_B, 1, h) : (m < order ? starter_step() : steady_step());
    if (st) return st;
    if (solout) {
        st = solout->after_step(*this);
        if (st) return st;
    }
    return EE_OK;
}

// Launch the steps that step() has accounted for but not yet run (small-system run-ahead).  At most two batches are in
// flight, so an observer never waits for more than ~2 x kSmallBatch steps of backlog.
void NBodyEngine::flush_pending() {
    EE_CUDA(cudaSetDevice(device));  // observers come through here: make the handle's device current even if nothing is pending
    if (pending == 0) return;
    if (!batch_ev[0]) {
        EE_CUDA(cudaEventCreateWithFlags(&batch_ev[0], cudaEventDisableTiming));
        EE_CUDA(cudaEventCreateWithFlags(&batch_ev[1], cudaEventDisableTiming));
    }
    if (batch_k >= 2) EE_CUDA(cudaEventSynchronize(batch_ev[batch_k & 1]));
    const int64_t k = pending;
    pending = 0;
    if (solout) solout->begin_batch(*this);
    small_steps(*this, m - k, solout ? solout->steps_done - k : 0, k);
    EE_CUDA(cudaEventRecord(batch_ev[batch_k & 1], stream));
    batch_k += 1;
}

// true when every batch launched so far has finished (cudaEventQuery, no synchronisation)
bool NBodyEngine::device_idle() {
    if (batch_k == 0 || !batch_ev[0]) return true;
    return cudaEventQuery(batch_ev[(batch_k - 1) & 1]) == cudaSuccess;
}

// Steady-state steps of a small system (persistent single-CTA kernel) with RUN-AHEAD: the call only evaluates the
// reference's per-step guards and advances the host-side clock and sampling schedule; the steps themselves are launched in
// batches (kSmallBatch, or when the sample buffers are full, or when an observer -- state, take_solution, clone, snapshot,
// sync -- needs the device to be current).  The Prediction Planner calls step() once per step (prediction.rs:429): a launch
// per call would cost several times the 1.2 us of arithmetic a 32-body step takes, so would any CUDA runtime call -- this
// function makes none unless a batch goes out.
int32_t NBodyEngine::step_small(int64_t nsteps) {
    int32_t st = EE_OK;
    int64_t s = 0;
    while (s < nsteps) {
        int64_t k = std::min<int64_t>(nsteps - s, kSmallBatch - pending);
        if (solout) {
            if (solout->room() == 0) solout->flush(*this);  // launches the pending steps first
            k = std::min(k, solout->room());
        }
        double tt = t;
        int64_t ok_steps = 0;
        for (; ok_steps < k; ++ok_steps) {  // the reference's per-step guards (multistep/mod.rs:203-209)
            if (tt >= bound) {
                st = EE_BOUND_REACHED;
                break;
            }
            if (tt + h == tt) {
                st = EE_STEP_SIZE_UNDERFLOW;
                break;
            }
            tt = tt + h;
        }
        if (ok_steps > 0) {
            if (solout) solout->advance_host(ok_steps);
            pending += ok_steps;
            m += ok_steps;
            t = tt;
            predicted = false;
            s += ok_steps;
        }
        if (pending >= kSmallBatch) {
            flush_pending();
        } else if (pending >= 256 && (pending & (pending - 1)) == 0 && device_idle()) {
            // The device has nothing left to do (an observer has just drained it, or the host fell behind): do not make it wait
            // for a full batch.  Asked only when the backlog reaches 256, 512, 1024, 2048 steps -- a query costs a microsecond.
            flush_pending();
        }
        if (st) break;
    }
    timed = false;  // with run-ahead, this call's steps have not (all) been launched yet
    return st;
}

int32_t NBodyEngine::step(int64_t nsteps) {
    const bool small = small_path_available(*this);
    if (small && m >= order) return step_small(nsteps);
    EE_CUDA(cudaSetDevice(device));
    accel_launches = 0;
    EE_CUDA(cudaEventRecord(ev0, stream));
    int32_t st = EE_OK;
    int64_t s = 0;
    while (s < nsteps) {
        if (small && m >= order) {  // start-up finished inside this call: the rest goes through the run-ahead path
            st = step_small(nsteps - s);
            break;
        }
        st = step_once();
        if (st) break;
        ++s;
    }
    EE_CUDA(cudaEventRecord(ev1, stream));
    timed = pending == 0;
    if (p2p_used) check_async_error();
    return st;
}

void NBodyEngine::sync() {
    EE_CUDA(cudaSetDevice(device));
    flush_pending();
    EE_CUDA(cudaStreamSynchronize(stream));
    if (p2p_used) check_async_error();
}

// SoA device state -> AoS `double[3]` rows (what Vec<DVec3> looks like to the caller), so that a state read is ONE
// contiguous device-to-host copy per array and no host-side transposition.
__global__ void k_pack_state(int64_t n, const double4* __restrict__ y, const double* __restrict__ dy, const double* __restrict__ a,
                             double* __restrict__ out_pos, double* __restrict__ out_vel, double* __restrict__ out_acc) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (out_pos) {
        const double4 p = y[k];
        out_pos[3 * k] = p.x;
        out_pos[3 * k + 1] = p.y;
        out_pos[3 * k + 2] = p.z;
    }
    if (out_vel) {
        out_vel[3 * k] = dy[k];
        out_vel[3 * k + 1] = dy[n + k];
        out_vel[3 * k + 2] = dy[2 * n + k];
    }
    if (out_acc) {
        out_acc[3 * k] = a[k];
        out_acc[3 * k + 1] = a[n + k];
        out_acc[3 * k + 2] = a[2 * n + k];
    }
}

// Enqueue a read of the current state into caller memory and return without waiting: the pack kernel runs on the
// compute stream (so it sees exactly the state after the steps queued so far, and later steps cannot overwrite what it
// reads), the device-to-host copies run on a separate copy stream and overlap the following steps.  Two device staging
// buffers alternate; state_wait() blocks until every enqueued read has landed.  Page-locked caller memory gives true
// overlap; pageable memory works but the driver stages it.
void NBodyEngine::state_async(double* time, double* pos, double* vel, double* acc) {
    EE_CUDA(cudaSetDevice(device));
    flush_pending();
    if (acc) ensure_a0();
    if (time) *time = t;
    if (p2p_used) check_async_error();
    EE_REQUIRE(!(world > 1 && exchange == EE_EXCHANGE_ALLGATHER), "internal: state_async on a target-sharded handle");
    if (!pos && !vel && !acc) return;
    if (!copy_stream) {
        EE_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            stage_d[k].alloc((size_t)9 * n);
            EE_CUDA(cudaEventCreateWithFlags(&stage_packed[k], cudaEventDisableTiming));
            EE_CUDA(cudaEventCreateWithFlags(&stage_copied[k], cudaEventDisableTiming));
        }
    }
    const int k = (int)(stage_k & 1);
    if (stage_k >= 2) EE_CUDA(cudaStreamWaitEvent(stream, stage_copied[k], 0));  // staging buffer k is free again
    double* dpos = stage_d[k].p;
    double* dvel = dpos + 3 * n;
    double* dacc = dpos + 6 * n;
    const int B = 256;
    k_pack_state<<<(unsigned)((n + B - 1) / B), B, 0, stream>>>(n, ry.p + (size_t)slot_of(m) * n, dy.p,
                                                               ra.p + (size_t)slot_of(m) * 3 * n, pos ? dpos : nullptr,
                                                               vel ? dvel : nullptr, acc ? dacc : nullptr);
    EE_CUDA(cudaGetLastError());
    count_launch();
    EE_CUDA(cudaEventRecord(stage_packed[k], stream));
    EE_CUDA(cudaStreamWaitEvent(copy_stream, stage_packed[k], 0));
    const size_t bytes = (size_t)3 * n * sizeof(double);
    if (pos) EE_CUDA(cudaMemcpyAsync(pos, dpos, bytes, cudaMemcpyDeviceToHost, copy_stream));
    if (vel) EE_CUDA(cudaMemcpyAsync(vel, dvel, bytes, cudaMemcpyDeviceToHost, copy_stream));
    if (acc) EE_CUDA(cudaMemcpyAsync(acc, dacc, bytes, cudaMemcpyDeviceToHost, copy_stream));
    EE_CUDA(cudaEventRecord(stage_copied[k], copy_stream));
    stage_k += 1;
}

void NBodyEngine::state_wait() {
    EE_CUDA(cudaSetDevice(device));
    if (copy_stream) EE_CUDA(cudaStreamSynchronize(copy_stream));
    if (p2p_used) check_async_error();
}

void NBodyEngine::state(double* time, double* pos, double* vel, double* acc) {
    EE_CUDA(cudaSetDevice(device));
    const bool local_only = world > 1 && exchange == EE_EXCHANGE_ALLGATHER;  // only the own slice of dy / ra is current
    if (!local_only) {
        state_async(time, pos, vel, acc);
        state_wait();
        return;
    }
    flush_pending();
    if (acc) ensure_a0();
    if (time) *time = t;
    const int64_t per = n / world, g0 = rank * per;
    std::vector<double> hv;
    if (pos) {  // positions are all-gathered every step: every rank holds all of them
        DBuf<double> dpos((size_t)3 * n);
        const int B = 256;
        k_pack_state<<<(unsigned)((n + B - 1) / B), B, 0, stream>>>(n, ry.p + (size_t)slot_of(m) * n, nullptr, nullptr, dpos.p,
                                                                   nullptr, nullptr);
        EE_CUDA(cudaGetLastError());
        count_launch();
        EE_CUDA(cudaMemcpyAsync(pos, dpos.p, dpos.bytes(), cudaMemcpyDeviceToHost, stream));
        EE_CUDA(cudaStreamSynchronize(stream));
    }
    auto fetch_soa = [&](const double* src, double* dst) {  // SoA [3][n], own slice valid: gather across ranks first
        DBuf<double> tmp((size_t)3 * per), all((size_t)3 * n);
        for (int c = 0; c < 3; ++c)
            EE_CUDA(cudaMemcpyAsync(tmp.p + (size_t)c * per, src + (size_t)c * n + g0, (size_t)per * sizeof(double),
                                    cudaMemcpyDeviceToDevice, stream));
        EE_NCCL(nccl().AllGather(tmp.p, all.p, (size_t)3 * per, ncclDouble, (ncclComm_t)comm, stream));
        hv.resize((size_t)3 * n);
        EE_CUDA(cudaMemcpyAsync(hv.data(), all.p, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        EE_CUDA(cudaStreamSynchronize(stream));
        for (int r = 0; r < world; ++r)
            for (int c = 0; c < 3; ++c)
                for (int64_t k = 0; k < per; ++k) dst[3 * (r * per + k) + c] = hv[(size_t)((r * 3 + c) * per + k)];
    };
    if (vel) fetch_soa(dy.p, vel);
    if (acc) fetch_soa(ra.p + (size_t)slot_of(m) * 3 * n, acc);
    EE_CUDA(cudaStreamSynchronize(stream));
}

double NBodyEngine::last_step_ms() {
    if (!timed) return 0.0;
    EE_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    EE_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    return (double)ms;
}

// Clone (prediction.rs:224-229 clones the propagator at every snapshot).  The copy is device-to-device on the clone's
// stream and never touches the host; scratch buffers of the throughput kernels are only allocated when a handle first
// evaluates an acceleration, so a snapshot clone that is never stepped costs the state arrays and nothing else.
// A sharded handle (pair units sharded: every rank holds the complete state) clones into an UNSHARDED replica on the
// calling rank's GPU -- it continues the same trajectory on one GPU.
NBodyEngine* NBodyEngine::clone() {
    EE_REQUIRE(world == 1 || exchange == EE_EXCHANGE_ALLREDUCE,
               "clone of a target-sharded (allgather) propagator is not supported: no rank holds the complete state");
    sync();  // launches pending run-ahead steps too
    NBodyEngine* c = new NBodyEngine(n, nullptr, nullptr, nullptr, t, h, method, mode, device, 0, 1, nullptr, 0);
    try {
        c->m = m;
        c->t = t;
        c->have_a0 = have_a0;
        c->predicted = predicted;
        EE_CUDA(cudaMemcpyAsync(c->ry.p, ry.p, ry.bytes(), cudaMemcpyDeviceToDevice, c->stream));
        EE_CUDA(cudaMemcpyAsync(c->ra.p, ra.p, ra.bytes(), cudaMemcpyDeviceToDevice, c->stream));
        EE_CUDA(cudaMemcpyAsync(c->dy.p, dy.p, dy.bytes(), cudaMemcpyDeviceToDevice, c->stream));
        if (solout) c->solout.reset(solout->clone(*c));
        EE_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

struct SnapHeader {
    uint64_t magic;
    int64_t n, m;
    int32_t method, mode, order, R;
    int32_t have_a0, predicted;
    double t, h;
    int64_t solout_bytes;  // 0 = no solout attached
};
static const uint64_t kSnapMagic = 0x45455f534e415032ull;  // "EE_SNAP2"

int64_t NBodyEngine::snapshot_bytes() const {
    return (int64_t)sizeof(SnapHeader) + (int64_t)(ry.bytes() + ra.bytes() + dy.bytes()) + (solout ? solout->blob_bytes() : 0);
}

void NBodyEngine::snapshot(void* blob) {
    EE_REQUIRE(world == 1 || exchange == EE_EXCHANGE_ALLREDUCE,
               "snapshot of a target-sharded (allgather) propagator is not supported: no rank holds the complete state");
    EE_CUDA(cudaSetDevice(device));
    flush_pending();
    SnapHeader hd{kSnapMagic, n, m, method, mode, order, R, have_a0 ? 1 : 0, predicted ? 1 : 0, t, h,
                  solout ? solout->blob_bytes() : 0};
    unsigned char* p = (unsigned char*)blob;
    std::memcpy(p, &hd, sizeof(hd));
    p += sizeof(hd);
    EE_CUDA(cudaMemcpyAsync(p, ry.p, ry.bytes(), cudaMemcpyDeviceToHost, stream));
    p += ry.bytes();
    EE_CUDA(cudaMemcpyAsync(p, ra.p, ra.bytes(), cudaMemcpyDeviceToHost, stream));
    p += ra.bytes();
    EE_CUDA(cudaMemcpyAsync(p, dy.p, dy.bytes(), cudaMemcpyDeviceToHost, stream));
    p += dy.bytes();
    EE_CUDA(cudaStreamSynchronize(stream));
    if (solout) solout->save(*this, p);
    if (p2p_used) check_async_error();
}

// Every rank of a sharded handle restores the same blob (the state is replicated); the caller keeps the ranks in step.
void NBodyEngine::restore(const void* blob, int64_t blob_bytes) {
    EE_REQUIRE(world == 1 || exchange == EE_EXCHANGE_ALLREDUCE,
               "restore of a target-sharded (allgather) propagator is not supported");
    EE_CUDA(cudaSetDevice(device));
    flush_pending();
    SnapHeader hd;
    EE_REQUIRE(blob_bytes >= (int64_t)sizeof(hd), "snapshot blob is truncated");
    std::memcpy(&hd, blob, sizeof(hd));
    EE_REQUIRE(hd.magic == kSnapMagic, "not a snapshot blob");
    EE_REQUIRE(hd.n == n && hd.method == method && hd.mode == mode && hd.R == R && hd.order == order,
               "snapshot does not match this handle");
    EE_REQUIRE(hd.m >= 0 && std::isfinite(hd.t) && std::isfinite(hd.h) && hd.h != 0.0 && hd.solout_bytes >= 0,
               "corrupt snapshot header");
    EE_REQUIRE(blob_bytes >= (int64_t)sizeof(hd) + (int64_t)(ry.bytes() + ra.bytes() + dy.bytes()) + hd.solout_bytes,
               "snapshot blob is truncated");
    const unsigned char* p = (const unsigned char*)blob + sizeof(hd);
    EE_CUDA(cudaMemcpyAsync(ry.p, p, ry.bytes(), cudaMemcpyHostToDevice, stream));
    p += ry.bytes();
    EE_CUDA(cudaMemcpyAsync(ra.p, p, ra.bytes(), cudaMemcpyHostToDevice, stream));
    p += ra.bytes();
    EE_CUDA(cudaMemcpyAsync(dy.p, p, dy.bytes(), cudaMemcpyHostToDevice, stream));
    p += dy.bytes();
    EE_CUDA(cudaStreamSynchronize(stream));  // the caller may reuse (or unpin) the blob as soon as this returns
    m = hd.m;
    t = hd.t;
    h = hd.h;
    hs = h * (1.0 / 4.0);
    have_a0 = hd.have_a0 != 0;
    predicted = hd.predicted != 0;
    if (hd.solout_bytes > 0)
        solout.reset(Solout::load(*this, p, hd.solout_bytes));
    else
        solout.reset();
}

// K steps, each bracketed by events, with an L2-evicting memset before each (outside the timed interval)
double NBodyEngine::step_timed(int64_t nsteps, int64_t flush_bytes, int32_t* status) {
    EE_CUDA(cudaSetDevice(device));
    flush_pending();
    if (flush_bytes > 0 && (int64_t)flush_buf.n < flush_bytes) flush_buf.alloc((size_t)flush_bytes);
    double total = 0.0;
    accel_launches = 0;
    *status = EE_OK;
    for (int64_t s = 0; s < nsteps; ++s) {
        if (flush_bytes > 0) EE_CUDA(cudaMemsetAsync(flush_buf.p, (int)(s & 0xff), (size_t)flush_bytes, stream));
        EE_CUDA(cudaEventRecord(ev0, stream));
        const int32_t st = step_once();
        EE_CUDA(cudaEventRecord(ev1, stream));
        EE_CUDA(cudaEventSynchronize(ev1));
        if (st) {
            *status = st;
            break;
        }
        float ms = 0.f;
        EE_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        total += (double)ms;
    }
    timed = false;
    if (p2p_used) check_async_error();
    return total;
}

// Sustained DFMA rate: every thread runs 8 independent FMA chains; 2 flop per FMA.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c);
            a1 = fma(a1, b, c);
            a2 = fma(a2, b, c);
            a3 = fma(a3, b, c);
            a4 = fma(a4, b, c);
            a5 = fma(a5, b, c);
            a6 = fma(a6, b, c);
            a7 = fma(a7, b, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

double fp64_fma_peak(int device) {
    int ndev = 0;
    EE_CUDA(cudaGetDeviceCount(&ndev));
    EE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU path)");
    EE_CUDA(cudaSetDevice(device));
    const int blocks = device_info(device).sm_count * 8, threads = 256, iters = 4096;
    DBuf<double> out((size_t)blocks * threads);
    cudaEvent_t e0, e1;
    EE_CUDA(cudaEventCreate(&e0));
    EE_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        EE_CUDA(cudaEventRecord(e0));
        for (int k = 0; k < 4; ++k) k_dfma_peak<<<blocks, threads>>>(out.p, iters, 1.0 + rep);
        EE_CUDA(cudaEventRecord(e1));
        EE_CUDA(cudaEventSynchronize(e1));
        EE_CUDA(cudaGetLastError());
        float ms = 0.f;
        EE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 4.0 * (double)blocks * threads * (double)iters * 64.0 * 2.0;
        if (rep > 0) best = std::max(best, flops / ((double)ms * 1e-3) / 1e12);
    }
    count_launch(20);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

// Work list of one rank.  Units are (tile row, 32-body chunk) in canonical order; the rank owns an equal contiguous share
// (to within one unit).  Items are runs of units inside one row with GUIDED sizes: remaining / spread rounded down to a
// power of two (spread = the number of items the remaining work is dealt into; the engine passes the number of persistent
// CTAs), at most max_chunks, at least one unit -- long items while there is plenty of work, single chunks at the end, so
// the dynamic queue's tail is one chunk.  Queue order = canonical order = slot order.
SymSchedule build_sym_schedule(int64_t n, int tile, int spread, int world, int rank, int max_chunks) {
    const int ctas = spread;
    EE_REQUIRE(n > 0 && tile >= 64 && tile % 32 == 0 && n % tile == 0, "n must be a positive multiple of the tile");
    EE_REQUIRE(world >= 1 && rank >= 0 && rank < world && ctas >= 1 && max_chunks >= 1, "bad schedule arguments");
    SymSchedule sc;
    const long long nch = n / 32, cpt = tile / 32, nt = n / tile;
    sc.u_total = sym_row_unit(nt, nch, cpt);
    sc.u_lo = sc.u_total * rank / world;
    sc.u_hi = sc.u_total * (rank + 1) / world;
    std::vector<int> row_count((size_t)nt, 0);
    long long u = sc.u_lo, ti = 0;
    int slot = 0;
    while (u < sc.u_hi) {
        const long long row_begin = sym_row_unit(ti, nch, cpt), row_end = sym_row_unit(ti + 1, nch, cpt);
        if (u >= row_end) {
            ++ti;
            continue;
        }
        const long long remaining = sc.u_hi - u;
        long long want = remaining / (long long)ctas, size = 1;
        while (size * 2 <= want && size * 2 <= max_chunks) size *= 2;
        size = std::min(size, std::min(row_end - u, remaining));
        sc.items.push_back(SymItem{(int)ti, (int)(ti * cpt + (u - row_begin)), (int)size, slot++});
        row_count[(size_t)ti] += 1;
        u += size;
    }
    sc.row_slot.assign((size_t)nt + 1, 0);
    for (long long r = 0; r < nt; ++r) sc.row_slot[(size_t)r + 1] = sc.row_slot[(size_t)r] + row_count[(size_t)r];
    return sc;
}

namespace {
// Launch with the programmatic-stream-serialization attribute (see pdl_wait/pdl_trigger in ee_sym.cuh).
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    static const bool off = getenv("EE_DEV_AIDS") && getenv("EE_DEV_AIDS")[0] == '1' && getenv("EE_PDL") && getenv("EE_PDL")[0] == '0';
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = off ? 0 : 1;
    EE_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

template <int TI, int NT, int MINB, int SBC>
void launch_sym_variant(NBodyEngine& e, const double4* y_in, const EpArgs& ep) {
    using Smem = SymSmem<NT / 32, SBC>;
    static std::atomic<uint64_t> attr_mask{0};  // opt-in shared memory size: a per-device function attribute
    const uint64_t bit = 1ull << (e.device & 63);
    if (!(attr_mask.load(std::memory_order_acquire) & bit)) {
        EE_CUDA(cudaFuncSetAttribute(k_accel_sym<TI, NT, MINB, SBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        attr_mask.fetch_or(bit, std::memory_order_release);
    }
    static const bool prof = getenv("EE_DEV_AIDS") && getenv("EE_DEV_AIDS")[0] == '1' && getenv("EE_SYM_PROF");
    if (prof) {  // developer aid: per-phase cycle counts of this launch to stderr (synchronises)
        static std::atomic<uint64_t> pmask{0};
        if (!(pmask.load() & bit)) {
            EE_CUDA(cudaFuncSetAttribute(k_accel_sym<TI, NT, MINB, SBC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
            pmask.fetch_or(bit);
        }
        const int G = MINB * e.sm_count;
        DBuf<long long> d((size_t)G * 9);
        k_accel_sym<TI, NT, MINB, SBC, true><<<G, NT, sizeof(Smem), e.stream>>>(
            (int)e.n, y_in, e.sym_items.p, e.sym_n_items, e.sym_counter.p, e.sym_part_i.p, e.sym_part_j.p, d.p);
        std::vector<long long> h((size_t)G * 9);
        EE_CUDA(cudaMemcpyAsync(h.data(), d.p, h.size() * 8, cudaMemcpyDeviceToHost, e.stream));
        EE_CUDA(cudaStreamSynchronize(e.stream));
        double tot[5] = {0, 0, 0, 0, 0}, cmin = 1e300, cmax = 0, imin = 1e300, imax = 0;
        long long t_first = h[5], t_last = 0;
        for (int g = 0; g < G; ++g) {
            for (int q = 0; q < 5; ++q) tot[q] += (double)h[(size_t)g * 9 + q];
            const double c = (double)(h[(size_t)g * 9] + h[(size_t)g * 9 + 1] + h[(size_t)g * 9 + 2] + h[(size_t)g * 9 + 3]);
            cmin = std::min(cmin, c);
            cmax = std::max(cmax, c);
            imin = std::min(imin, (double)h[(size_t)g * 9 + 4]);
            imax = std::max(imax, (double)h[(size_t)g * 9 + 4]);
            t_first = std::min(t_first, h[(size_t)g * 9 + 5]);
            t_last = std::max(t_last, h[(size_t)g * 9 + 6]);
        }
        if (getenv("EE_SYM_PROF_DUMP")) {  // per CTA: start, end, start of the last item (us from the first start) and its size
            for (int g = 0; g < G; ++g)
                fprintf(stderr, "[sym-prof-cta] %d start %.1f end %.1f last_item_start %.1f last_item_chunks %lld items %lld\n", g,
                        (h[(size_t)g * 9 + 5] - t_first) * 1e-3, (h[(size_t)g * 9 + 6] - t_first) * 1e-3,
                        (h[(size_t)g * 9 + 7] - t_first) * 1e-3, h[(size_t)g * 9 + 8], h[(size_t)g * 9 + 4]);
        }
        fprintf(stderr, "[sym-prof] makespan %.1f us (first CTA start to last CTA end)\n", (t_last - t_first) * 1e-3);
        fprintf(stderr, "[sym-prof] busy cycles per CTA: min %.0f max %.0f; items per CTA: min %.0f max %.0f\n", cmin, cmax, imin, imax);
        const double all = tot[0] + tot[1] + tot[2] + tot[3];
        fprintf(stderr, "[sym-prof] items %d over %d CTAs (%.1f/CTA): cycles per CTA %.0f = prologue %.1f%% + chunks %.1f%% + merge %.1f%% + "
                        "i-store %.1f%%; per item: prologue %.0f merge %.0f i-store %.0f cycles\n",
                e.sym_n_items, G, tot[4] / G, all / G, 100 * tot[0] / all, 100 * tot[1] / all, 100 * tot[2] / all, 100 * tot[3] / all,
                tot[0] / tot[4], tot[2] / tot[4], tot[3] / tot[4]);
    } else {
        launch_pdl(k_accel_sym<TI, NT, MINB, SBC>, dim3(MINB * e.sm_count), dim3(NT), sizeof(Smem), e.stream, (int)e.n, y_in,
                   (const SymItem*)e.sym_items.p, e.sym_n_items, e.sym_counter.p, e.sym_part_i.p, e.sym_part_j.p, (long long*)nullptr);
    }
    constexpr int KL = NT >= 128 ? 8 : 32, KB = NT >= 128 ? 32 : 8;  // threads per body, bodies per CTA (see k_sym_reduce)
    const unsigned rg = (unsigned)((e.n + KB - 1) / KB);
    launch_pdl(k_sym_reduce<TI * NT, KL, KB>, dim3(rg), dim3(KL * KB), 0, e.stream, (int)e.n, e.sym_share, (const int*)e.sym_row_slot.p,
               (const double*)e.sym_part_i.p, (const double*)e.sym_part_j.p, e.sym_counter.p, (unsigned)(MINB * e.sm_count), ep);
}
}  // namespace

// Pair-symmetric acceleration of the positions in y_in (this rank's share of the units) + epilogue `ep`: two launches.
void NBodyEngine::launch_sym(const double4* y_in, const EpArgs& ep) {
    const int key = ((sym_ti * 1000 + sym_nt) * 10 + sym_minb) * 100 + sym_sbc;
    switch (key) {
        case 4256216: launch_sym_variant<4, 256, 2, 16>(*this, y_in, ep); break;
        case 4128316: launch_sym_variant<4, 128, 3, 16>(*this, y_in, ep); break;
        case 4033604: launch_sym_variant<4, 32, 16, 4>(*this, y_in, ep); break;
        case 4128404: launch_sym_variant<4, 128, 4, 4>(*this, y_in, ep); break;
        // other shapes were measured and dropped (profiles/r02/mid_probe_*.jsonl, profiles/README.md): TI = 2 or 8, one CTA
        // of 8 warps per SM, 2-warp CTAs
        default: throw Error(EE_ERR_INVALID, "unknown pair-symmetric kernel variant (TI,NT,MINB,SBC)");
    }
    EE_CUDA(cudaGetLastError());
    count_launch(2);
    accel_launches++;
}

// stand-alone NewtonianGravity::eval
void gravity_eval(int64_t n, const double* pos, const double* mus, int mode, int device, double* acc) {
    std::vector<double> vel((size_t)3 * n, 0.0);
    NBodyEngine e(n, pos, vel.data(), mus, 0.0, 1.0, EE_QUINLAN_TREMAINE_12, mode, device, 0, 1, nullptr, 0);
    e.state(nullptr, nullptr, nullptr, acc);
}

}  // namespace ee
