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

#include "ee_coeffs.h"
#include "ee_engine.h"
#include "ee_kernels.cuh"
#include "ee_sym.cuh"

namespace ee {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};

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
    throw Error(EE_ERR_INVALID, "unknown method id (12 = QuinlanTremaine12, 13 = Stormer13)");
}

NBodyEngine::NBodyEngine(int64_t n_, const double* pos, const double* vel, const double* mus, double t0, double h_signed,
                         int method_, int mode_, int device_, int rank_, int world_, const void* uid, int exchange_)
    : n(n_), method(method_), mode(mode_), device(device_), rank(rank_), world(world_), exchange(exchange_) {
    EE_REQUIRE(n >= 1, "n must be >= 1");
    EE_REQUIRE(pos && vel && mus, "null input array");
    EE_REQUIRE(mode == EE_MODE_PARITY || mode == EE_MODE_THROUGHPUT, "unknown mode");
    EE_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    MethodTable mt = method_table(method);
    order = mt.order;
    R = order + 1;
    h = h_signed;
    hs = h_signed * (1.0 / 4.0);  // Substepper::new -- multistep/mod.rs:53-58
    t = t0;
    bound = std::numeric_limits<double>::infinity();  // nbody.rs:112
    int ndev = 0;
    EE_CUDA(cudaGetDeviceCount(&ndev));
    EE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this engine has no CPU path)");
    EE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    EE_CUDA(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
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
    ry.alloc((size_t)R * n);
    ra.alloc((size_t)R * 3 * n);
    dy.alloc((size_t)3 * n);
    ytmp[0].alloc((size_t)n);
    ytmp[1].alloc((size_t)n);
    a_scr.alloc((size_t)3 * n);
    EE_CUDA(cudaMemsetAsync(ry.p, 0, ry.bytes(), stream));
    EE_CUDA(cudaMemsetAsync(ra.p, 0, ra.bytes(), stream));
    std::vector<double4> hp((size_t)n);
    std::vector<double> hv((size_t)3 * n);
    for (int64_t k = 0; k < n; ++k) {
        hp[(size_t)k] = make_double4(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], mus[k]);
        for (int c = 0; c < 3; ++c) hv[(size_t)(c * n + k)] = vel[3 * k + c];
    }
    EE_CUDA(cudaMemcpyAsync(ry.p, hp.data(), (size_t)n * sizeof(double4), cudaMemcpyHostToDevice, stream));
    EE_CUDA(cudaMemcpyAsync(dy.p, hv.data(), hv.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    EE_CUDA(cudaStreamSynchronize(stream));
    plan_launch();
}

NBodyEngine::~NBodyEngine() {
    for (void* p : p2p_opened) cudaIpcCloseMemHandle(p);
    delete (PeerTable*)p2p_table;
    if (comm) nccl().CommDestroy((ncclComm_t)comm);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
}

// Choose the throughput-kernel decomposition: tiles x splits blocks, with the block count a whole number of waves
// (sm_count x resident CTAs per SM) whenever the problem is big enough to allow it.
void NBodyEngine::plan_launch() {
    // Large systems take the pair-symmetric kernel (ee_sym.cuh): single GPU, or sources sharded + allreduce.
    const char* env = getenv("EE_SYM");
    const bool sym_allowed = !(env && env[0] == '0');
    use_sym = sym_allowed && mode == EE_MODE_THROUGHPUT && n % kSymTile == 0 && n >= 32768 &&
              (world == 1 || exchange == EE_EXCHANGE_ALLREDUCE);
    if (use_sym) {
        // superchunk size: 512 on one GPU (4160 items for 296 resident CTAs); 256 when sharded, so that every rank still
        // has several items per CTA (measured at 2 and 8 GPUs: 256 beats both 512 and 128, profiles/r01/README.md)
        const long long nt = n / kSymTile;
        const char* jsenv = getenv("EE_SYM_JS");
        sym_js = world == 1 ? 512 : 256;
        if (jsenv) sym_js = atoi(jsenv);
        EE_REQUIRE(sym_js == 512 || sym_js == 256 || sym_js == 128, "EE_SYM_JS must be 512, 256 or 128");
        const long long ns = n / sym_js;
        const long long total = sym_item_prefix(nt, ns, kSymTile / sym_js);
        sym_lo = total * rank / world;
        sym_hi = total * (rank + 1) / world;
        if (const char* sh = getenv("EE_SYM_SHARE")) {  // developer aid: time one rank's share of a G-way split on one GPU
            const int g = atoi(sh);
            if (g > 1 && world == 1) sym_hi = total / g;
        }
        if (const char* rg = getenv("EE_SYM_RANGE")) {  // developer aid "a/b": behave like rank a of b on one GPU (partial sums)
            int a = 0, b = 1;
            if (sscanf(rg, "%d/%d", &a, &b) == 2 && b >= 1 && a >= 0 && a < b && world == 1) {
                sym_lo = total * a / b;
                sym_hi = total * (a + 1) / b;
            }
        }
        const char* senv = getenv("EE_SYM_STATIC");
        sym_static = senv ? senv[0] == '1' : false;
        if (sym_static) {
            // static chunk-granular split: CTA g owns units [u_len*g/G, u_len*(g+1)/G) of this rank's list; mark the items a
            // boundary falls strictly inside of (their i-side sum arrives in two slots)
            const long long chunks = sym_js / 32, G = 2LL * sm_count, items = sym_hi - sym_lo, u_len = items * chunks;
            EE_REQUIRE(u_len / G >= chunks, "static split needs at least one item per CTA");
            std::vector<unsigned char> sp((size_t)std::max<long long>(1, items), 0);
            for (long long g = 1; g < G; ++g) {
                const long long ub = u_len * g / G;
                if (ub % chunks != 0 && ub < u_len) sp[(size_t)(ub / chunks)] = 1;
            }
            sym_split.alloc(sp.size());
            EE_CUDA(cudaMemcpy(sym_split.p, sp.data(), sp.size(), cudaMemcpyHostToDevice));
            EE_CUDA(cudaFuncSetAttribute(k_accel_sym_static<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymSmem<512>)));
            EE_CUDA(cudaFuncSetAttribute(k_accel_sym_static<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymSmem<256>)));
            EE_CUDA(cudaFuncSetAttribute(k_accel_sym_static<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymSmem<128>)));
        }
        sym_part_i.alloc((size_t)ns * (sym_static ? 2 : 1) * 3 * n);
        sym_part_j.alloc((size_t)nt * 3 * n);
        sym_counter.alloc(1);
        EE_CUDA(cudaFuncSetAttribute(k_accel_sym<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymSmem<512>)));
        EE_CUDA(cudaFuncSetAttribute(k_accel_sym<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymSmem<256>)));
        EE_CUDA(cudaFuncSetAttribute(k_accel_sym<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymSmem<128>)));
    }
    block = n >= 16384 ? 256 : 128;
    const int64_t targets = i1 - i0, sources = j1 - j0;
    tiles = (int)((targets + block - 1) / block);
    int occ = 0;
    if (block == 256)
        EE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_accel_fast<256>, 256, 0));
    else
        EE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_accel_fast<128>, 128, 0));
    occ = std::max(1, occ);
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
    if (splits > 1) {
        part.alloc((size_t)splits * 3 * n);
        tickets.alloc((size_t)tiles);
        EE_CUDA(cudaMemsetAsync(tickets.p, 0, tickets.bytes(), stream));
    }
}

void NBodyEngine::exchange_y(double4* buf) {
    if (world == 1 || exchange != EE_EXCHANGE_ALLGATHER) return;
    const int64_t per = n / world;
    EE_NCCL(nccl().AllGather((const void*)(buf + i0), (void*)buf, (size_t)per * 4, ncclDouble, (ncclComm_t)comm, stream));
}

// One acceleration evaluation of the positions in `y_in`, followed by epilogue `ep` for every local target.
void NBodyEngine::accel(const double4* y_in, EpArgs ep) {
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
        EE_CUDA(cudaMemsetAsync(sym_counter.p, 0, sizeof(unsigned long long), stream));
        const unsigned rg = (unsigned)((n + 255) / 256);
        if (sym_static) {
#define EE_SYM_STATIC_LAUNCH(JS)                                                                                                  \
    k_accel_sym_static<JS><<<2 * sm_count, kSymThreads, sizeof(SymSmem<JS>), stream>>>(n, y_in, sym_lo, sym_hi, sym_part_i.p,     \
                                                                                       sym_part_j.p);                            \
    k_sym_reduce_static<JS><<<rg, 256, 0, stream>>>(n, sym_lo, sym_hi, sym_part_i.p, sym_part_j.p, sym_split.p, kep);
            if (sym_js == 512) {
                EE_SYM_STATIC_LAUNCH(512)
            } else if (sym_js == 256) {
                EE_SYM_STATIC_LAUNCH(256)
            } else {
                EE_SYM_STATIC_LAUNCH(128)
            }
#undef EE_SYM_STATIC_LAUNCH
        } else if (sym_js == 512) {
            k_accel_sym<512><<<2 * sm_count, kSymThreads, sizeof(SymSmem<512>), stream>>>(n, y_in, sym_lo, sym_hi, sym_counter.p,
                                                                                         sym_part_i.p, sym_part_j.p);
            k_sym_reduce<512><<<rg, 256, 0, stream>>>(n, sym_lo, sym_hi, sym_part_i.p, sym_part_j.p, kep);
        } else if (sym_js == 256) {
            k_accel_sym<256><<<2 * sm_count, kSymThreads, sizeof(SymSmem<256>), stream>>>(n, y_in, sym_lo, sym_hi, sym_counter.p,
                                                                                         sym_part_i.p, sym_part_j.p);
            k_sym_reduce<256><<<rg, 256, 0, stream>>>(n, sym_lo, sym_hi, sym_part_i.p, sym_part_j.p, kep);
        } else {
            k_accel_sym<128><<<2 * sm_count, kSymThreads, sizeof(SymSmem<128>), stream>>>(n, y_in, sym_lo, sym_hi, sym_counter.p,
                                                                                         sym_part_i.p, sym_part_j.p);
            k_sym_reduce<128><<<rg, 256, 0, stream>>>(n, sym_lo, sym_hi, sym_part_i.p, sym_part_j.p, kep);
        }
        EE_CUDA(cudaGetLastError());
        count_launch();
    } else if (mode == EE_MODE_PARITY) {
        constexpr int B = 128;
        const int grid = (int)((i1 - i0 + B - 1) / B);
        k_accel_parity<B><<<grid, B, 0, stream>>>(n, i0, i1, y_in, kep);
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

// One call of LinearMultistepIntegrator::advance while the starter is active: 4 x BlanesMoan6B(h/4), then the
// acceleration at the new state (which is the starter's own last-stage evaluation: A[6] = 0 leaves y untouched).
int32_t NBodyEngine::starter_step() {
    ensure_a0();
    const double4* y_in = ry.p + (size_t)slot_of(m) * n;
    const double* a_in = ra.p + (size_t)slot_of(m) * 3 * n;
    int pp = 0;
    for (int sub = 0; sub < 4; ++sub) {
        if (t >= bound) return EE_BOUND_REACHED;             // runge_kutta/mod.rs:113-115
        if (t + hs == t) return EE_STEP_SIZE_UNDERFLOW;      // :117-119
        for (int s = 0; s < EE_BM6B_STAGES; ++s) {
            const bool final_stage = sub == 3 && s == EE_BM6B_STAGES - 1;
            double4* y_out = final_stage ? ry.p + (size_t)slot_of(m + 1) * n : ytmp[pp].p;
            EpArgs ep{};
            ep.kind = EP_KD;
            ep.kd.hb = hs * EE_BM6B_B[s];
            ep.kd.ha = hs * EE_BM6B_A[s];
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
        t = t + hs;  // symplectic.rs:98
    }
    m += 1;
    predicted = false;
    return EE_OK;
}

struct P2PBlob {
    cudaIpcMemHandle_t a_part, ry, flags, ra, dy, unused[3];
};
static_assert(sizeof(P2PBlob) == 512, "8 IPC handle slots of 64 bytes");

void NBodyEngine::p2p_export(void* blob256 /* 512 bytes */) {
    EE_REQUIRE(world > 1 && use_sym && exchange == EE_EXCHANGE_ALLREDUCE,
               "the peer path needs a sharded handle on the pair-symmetric kernel (throughput mode, allreduce layout, n >= 32768)");
    EE_REQUIRE(world <= kMaxPeers, "at most 8 peers");
    EE_CUDA(cudaSetDevice(device));
    if (!p2p_flags.p) {
        p2p_flags.alloc(kMaxPeers);
        p2p_err.alloc(1);
        EE_CUDA(cudaMemset(p2p_flags.p, 0, p2p_flags.bytes()));
        EE_CUDA(cudaMemset(p2p_err.p, 0, sizeof(int)));
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

void NBodyEngine::check_async_error() {
    if (!p2p_err.p) return;
    int e = 0;
    EE_CUDA(cudaMemcpy(&e, p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) throw Error(EE_ERR_CUDA, "peer barrier timed out (a rank did not arrive)");
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
    EE_CUDA(cudaMemsetAsync(sym_counter.p, 0, sizeof(unsigned long long), stream));
    const int64_t per = n / world, b0 = rank * per, b1 = b0 + per;
    const unsigned rg = (unsigned)((n + 255) / 256), fg = (unsigned)((per + 127) / 128);
#define EE_P2P_LAUNCH(JS)                                                                                                        \
    k_accel_sym<JS><<<2 * sm_count, kSymThreads, sizeof(SymSmem<JS>), stream>>>(n, y_in, sym_lo, sym_hi, sym_counter.p,          \
                                                                                sym_part_i.p, sym_part_j.p);                    \
    k_sym_reduce<JS><<<rg, 256, 0, stream>>>(n, sym_lo, sym_hi, sym_part_i.p, sym_part_j.p, store);
#define EE_P2P_LAUNCH_STATIC(JS)                                                                                                 \
    k_accel_sym_static<JS><<<2 * sm_count, kSymThreads, sizeof(SymSmem<JS>), stream>>>(n, y_in, sym_lo, sym_hi, sym_part_i.p,     \
                                                                                       sym_part_j.p);                            \
    k_sym_reduce_static<JS><<<rg, 256, 0, stream>>>(n, sym_lo, sym_hi, sym_part_i.p, sym_part_j.p, sym_split.p, store);
    if (sym_static) {
        if (sym_js == 512) {
            EE_P2P_LAUNCH_STATIC(512)
        } else if (sym_js == 256) {
            EE_P2P_LAUNCH_STATIC(256)
        } else {
            EE_P2P_LAUNCH_STATIC(128)
        }
    } else if (sym_js == 512) {
        EE_P2P_LAUNCH(512)
    } else if (sym_js == 256) {
        EE_P2P_LAUNCH(256)
    } else {
        EE_P2P_LAUNCH(128)
    }
#undef EE_P2P_LAUNCH
#undef EE_P2P_LAUNCH_STATIC
    k_peer_barrier<<<1, 32, 0, stream>>>(T, ++p2p_epoch, p2p_err.p);
    k_peer_finish<<<fg, 128, 0, stream>>>(n, b0, b1, T, ep);
    k_peer_barrier<<<1, 32, 0, stream>>>(T, ++p2p_epoch, p2p_err.p);
    EE_CUDA(cudaGetLastError());
    count_launch(5);
    accel_launches++;
    p2p_used = true;
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
    int32_t st = m < order ? starter_step() : steady_step();
    if (st) return st;
    if (solout) {
        st = solout->after_step(*this);
        if (st) return st;
    }
    return EE_OK;
}

int32_t NBodyEngine::step(int64_t nsteps) {
    EE_CUDA(cudaSetDevice(device));
    accel_launches = 0;
    EE_CUDA(cudaEventRecord(ev0, stream));
    int32_t st = EE_OK;
    const bool small = small_path_available(*this);
    int64_t s = 0;
    while (s < nsteps) {
        if (small && m >= order) {
            // persistent single-CTA path: as many steady-state steps per launch as the sample buffers allow
            int64_t k = std::min<int64_t>(nsteps - s, 1 << 16);
            if (solout) {
                if (solout->room() == 0) solout->flush(*this);
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
                if (solout) solout->begin_batch(*this);
                small_steps(*this, ok_steps);
                if (solout) solout->advance_host(ok_steps);
                m += ok_steps;
                t = tt;
                predicted = false;
                s += ok_steps;
            }
            if (st) break;
            continue;
        }
        st = step_once();
        if (st) break;
        ++s;
    }
    EE_CUDA(cudaEventRecord(ev1, stream));
    timed = true;
    return st;
}

void NBodyEngine::sync() {
    EE_CUDA(cudaSetDevice(device));
    EE_CUDA(cudaStreamSynchronize(stream));
}

void NBodyEngine::state(double* time, double* pos, double* vel, double* acc) {
    EE_CUDA(cudaSetDevice(device));
    if (acc) ensure_a0();
    if (time) *time = t;
    const bool local_only = world > 1 && exchange == EE_EXCHANGE_ALLGATHER;  // only the own slice is current
    const int64_t g0 = rank * (n / world);
    if (p2p_used) check_async_error();
    std::vector<double4> hp;
    std::vector<double> hv;
    if (pos) {
        hp.resize((size_t)n);
        EE_CUDA(cudaMemcpyAsync(hp.data(), ry.p + (size_t)slot_of(m) * n, (size_t)n * sizeof(double4), cudaMemcpyDeviceToHost,
                                stream));
        EE_CUDA(cudaStreamSynchronize(stream));
        for (int64_t k = 0; k < n; ++k) {
            pos[3 * k] = hp[(size_t)k].x;
            pos[3 * k + 1] = hp[(size_t)k].y;
            pos[3 * k + 2] = hp[(size_t)k].z;
        }
    }
    auto fetch_soa = [&](const double* src, double* dst) {
        // SoA [3][n] -> AoS; when sharded by targets only [i0,i1) is valid locally: gather across ranks first
        const double* from = src;
        if (local_only) {
            const int64_t per = n / world;
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
                    for (int64_t k = 0; k < per; ++k)
                        dst[3 * (r * per + k) + c] = hv[(size_t)((r * 3 + c) * per + k)];
            return;
        }
        hv.resize((size_t)3 * n);
        EE_CUDA(cudaMemcpyAsync(hv.data(), from, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        EE_CUDA(cudaStreamSynchronize(stream));
        for (int64_t k = 0; k < n; ++k)
            for (int c = 0; c < 3; ++c) dst[3 * k + c] = hv[(size_t)(c * n + k)];
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

NBodyEngine* NBodyEngine::clone() {
    EE_REQUIRE(world == 1, "clone of a sharded propagator is not supported");
    sync();
    std::vector<double> p((size_t)3 * n, 0.0), v((size_t)3 * n, 0.0), mu((size_t)n, 1.0);
    NBodyEngine* c = new NBodyEngine(n, p.data(), v.data(), mu.data(), t, h, method, mode, device, 0, 1, nullptr, 0);
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
};
static const uint64_t kSnapMagic = 0x45455f534e415031ull;  // "EE_SNAP1"

int64_t NBodyEngine::snapshot_bytes() const {
    return (int64_t)sizeof(SnapHeader) + (int64_t)(ry.bytes() + ra.bytes() + dy.bytes());
}

void NBodyEngine::snapshot(void* blob) {
    EE_REQUIRE(world == 1, "snapshot of a sharded propagator is not supported");
    EE_REQUIRE(!solout, "snapshot with a solout attached is not supported (use clone)");
    EE_CUDA(cudaSetDevice(device));
    SnapHeader hd{kSnapMagic, n, m, method, mode, order, R, have_a0 ? 1 : 0, predicted ? 1 : 0, t, h};
    unsigned char* p = (unsigned char*)blob;
    std::memcpy(p, &hd, sizeof(hd));
    p += sizeof(hd);
    EE_CUDA(cudaMemcpyAsync(p, ry.p, ry.bytes(), cudaMemcpyDeviceToHost, stream));
    p += ry.bytes();
    EE_CUDA(cudaMemcpyAsync(p, ra.p, ra.bytes(), cudaMemcpyDeviceToHost, stream));
    p += ra.bytes();
    EE_CUDA(cudaMemcpyAsync(p, dy.p, dy.bytes(), cudaMemcpyDeviceToHost, stream));
    EE_CUDA(cudaStreamSynchronize(stream));
}

void NBodyEngine::restore(const void* blob) {
    EE_REQUIRE(world == 1, "restore of a sharded propagator is not supported");
    EE_REQUIRE(!solout, "restore with a solout attached is not supported");
    EE_CUDA(cudaSetDevice(device));
    SnapHeader hd;
    std::memcpy(&hd, blob, sizeof(hd));
    EE_REQUIRE(hd.magic == kSnapMagic, "not a snapshot blob");
    EE_REQUIRE(hd.n == n && hd.method == method && hd.mode == mode && hd.R == R, "snapshot does not match this handle");
    const unsigned char* p = (const unsigned char*)blob + sizeof(hd);
    EE_CUDA(cudaMemcpyAsync(ry.p, p, ry.bytes(), cudaMemcpyHostToDevice, stream));
    p += ry.bytes();
    EE_CUDA(cudaMemcpyAsync(ra.p, p, ra.bytes(), cudaMemcpyHostToDevice, stream));
    p += ra.bytes();
    EE_CUDA(cudaMemcpyAsync(dy.p, p, dy.bytes(), cudaMemcpyHostToDevice, stream));
    m = hd.m;
    t = hd.t;
    h = hd.h;
    hs = h * (1.0 / 4.0);
    have_a0 = hd.have_a0 != 0;
    predicted = hd.predicted != 0;
}

// K steps, each bracketed by events, with an L2-evicting memset before each (outside the timed interval)
double NBodyEngine::step_timed(int64_t nsteps, int64_t flush_bytes, int32_t* status) {
    EE_CUDA(cudaSetDevice(device));
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
    cudaDeviceProp prop;
    EE_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
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

int64_t pair_items_total(int64_t n, int js) {
    const long long nt = n / kSymTile, ns = n / js;
    return sym_item_prefix(nt, ns, kSymTile / js);
}

void pair_item_decode(int64_t n, int js, int64_t item, int64_t* ti_out, int64_t* sj_out) {
    const long long nt = n / kSymTile, ns = n / js, ratio = kSymTile / js;
    long long lo = 0, hi = nt - 1;
    while (lo < hi) {  // same search as the kernels: largest tile whose first item is <= item
        const long long mid = (lo + hi + 1) >> 1;
        if (sym_item_prefix(mid, ns, ratio) <= item) lo = mid; else hi = mid - 1;
    }
    *ti_out = lo;
    *sj_out = ratio * lo + (item - sym_item_prefix(lo, ns, ratio));
}

// stand-alone NewtonianGravity::eval
void gravity_eval(int64_t n, const double* pos, const double* mus, int mode, int device, double* acc) {
    std::vector<double> vel((size_t)3 * n, 0.0);
    NBodyEngine e(n, pos, vel.data(), mus, 0.0, 1.0, EE_QUINLAN_TREMAINE_12, mode, device, 0, 1, nullptr, 0);
    e.state(nullptr, nullptr, nullptr, acc);
}

}  // namespace ee
