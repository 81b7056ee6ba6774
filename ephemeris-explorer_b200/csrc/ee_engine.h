// ee_engine.h -- host-side classes behind the C ABI (declarations only; no device code).
#pragma once
#include <memory>

#include "ee_common.cuh"
#include "ee_sym_types.h"

namespace ee {

struct EpArgs;
struct QtArgs;
struct NBodyEngine;
struct Ephem;

void nccl_unique_id(void* out128);
int64_t sampling_stride(double delta, double period);

// Work list of the pair-symmetric kernel for one rank (host logic, no CUDA call): see ee_sym.cuh.
struct SymSchedule {
    long long u_total = 0, u_lo = 0, u_hi = 0;  // canonical units: all, and this rank's range
    std::vector<SymItem> items;                 // queue order = canonical order, guided (decreasing) sizes
    std::vector<int> row_slot;                  // [nt + 1] prefix: slots of tile row ti are [row_slot[ti], row_slot[ti+1])
};
SymSchedule build_sym_schedule(int64_t n, int tile, int spread, int world, int rank, int max_chunks);
bool small_path_available(const NBodyEngine& e);
void small_steps(NBodyEngine& e, int64_t m0, int64_t steps_done0, int64_t k);
double fp64_fma_peak(int device);
void gravity_eval(int64_t n, const double* pos, const double* mus, int mode, int device, double* acc);
void lsq_fit_batch(int64_t n_fits, const int32_t* degrees, const double* ts9, const double* samples, int device,
                   double* coeffs, int32_t* n_coef);

// Host-side description of a spline solution (Vec<UniformSpline<DVec3>>), coefficients zero-padded to 9 per polynomial.
struct HostSolution {
    std::vector<double> start, interval;
    std::vector<int64_t> n_poly;
    std::vector<double> coeffs;  // sum(n_poly) * 27
    std::vector<int32_t> n_coef;
};

// SplineInterpolators<D, DVec3, LeastSquaresFit> + its Solution -- ephemeris/src/propagators/nbody.rs:309-517
struct Solout {
    int64_t n = 0;
    double delta = 0.0;
    bool backward = false;
    std::vector<double> period;
    std::vector<int32_t> degree;
    // reference interpolator state, mirrored on the host (exact f64 arithmetic of nbody.rs:389-396)
    std::vector<double> last_sample_time;
    std::vector<int64_t> stride;   // steps between samples (0 = the f64 accumulation never hits the period)
    std::vector<int64_t> since;    // steps since the last sample
    // device sample buffers
    int64_t steps_done = 0;        // steps since the solout was attached
    std::vector<int64_t> off, cap, held, qbase;
    DBuf<double> samples;          // [sum cap][3]
    DBuf<int64_t> d_stride, d_off, d_qbase;
    bool dirty_meta = true;
    // fitted polynomials: append-only pool on the device
    DBuf<double> pool;             // [pool_cap][27]
    DBuf<int32_t> pool_nc;
    int64_t pool_cap = 0, pool_len = 0;
    std::vector<std::vector<std::pair<int64_t, int64_t>>> segs;  // per body (pool offset, count), chronological
    std::vector<int64_t> done;     // polynomials fitted so far per body
    // Solution meta (UniformSpline::start at new_solution time)
    std::vector<double> sol_start, sol_interval;

    Solout(NBodyEngine& e, double delta, const double* periods, const int32_t* degrees);
    int32_t after_step(NBodyEngine& e);
    int64_t room() const;                          // steps that fit in the sample buffers before a flush is needed
    void begin_batch(NBodyEngine& e);              // make the device-side sampling tables current
    void advance_host(int64_t k);                  // account for k steps whose samples a kernel wrote itself
    void flush(NBodyEngine& e);                    // fit every complete 9-sample group, compact the buffers
    void new_solution(const NBodyEngine& e);       // nbody.rs:455-468
    int64_t n_poly(int64_t b) const { return done[(size_t)b] + (held[(size_t)b] - 1) / 8; }
    double bound(int64_t b) const;                 // SplineBound::bound -- nbody.rs:411-443
    double solution_time() const;                  // nbody.rs:501-508
    bool has_reached(double epoch) const;          // nbody.rs:510-516
    int64_t steps_until(double epoch) const;       // steps after which has_reached(epoch) first holds (0 = already)
    void take(NBodyEngine& e, HostSolution& out);  // Propagator::take_solution -- nbody.rs:181-189
    Solout* clone(NBodyEngine& owner) const;
    int64_t blob_bytes() const;                                   // serialised size (host state + device buffers)
    void save(NBodyEngine& e, unsigned char* out);
    static Solout* load(NBodyEngine& e, const unsigned char* in, int64_t bytes);
    double interp_time(int64_t b) const;           // SplineInterpolator::time -- nbody.rs:318-322

  private:
    Solout() = default;
    void upload_meta(NBodyEngine& e);
    void grow_pool(NBodyEngine& e, int64_t need);
};

struct NBodyEngine {
    int64_t n;
    int method, mode, device;
    int order = 12, R = 13;
    double h = 0, hs = 0, t = 0, bound = 0;
    int64_t m = 0;  // completed steps; state of step s lives in ring slot s % R
    bool have_a0 = false, predicted = false;
    // small-system run-ahead: steps accounted on the host (m, t, sampling schedule) but not launched yet
    int64_t pending = 0, batch_k = 0;
    cudaEvent_t batch_ev[2] = {nullptr, nullptr};
    void flush_pending();
    bool device_idle();
    // sharding
    int rank = 0, world = 1, exchange = 0;
    void* comm = nullptr;
    int64_t i0 = 0, i1 = 0, j0 = 0, j1 = 0;
    // launch plan (throughput kernel)
    int sm_count = 0, block = 128, tiles = 1, splits = 1;
    int64_t chunk = 0;
    // device state
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int64_t accel_launches = 0;
    DBuf<double4> ry;
    DBuf<double> ra, dy, a_scr, part;
    DBuf<double4> ytmp[2];
    DBuf<unsigned> tickets;
    // symmetric (Newton's third law) throughput path
    bool use_sym = false;
    int sym_ti = 4, sym_nt = 256, sym_minb = 2, sym_sbc = 16;  // kernel variant: targets/lane, threads/CTA, CTAs/SM, chunks/sub-block
    int sym_n_items = 0;
    SymShare sym_share{};
    DBuf<SymItem> sym_items;
    DBuf<int> sym_row_slot;
    DBuf<double> sym_part_i, sym_part_j;
    DBuf<unsigned> sym_counter;
    size_t sym_part_i_count = 0, sym_part_j_count = 0;
    bool scratch_ready = false;
    void ensure_scratch();
    void launch_sym(const double4* y_in, const EpArgs& ep);
    // NVLink peer path (CUDA IPC): see ee_sym.cuh
    bool p2p_ready = false, p2p_used = false;
    unsigned long long p2p_epoch = 0;
    DBuf<unsigned long long> p2p_flags;
    int* p2p_err_h = nullptr;             // sticky error flag of the peer barriers: pinned host memory, device-mapped
    int* p2p_err_d = nullptr;             // device alias of p2p_err_h
    DBuf<int> p2p_err_dev;                // the same flag in device memory (what the kernels poll)
    void* p2p_table = nullptr;            // PeerTable (host copy)
    std::vector<void*> p2p_opened;        // cudaIpcOpenMemHandle results to close
    bool p2p_trace_on = false;            // ee_nbody_p2p_trace: per-launch event timing of the peer step
    cudaEvent_t p2p_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double p2p_trace_ms[4] = {0, 0, 0, 0};  // pair units + local reduce | barrier | slice finish | barrier
    int64_t p2p_trace_steps = 0;
    void p2p_export(void* blob256);
    void p2p_connect(const void* all_blobs);
    void p2p_step(const EpArgs& ep);
    void check_async_error();
    std::unique_ptr<Solout> solout;

    NBodyEngine(int64_t n, const double* pos, const double* vel, const double* mus, double t0, double h_signed, int method,
                int mode, int device, int rank, int world, const void* uid, int exchange);
    ~NBodyEngine();
    NBodyEngine(const NBodyEngine&) = delete;
    void init(const double* pos, const double* vel, const double* mus, double t0, double h_signed, const void* uid);
    void release_all();

    int slot_of(int64_t s) const { return (int)(((s % R) + R) % R); }
    const double4* positions_dev() const { return ry.p + (size_t)slot_of(m) * n; }
    void plan_launch();
    void accel(const double4* y_in, EpArgs ep);
    void epilogue_from(const double* a_in, EpArgs ep);
    void exchange_y(double4* buf);
    QtArgs qt_args(int64_t newest, int64_t next) const;
    void ensure_a0();
    int pair_variant = 0;    // reading of particular's pair kernel (parity kernels), fixed at creation: ee_set_pair_variant
    bool srkn_main = false;  // method = BlanesMoan14A: every step is an SRKN step
    int32_t srkn_step(int stages, const double* CA, const double* CB, int substeps, double h_sub);
    int32_t starter_step();
    int32_t steady_step();
    int32_t step_once();
    int32_t step(int64_t nsteps);
    int32_t step_small(int64_t nsteps);
    void sync();
    void state(double* time, double* pos, double* vel, double* acc);
    void state_async(double* time, double* pos, double* vel, double* acc);
    void state_wait();
    cudaStream_t copy_stream = nullptr;
    DBuf<double> stage_d[2];              // [9][n]: AoS positions, velocities, accelerations
    cudaEvent_t stage_packed[2] = {nullptr, nullptr}, stage_copied[2] = {nullptr, nullptr};
    int64_t stage_k = 0;
    double last_step_ms();
    NBodyEngine* clone();
    int64_t snapshot_bytes() const;
    void snapshot(void* blob);
    void restore(const void* blob, int64_t blob_bytes);
    double step_timed(int64_t nsteps, int64_t flush_bytes, int32_t* status);
    DBuf<unsigned char> flush_buf;
};

// Device-resident ephemeris table: per body a uniform spline of polynomials with <= 9 DVec3 coefficients.
struct Ephem {
    int device = 0;
    int64_t nb = 0, total = 0;
    std::vector<double> mu, start, interval;
    std::vector<int64_t> n_poly, first;  // first[b] = index of body b's first polynomial in the table
    DBuf<double> coef;                   // [total][27]
    DBuf<int32_t> ncoef;                 // [total]
    DBuf<double> d_mu, d_start, d_interval;
    DBuf<int64_t> d_npoly, d_first;
    Ephem(int64_t nb, const double* mus, const double* start, const double* interval, const int64_t* n_poly,
          const double* coeffs, const int32_t* n_coef, int device);
    void evaluate(int64_t n_times, const double* times, double* pos, double* vel, int32_t* ok);
    void evaluate_relative(int body, int reference, int64_t n_times, const double* times, double* pos, double* vel, int32_t* ok);
};

}  // namespace ee
