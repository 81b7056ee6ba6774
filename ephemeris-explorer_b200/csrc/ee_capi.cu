// ee_capi.cu -- the extern "C" surface declared in include/ee_b200.h.  Every entry point catches C++ exceptions
// and turns them into status codes + ee_last_error().
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ee_engine.h"
#include "ee_ships.h"

using namespace ee;

struct ee_nbody {
    NBodyEngine* e;
};
struct ee_ephem {
    Ephem* e;
};
struct ee_ships {
    Ships* s;
};

namespace {
template <class F>
int32_t guarded(F&& f) {
    try {
        return f();
    } catch (const Error& err) {
        g_last_error = err.what();
        return err.code;
    } catch (const std::exception& err) {
        g_last_error = err.what();
        return EE_ERR_INVALID;
    }
}
#define EE_ARG(cond)                                                        \
    do {                                                                    \
        if (!(cond)) throw Error(EE_ERR_INVALID, "invalid argument: " #cond); \
    } while (0)
}  // namespace

extern "C" {

const char* ee_last_error(void) { return g_last_error.c_str(); }
int32_t ee_version(void) { return 100; }
uint64_t ee_launch_count(void) { return g_launch_count.load(); }

int32_t ee_set_pair_variant(int32_t variant) {
    if (variant != 0 && variant != 1) return EE_ERR_INVALID;
    g_pair_variant.store(variant);
    return EE_OK;
}

int64_t ee_host_sampling_stride(double delta, double period) { return sampling_stride(delta, period); }

namespace {
struct Xoshiro256ss {  // Blackman & Vigna's xoshiro256**, state from splitmix64(seed)
    uint64_t s[4];
    explicit Xoshiro256ss(uint64_t seed) {
        for (uint64_t& w : s) {
            uint64_t z = (seed += 0x9e3779b97f4a7c15ull);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            w = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * 0x1.0p-53; }  // [0, 1)
};
}  // namespace

int32_t ee_host_plummer(int64_t n, uint64_t seed, double* pos, double* vel, double* mus) {
    return guarded([&] {
        EE_ARG(n >= 1 && pos && vel && mus);
        Xoshiro256ss rng(seed);
        auto direction = [&](double len, double* out) {
            const double z = 2.0 * rng.uniform() - 1.0, phi = 2.0 * 3.14159265358979323846 * rng.uniform();
            const double sxy = std::sqrt(1.0 - z * z);
            out[0] = len * sxy * std::cos(phi);
            out[1] = len * sxy * std::sin(phi);
            out[2] = len * z;
        };
        for (int64_t i = 0; i < n; ++i) {
            double r;
            do {
                double u;
                do u = rng.uniform();
                while (u <= 0.0);
                r = 1.0 / std::sqrt(std::pow(u, -2.0 / 3.0) - 1.0);
            } while (!(r <= 20.0));
            double q;
            for (;;) {
                q = rng.uniform();
                const double y = 0.1 * rng.uniform();
                if (y < q * q * std::pow(1.0 - q * q, 3.5)) break;
            }
            direction(r, pos + 3 * i);
            direction(q * std::sqrt(2.0) * std::pow(1.0 + r * r, -0.25), vel + 3 * i);
            mus[i] = 1.0 / (double)n;
        }
        for (double* a : {pos, vel}) {  // equal masses: the centre of mass is the mean
            double m[3] = {0.0, 0.0, 0.0};
            for (int64_t i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) m[c] += a[3 * i + c];
            for (int c = 0; c < 3; ++c) m[c] /= (double)n;
            for (int64_t i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) a[3 * i + c] -= m[c];
        }
        return (int32_t)EE_OK;
    });
}

int32_t ee_host_pair_schedule(int64_t n, int32_t tile, int32_t spread, int32_t world, int32_t rank, int32_t max_chunks,
                              int64_t* units_total, int64_t* unit_lo, int64_t* unit_hi, int64_t* n_items, int32_t* items4,
                              int64_t items_cap, int32_t* row_slot) {
    return guarded([&] {
        EE_ARG(n > 0 && tile >= 256 && tile % 256 == 0 && n % tile == 0 && spread >= 1 && world >= 1 && rank >= 0 && rank < world &&
               max_chunks >= 1);
        const SymSchedule sc = build_sym_schedule(n, tile, spread, world, rank, max_chunks);
        if (units_total) *units_total = sc.u_total;
        if (unit_lo) *unit_lo = sc.u_lo;
        if (unit_hi) *unit_hi = sc.u_hi;
        if (n_items) *n_items = (int64_t)sc.items.size();
        if (items4) {
            EE_ARG(items_cap >= (int64_t)sc.items.size());
            std::memcpy(items4, sc.items.data(), sc.items.size() * sizeof(SymItem));
        }
        if (row_slot) std::memcpy(row_slot, sc.row_slot.data(), sc.row_slot.size() * sizeof(int));
        return (int32_t)EE_OK;
    });
}

int32_t ee_nccl_unique_id(void* out128) {
    return guarded([&] {
        EE_ARG(out128);
        nccl_unique_id(out128);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_create_sharded(int64_t n, const double* positions, const double* velocities, const double* mus, double t0,
                                double h_signed, int32_t method, int32_t mode, int32_t device, int32_t rank, int32_t world,
                                const void* unique_id128, int32_t exchange, ee_nbody** out) {
    return guarded([&] {
        EE_ARG(out);
        *out = nullptr;
        NBodyEngine* e =
            new NBodyEngine(n, positions, velocities, mus, t0, h_signed, method, mode, device, rank, world, unique_id128, exchange);
        *out = new ee_nbody{e};
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_create(int64_t n, const double* positions, const double* velocities, const double* mus, double t0,
                        double h_signed, int32_t method, int32_t mode, int32_t device, ee_nbody** out) {
    return ee_nbody_create_sharded(n, positions, velocities, mus, t0, h_signed, method, mode, device, 0, 1, nullptr, 0, out);
}

int32_t ee_nbody_p2p_export(ee_nbody* h, void* blob256) {
    return guarded([&] {
        EE_ARG(h && blob256);
        h->e->p2p_export(blob256);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_p2p_connect(ee_nbody* h, const void* all_blobs) {
    return guarded([&] {
        EE_ARG(h && all_blobs);
        h->e->p2p_connect(all_blobs);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_p2p_trace(ee_nbody* h, int32_t enable, double* mean_ms4, int64_t* steps) {
    return guarded([&] {
        EE_ARG(h);
        NBodyEngine& e = *h->e;
        if (mean_ms4)
            for (int k = 0; k < 4; ++k) mean_ms4[k] = e.p2p_trace_steps ? e.p2p_trace_ms[k] / (double)e.p2p_trace_steps : 0.0;
        if (steps) *steps = e.p2p_trace_steps;
        if ((enable != 0) != e.p2p_trace_on) {
            e.p2p_trace_on = enable != 0;
            for (double& v : e.p2p_trace_ms) v = 0.0;
            e.p2p_trace_steps = 0;
        }
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_set_solout(ee_nbody* h, double delta, const double* sample_periods, const int32_t* degrees) {
    return guarded([&] {
        EE_ARG(h && sample_periods && degrees);
        EE_ARG(h->e->m == 0);
        EE_CUDA(cudaSetDevice(h->e->device));
        h->e->solout.reset(new Solout(*h->e, delta, sample_periods, degrees));
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_step(ee_nbody* h, int64_t n_steps) {
    return guarded([&] {
        EE_ARG(h && n_steps >= 0);
        return h->e->step(n_steps);
    });
}

int32_t ee_nbody_step_to(ee_nbody* h, double epoch) {
    return guarded([&] {
        EE_ARG(h);
        if (!h->e->solout) throw Error(EE_ERR_INVALID, "step_to needs a solout (the spline solution defines `time`)");
        EE_CUDA(cudaSetDevice(h->e->device));
        for (;;) {  // IncrementalPropagator::step_to -- ephemeris/src/lib.rs:47-58
            if (h->e->solout->has_reached(epoch)) return (int32_t)EE_OK;
            // the stopping step is known in advance from the sampling schedule, so the steps go out in batches
            const int64_t k = h->e->solout->steps_until(epoch);
            if (k >= ((int64_t)1 << 62)) throw Error(EE_ERR_INVALID, "step_to: a body never completes a sample period");
            int32_t st = h->e->step(std::max<int64_t>(1, k));
            if (st) return st;
        }
    });
}

int32_t ee_nbody_sync(ee_nbody* h) {
    return guarded([&] {
        EE_ARG(h);
        h->e->sync();
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_state(ee_nbody* h, double* time, double* positions, double* velocities, double* accelerations) {
    return guarded([&] {
        EE_ARG(h);
        h->e->state(time, positions, velocities, accelerations);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_state_async(ee_nbody* h, double* time, double* positions, double* velocities) {
    return guarded([&] {
        EE_ARG(h);
        h->e->state_async(time, positions, velocities, nullptr);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_state_wait(ee_nbody* h) {
    return guarded([&] {
        EE_ARG(h);
        h->e->state_wait();
        return (int32_t)EE_OK;
    });
}

double ee_nbody_delta(const ee_nbody* h) { return h ? h->e->h : 0.0; }
int64_t ee_nbody_step_count(const ee_nbody* h) { return h ? h->e->m : 0; }

int32_t ee_nbody_solution_time(ee_nbody* h, double* epoch) {
    return guarded([&] {
        EE_ARG(h && epoch && h->e->solout);
        *epoch = h->e->solout->solution_time();
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_has_reached(ee_nbody* h, double epoch, int32_t* reached) {
    return guarded([&] {
        EE_ARG(h && reached && h->e->solout);
        *reached = h->e->solout->has_reached(epoch) ? 1 : 0;
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_solution_sizes(ee_nbody* h, int64_t* n_poly) {
    return guarded([&] {
        EE_ARG(h && n_poly && h->e->solout);
        for (int64_t b = 0; b < h->e->n; ++b) n_poly[b] = h->e->solout->n_poly(b);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_take_solution(ee_nbody* h, double* start, double* interval, double* coeffs, int32_t* n_coef) {
    return guarded([&] {
        EE_ARG(h && start && interval && h->e->solout);
        EE_CUDA(cudaSetDevice(h->e->device));
        HostSolution sol;
        h->e->solout->take(*h->e, sol);
        std::memcpy(start, sol.start.data(), sol.start.size() * 8);
        std::memcpy(interval, sol.interval.data(), sol.interval.size() * 8);
        if (!sol.coeffs.empty()) {
            EE_ARG(coeffs && n_coef);
            std::memcpy(coeffs, sol.coeffs.data(), sol.coeffs.size() * 8);
            std::memcpy(n_coef, sol.n_coef.data(), sol.n_coef.size() * 4);
        }
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_take_solution_ephem(ee_nbody* h, ee_ephem** out) {
    return guarded([&] {
        EE_ARG(h && out && h->e->solout);
        EE_CUDA(cudaSetDevice(h->e->device));
        HostSolution sol;
        h->e->solout->take(*h->e, sol);
        std::vector<double> mu((size_t)h->e->n);
        {
            std::vector<double4> hp((size_t)h->e->n);
            EE_CUDA(cudaMemcpy(hp.data(), h->e->positions_dev(), hp.size() * sizeof(double4), cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < hp.size(); ++k) mu[k] = hp[k].w;
        }
        Ephem* e = new Ephem(h->e->n, mu.data(), sol.start.data(), sol.interval.data(), sol.n_poly.data(), sol.coeffs.data(),
                             sol.n_coef.data(), h->e->device);
        *out = new ee_ephem{e};
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_snapshot_size(const ee_nbody* h, int64_t* bytes) {
    return guarded([&] {
        EE_ARG(h && bytes);
        *bytes = h->e->snapshot_bytes();
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_snapshot(ee_nbody* h, void* blob) {
    return guarded([&] {
        EE_ARG(h && blob);
        h->e->snapshot(blob);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_restore(ee_nbody* h, const void* blob, int64_t blob_bytes) {
    return guarded([&] {
        EE_ARG(h && blob && blob_bytes > 0);
        h->e->restore(blob, blob_bytes);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_step_timed(ee_nbody* h, int64_t n_steps, int64_t flush_bytes, double* total_ms) {
    return guarded([&] {
        EE_ARG(h && total_ms && n_steps >= 0 && flush_bytes >= 0);
        int32_t st = EE_OK;
        *total_ms = h->e->step_timed(n_steps, flush_bytes, &st);
        return st;
    });
}

int32_t ee_fp64_fma_peak(int32_t device, double* tflops) {
    return guarded([&] {
        EE_ARG(tflops);
        *tflops = fp64_fma_peak(device);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_clone(ee_nbody* h, ee_nbody** out) {
    return guarded([&] {
        EE_ARG(h && out);
        *out = new ee_nbody{h->e->clone()};
        return (int32_t)EE_OK;
    });
}

void ee_nbody_destroy(ee_nbody* h) {
    if (!h) return;
    delete h->e;
    delete h;
}

int32_t ee_gravity_eval(int64_t n, const double* positions, const double* mus, int32_t mode, int32_t device,
                        double* accelerations) {
    return guarded([&] {
        EE_ARG(positions && mus && accelerations);
        gravity_eval(n, positions, mus, mode, device, accelerations);
        return (int32_t)EE_OK;
    });
}

int32_t ee_nbody_last_timing(const ee_nbody* h, double* ms, int64_t* launches) {
    return guarded([&] {
        EE_ARG(h);
        if (ms) *ms = h->e->last_step_ms();
        if (launches) *launches = h->e->accel_launches;
        return (int32_t)EE_OK;
    });
}

// ---- ephemeris ---------------------------------------------------------------------------------------------
int32_t ee_ephem_create(int64_t n_bodies, const double* mus, const double* start, const double* interval,
                        const int64_t* n_poly, const double* coeffs, const int32_t* n_coef, int32_t device, ee_ephem** out) {
    return guarded([&] {
        EE_ARG(out && mus && start && interval && n_poly);
        *out = nullptr;
        Ephem* e = new Ephem(n_bodies, mus, start, interval, n_poly, coeffs, n_coef, device);
        *out = new ee_ephem{e};
        return (int32_t)EE_OK;
    });
}

int32_t ee_ephem_evaluate(ee_ephem* e, int64_t n_times, const double* times, double* positions, double* velocities,
                          int32_t* ok) {
    return guarded([&] {
        EE_ARG(e);
        e->e->evaluate(n_times, times, positions, velocities, ok);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ephem_sizes(ee_ephem* e, int64_t* n_bodies, int64_t* n_poly) {
    return guarded([&] {
        EE_ARG(e);
        if (n_bodies) *n_bodies = e->e->nb;
        if (n_poly) std::memcpy(n_poly, e->e->n_poly.data(), (size_t)e->e->nb * 8);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ephem_get(ee_ephem* e, double* mus, double* start, double* interval, double* coeffs, int32_t* n_coef) {
    return guarded([&] {
        EE_ARG(e);
        Ephem& E = *e->e;
        EE_CUDA(cudaSetDevice(E.device));
        if (mus) std::memcpy(mus, E.mu.data(), (size_t)E.nb * 8);
        if (start) std::memcpy(start, E.start.data(), (size_t)E.nb * 8);
        if (interval) std::memcpy(interval, E.interval.data(), (size_t)E.nb * 8);
        if (coeffs && E.total) EE_CUDA(cudaMemcpy(coeffs, E.coef.p, (size_t)E.total * 27 * 8, cudaMemcpyDeviceToHost));
        if (n_coef && E.total) EE_CUDA(cudaMemcpy(n_coef, E.ncoef.p, (size_t)E.total * 4, cudaMemcpyDeviceToHost));
        return (int32_t)EE_OK;
    });
}

void ee_ephem_destroy(ee_ephem* e) {
    if (!e) return;
    delete e->e;
    delete e;
}

int32_t ee_lsq_fit(int64_t n_fits, const int32_t* degrees, const double* ts9, const double* samples, int32_t device,
                   double* coeffs, int32_t* n_coef) {
    return guarded([&] {
        EE_ARG(degrees && ts9 && samples && coeffs && n_coef);
        lsq_fit_batch(n_fits, degrees, ts9, samples, device, coeffs, n_coef);
        return (int32_t)EE_OK;
    });
}

// ---- ships -------------------------------------------------------------------------------------------------
int32_t ee_ships_create(ee_ephem* ephem, int64_t n_ships, const double* t0, const double* states,
                        const ee_adaptive_params* params, const int64_t* burn_offsets, const double* burn_start,
                        const double* burn_end, const double* burn_acc, const int32_t* burn_ref, ee_ships** out) {
    return guarded([&] {
        EE_ARG(ephem && out);
        *out = nullptr;
        Ships* s = new Ships(ephem->e, n_ships, t0, states, params, burn_offsets, burn_start, burn_end, burn_acc, burn_ref);
        *out = new ee_ships{s};
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_step_to(ee_ships* h, double t_end, int64_t max_steps) {
    return guarded([&] {
        EE_ARG(h);
        h->s->step_to(t_end, max_steps);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_info(ee_ships* h, int32_t* status, double* time, int64_t* n_knots, uint32_t* n_attempts, uint64_t* rhs_evals) {
    return guarded([&] {
        EE_ARG(h);
        h->s->info(status, time, n_knots, n_attempts, rhs_evals);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_take_knots(ee_ships* h, const int64_t* knot_offsets, double* knots7) {
    return guarded([&] {
        EE_ARG(h && knot_offsets && knots7);
        h->s->take_knots(knot_offsets, knots7);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_enable_analytics(ee_ships* h, const double* soi_radius) {
    return guarded([&] {
        EE_ARG(h && soi_radius);
        h->s->enable_analytics(soi_radius);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_analytics_counts(ee_ships* h, int32_t* n_transitions, int32_t* n_apsides) {
    return guarded([&] {
        EE_ARG(h);
        h->s->analytics_counts(n_transitions, n_apsides);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_read_analytics(ee_ships* h, const int64_t* transition_offsets, double* transition_time, int32_t* transition_body,
                                const int64_t* apsis_offsets, double* apsis_time, double* apsis_distance, int32_t* apsis_body,
                                int32_t* apsis_kind) {
    return guarded([&] {
        EE_ARG(h && transition_offsets && apsis_offsets);
        h->s->read_analytics(transition_offsets, transition_time, transition_body, apsis_offsets, apsis_time, apsis_distance,
                             apsis_body, apsis_kind);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ephem_evaluate_relative(ee_ephem* e, int32_t body, int32_t reference, int64_t n_times, const double* times, double* pos,
                                   double* vel, int32_t* ok) {
    return guarded([&] {
        EE_ARG(e);
        e->e->evaluate_relative(body, reference, n_times, times, pos, vel, ok);
        return (int32_t)EE_OK;
    });
}

int32_t ee_ships_evaluate_relative(ee_ships* h, int64_t ship, int32_t reference, int64_t n_times, const double* times, double* pos,
                                   double* vel, int32_t* ok) {
    return guarded([&] {
        EE_ARG(h);
        h->s->evaluate_relative(ship, reference, n_times, times, pos, vel, ok);
        return (int32_t)EE_OK;
    });
}

double ee_ships_last_ms(ee_ships* h) { return h ? h->s->last_ms : 0.0; }

void ee_ships_destroy(ee_ships* h) {
    if (!h) return;
    delete h->s;
    delete h;
}

}  // extern "C"
