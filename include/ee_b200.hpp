// ee_b200.hpp -- header-only C++ mirror of the reference's propagator trait surface over the C ABI (ee_b200.h).
//
//   ee::NBodyPropagator        ephemeris::NBodyPropagator + Propagator / IncrementalPropagator / DirectionalPropagator /
//                              BoundedPropagator (ephemeris/src/lib.rs:9-79, propagators/nbody.rs:65-235)
//   ee::UniformSpline          ephemeris::UniformSpline<DVec3> (trajectory.rs:412-417)
//   ee::Ephemeris              Vec<UniformSpline<DVec3>> resident on the device (the ships' AccelerationModel context)
//   ee::SpacecraftPropagator   a batch of ephemeris::SpacecraftPropagator<T, .., M, ..>, M = any IntegrationMethod, with the
//                              app's SpacecraftSolout analytics and RelativeTrajectory sampling
//
// Errors the reference returns from `step()` (StepError / NBodyPropagatorError) surface as ee::StepError; engine
// failures (CUDA, NCCL, bad arguments) as ee::EngineError.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <iterator>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ee_b200.h"

namespace ee {

struct EngineError : std::runtime_error {
    int32_t code;
    EngineError(int32_t c, const std::string& where) : std::runtime_error(where + ": " + ee_last_error()), code(c) {}
};
struct StepError : std::runtime_error {  // integration/src/lib.rs:312-337
    int32_t code;
    explicit StepError(int32_t c)
        : std::runtime_error(c == 1   ? "step size underflow"
                             : c == 2 ? "max iterations reached"
                             : c == 3 ? "integration bound reached"
                             : c == 4 ? "failed to evaluate ODE"
                                      : "solout exit"),
          code(c) {}
};
inline void check(int32_t st, const char* where) {
    if (st == EE_OK) return;
    if (st < 100) throw StepError(st);
    throw EngineError(st, where);
}

using Vec3 = std::array<double, 3>;

// The synthetic Plummer sphere of the bench configurations (ee_host_plummer): the same bits in every host language.
struct SyntheticSystem {
    std::vector<Vec3> positions, velocities;
    std::vector<double> mus;
};
inline SyntheticSystem plummer(int64_t n, uint64_t seed = 20260924) {
    SyntheticSystem s;
    s.positions.resize((size_t)n);
    s.velocities.resize((size_t)n);
    s.mus.resize((size_t)n);
    check(ee_host_plummer(n, seed, s.positions.data()->data(), s.velocities.data()->data(), s.mus.data()), "ee_host_plummer");
    return s;
}

struct Polynomial {  // lowest-order coefficient first (trajectory.rs:339-340)
    std::vector<Vec3> coeffs;
};
// Container side of ephemeris::UniformSpline (trajectory.rs:427-447, :474-540, :571-617): what the Prediction Planner
// applies to every solution it takes (PredictionTarget::merge, dynamics/celestial.rs:194-235).  Evaluation stays on the
// device (ee_ephem_evaluate).
struct UniformSpline {
    double start = 0, interval = 0;
    std::vector<Polynomial> polynomials;
    double span() const { return interval * (double)polynomials.size(); }
    double end() const { return start + span(); }
    bool contains(double time) const {
        const double local = time - start;
        return !std::signbit(local) && local <= span();
    }
    size_t segment_count() const { return polynomials.size(); }

    // index rules; -1 = None.  `as usize` saturates, is_negative looks at the sign bit (ftime/src/duration.rs:83-85)
    int64_t get_index(double at) const {
        const double local = at - start;
        if (std::signbit(local) || local >= span()) return -1;
        return as_index(local / interval);
    }
    int64_t get_index_exclusive(double at) const {
        const double local = at - start;
        if (std::signbit(local) || local > span()) return -1;
        const int64_t i = as_index(std::ceil(local / interval));
        return i > 0 ? i - 1 : 0;
    }
    void push_front(Polynomial p) {
        polynomials.insert(polynomials.begin(), std::move(p));
        start -= interval;
    }
    void push_back(Polynomial p) { polynomials.push_back(std::move(p)); }
    void prepend(UniformSpline t) {
        if (!(start == t.start + t.span()) || !(interval == t.interval)) throw std::logic_error("prepend: the pieces do not meet");
        start = t.start;
        polynomials.insert(polynomials.begin(), std::make_move_iterator(t.polynomials.begin()),
                           std::make_move_iterator(t.polynomials.end()));
    }
    void append(UniformSpline t) {
        if (!(start + span() == t.start) || !(interval == t.interval)) throw std::logic_error("append: the pieces do not meet");
        polynomials.insert(polynomials.end(), std::make_move_iterator(t.polynomials.begin()),
                           std::make_move_iterator(t.polynomials.end()));
    }
    void clear_before(double at) {
        const int64_t idx = get_index_exclusive(at + interval);
        if (idx < 0) return;
        start += interval * (double)idx;
        polynomials.erase(polynomials.begin(), polynomials.begin() + idx);
    }
    void clear_after(double at) {
        const int64_t idx = get_index(at);
        if (idx >= 0) polynomials.resize((size_t)idx);
    }
    void merge_forward(UniformSpline propagated) {  // CelestialTrajectory<Forward>::merge
        clear_after(propagated.start);
        append(std::move(propagated));
    }
    void merge_backward(UniformSpline propagated) {  // CelestialTrajectory<Backward>::merge
        clear_before(propagated.end());
        prepend(std::move(propagated));
    }

  private:
    static int64_t as_index(double x) { return !(x > 0.0) ? 0 : x >= 9.2e18 ? INT64_MAX : (int64_t)x; }
};

// CubicHermiteSpline as a knot list (t, position, velocity).  `join` is the Planner's merge of a ship solution
// (SpacecraftPropagator::join, ephemeris/src/propagators/spacecraft.rs:558-561: clear_after(rhs.start()) + extend;
// clear_after keeps the knots strictly before `at`, trajectory.rs:835-838).
using Knot = std::array<double, 7>;
inline void join(std::vector<Knot>& lhs, const std::vector<Knot>& rhs) {
    if (!rhs.empty()) {
        const double at = rhs.front()[0];
        size_t keep = 0;
        for (const Knot& k : lhs)
            if (at > k[0]) lhs[keep++] = k;
        lhs.resize(keep);
    }
    lhs.insert(lhs.end(), rhs.begin(), rhs.end());
}

struct Forward {  // propagators/mod.rs:23-57
    double delta;
    double signed_delta() const { return delta < 0 ? -delta : delta; }
};
struct Backward {  // propagators/mod.rs:59-93
    double delta;
    double signed_delta() const { return delta < 0 ? delta : -delta; }
};

class Ephemeris {
  public:
    explicit Ephemeris(ee_ephem* h) : h_(h) {}
    Ephemeris(const Ephemeris&) = delete;
    Ephemeris(Ephemeris&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~Ephemeris() { ee_ephem_destroy(h_); }
    ee_ephem* raw() const { return h_; }

  private:
    ee_ephem* h_;
};

class NBodyPropagator {
  public:
    // NBodyPropagator::new (nbody.rs:93-121); Direction = Forward | Backward
    template <class Direction>
    NBodyPropagator(Direction direction, double initial_time, const std::vector<Vec3>& positions,
                    const std::vector<Vec3>& velocities, const std::vector<double>& gravitational_parameters,
                    int method = EE_QUINLAN_TREMAINE_12, int mode = EE_MODE_PARITY, int device = 0)
        : n_((int64_t)gravitational_parameters.size()) {
        check(ee_nbody_create(n_, positions.data()->data(), velocities.data()->data(), gravitational_parameters.data(),
                              initial_time, direction.signed_delta(), method, mode, device, &h_),
              "ee_nbody_create");
    }
    NBodyPropagator(const NBodyPropagator& o) : n_(o.n_) { check(ee_nbody_clone(o.h_, &h_), "ee_nbody_clone"); }  // Clone
    NBodyPropagator(NBodyPropagator&& o) noexcept : h_(o.h_), n_(o.n_) { o.h_ = nullptr; }
    ~NBodyPropagator() { ee_nbody_destroy(h_); }

    // SplineInterpolators::new (nbody.rs:332-340) with LeastSquaresFit{degree}
    void with_solout(double delta, const std::vector<double>& sample_periods, const std::vector<int32_t>& degrees) {
        check(ee_nbody_set_solout(h_, delta, sample_periods.data(), degrees.data()), "ee_nbody_set_solout");
    }
    void step(int64_t n_steps = 1) { check(ee_nbody_step(h_, n_steps), "ee_nbody_step"); }       // IncrementalPropagator::step
    void step_to(double time) { check(ee_nbody_step_to(h_, time), "ee_nbody_step_to"); }          // ::step_to
    std::vector<UniformSpline> propagate(double to) {                                             // BoundedPropagator
        step_to(to);
        return take_solution();
    }
    double time() const {  // DirectionalPropagator::time
        double t = 0;
        check(ee_nbody_solution_time(h_, &t), "ee_nbody_solution_time");
        return t;
    }
    bool has_reached(double time) const {
        int32_t r = 0;
        check(ee_nbody_has_reached(h_, time, &r), "ee_nbody_has_reached");
        return r != 0;
    }
    double delta() const { return ee_nbody_delta(h_); }
    // non-blocking read of problem.time / state.y / state.dy: the copies overlap the following steps; the vectors must
    // stay alive and untouched until state_wait()
    double state_async(std::vector<Vec3>& positions, std::vector<Vec3>& velocities) {
        positions.resize((size_t)n_);
        velocities.resize((size_t)n_);
        double t = 0.0;
        check(ee_nbody_state_async(h_, &t, positions.data()->data(), velocities.data()->data()), "ee_nbody_state_async");
        return t;
    }
    void state_wait() { check(ee_nbody_state_wait(h_), "ee_nbody_state_wait"); }
    // problem.time / state.y / state.dy
    double state(std::vector<Vec3>& positions, std::vector<Vec3>& velocities) const {
        positions.resize((size_t)n_);
        velocities.resize((size_t)n_);
        double t = 0;
        check(ee_nbody_state(h_, &t, positions.data()->data(), velocities.data()->data(), nullptr), "ee_nbody_state");
        return t;
    }
    std::vector<UniformSpline> take_solution() {  // Propagator::take_solution
        std::vector<int64_t> np((size_t)n_);
        check(ee_nbody_solution_sizes(h_, np.data()), "ee_nbody_solution_sizes");
        int64_t total = 0;
        for (auto c : np) total += c;
        std::vector<double> start((size_t)n_), interval((size_t)n_), coeffs((size_t)(total ? total : 1) * 27);
        std::vector<int32_t> nc((size_t)(total ? total : 1));
        check(ee_nbody_take_solution(h_, start.data(), interval.data(), coeffs.data(), nc.data()), "ee_nbody_take_solution");
        std::vector<UniformSpline> out((size_t)n_);
        size_t k = 0;
        for (int64_t b = 0; b < n_; ++b) {
            out[(size_t)b].start = start[(size_t)b];
            out[(size_t)b].interval = interval[(size_t)b];
            for (int64_t p = 0; p < np[(size_t)b]; ++p, ++k) {
                Polynomial poly;
                for (int i = 0; i < nc[k]; ++i)
                    poly.coeffs.push_back({coeffs[k * 27 + 3 * i], coeffs[k * 27 + 3 * i + 1], coeffs[k * 27 + 3 * i + 2]});
                out[(size_t)b].polynomials.push_back(std::move(poly));
            }
        }
        return out;
    }
    Ephemeris take_solution_ephemeris() {
        ee_ephem* e = nullptr;
        check(ee_nbody_take_solution_ephem(h_, &e), "ee_nbody_take_solution_ephem");
        return Ephemeris(e);
    }

  private:
    ee_nbody* h_ = nullptr;
    int64_t n_;
};

struct Burn {  // (start, end, ConstantThrust{acceleration, ReferenceFrame}) -- spacecraft.rs:30-57
    double start, end;
    Vec3 acceleration;
    int32_t reference = -1;  // body index (TNB frame relative to it) or -1 = inertial
};

class SpacecraftPropagator {
  public:
    // n x SpacecraftPropagator::new (spacecraft.rs:453-477)
    SpacecraftPropagator(const Ephemeris& context, const std::vector<double>& initial_times,
                         const std::vector<std::array<double, 6>>& initial_states, const ee_adaptive_params& params,
                         const std::vector<std::vector<Burn>>& timelines)
        : n_((int64_t)initial_states.size()) {
        std::vector<int64_t> off{0};
        std::vector<double> bs, be, ba;
        std::vector<int32_t> br;
        for (const auto& tl : timelines) {
            for (const auto& b : tl) {
                bs.push_back(b.start);
                be.push_back(b.end);
                ba.insert(ba.end(), b.acceleration.begin(), b.acceleration.end());
                br.push_back(b.reference);
            }
            off.push_back((int64_t)bs.size());
        }
        while ((int64_t)off.size() < n_ + 1) off.push_back((int64_t)bs.size());
        check(ee_ships_create(context.raw(), n_, initial_times.data(), initial_states.data()->data(), &params, off.data(),
                              bs.data(), be.data(), ba.data(), br.data(), &h_),
              "ee_ships_create");
    }
    ~SpacecraftPropagator() { ee_ships_destroy(h_); }
    SpacecraftPropagator(const SpacecraftPropagator&) = delete;
    void step_to(double time, int64_t max_steps = 1 << 14) { check(ee_ships_step_to(h_, time, max_steps), "ee_ships_step_to"); }
    // SpacecraftSolout (ephemeris_explorer/src/dynamics/spacecraft.rs:448-586): SOI transitions and apsides next to the knots
    struct Transition {
        double time;
        int32_t body;
    };
    struct Apsis {
        double time, distance;
        int32_t body, kind;  // kind 0 = periapsis, 1 = apoapsis
    };
    void enable_analytics(const std::vector<double>& soi_radius) {
        check(ee_ships_enable_analytics(h_, soi_radius.data()), "ee_ships_enable_analytics");
    }
    // read before take_solution(), which starts the next solution
    void analytics(std::vector<std::vector<Transition>>& transitions, std::vector<std::vector<Apsis>>& apsides) {
        std::vector<int32_t> ntr((size_t)n_), nap((size_t)n_);
        check(ee_ships_analytics_counts(h_, ntr.data(), nap.data()), "ee_ships_analytics_counts");
        std::vector<int64_t> to((size_t)n_ + 1, 0), ao((size_t)n_ + 1, 0);
        for (int64_t i = 0; i < n_; ++i) {
            to[(size_t)i + 1] = to[(size_t)i] + ntr[(size_t)i];
            ao[(size_t)i + 1] = ao[(size_t)i] + nap[(size_t)i];
        }
        std::vector<double> tt((size_t)to.back() + 1), at((size_t)ao.back() + 1), ad((size_t)ao.back() + 1);
        std::vector<int32_t> tb((size_t)to.back() + 1), ab((size_t)ao.back() + 1), ak((size_t)ao.back() + 1);
        check(ee_ships_read_analytics(h_, to.data(), tt.data(), tb.data(), ao.data(), at.data(), ad.data(), ab.data(), ak.data()),
              "ee_ships_read_analytics");
        transitions.assign((size_t)n_, {});
        apsides.assign((size_t)n_, {});
        for (int64_t i = 0; i < n_; ++i) {
            for (int64_t k = to[(size_t)i]; k < to[(size_t)i + 1]; ++k) transitions[(size_t)i].push_back({tt[(size_t)k], tb[(size_t)k]});
            for (int64_t k = ao[(size_t)i]; k < ao[(size_t)i + 1]; ++k)
                apsides[(size_t)i].push_back({at[(size_t)k], ad[(size_t)k], ab[(size_t)k], ak[(size_t)k]});
        }
    }
    // RelativeTrajectory::state_vector of one ship's spline w.r.t. a body (-1 = none), batched (trajectory.rs:315-335)
    void evaluate_relative(int64_t ship, int32_t reference, const std::vector<double>& times, std::vector<double>& pos,
                           std::vector<double>& vel, std::vector<int32_t>& ok) {
        pos.assign(times.size() * 3, 0.0);
        vel.assign(times.size() * 3, 0.0);
        ok.assign(times.size(), 0);
        check(ee_ships_evaluate_relative(h_, ship, reference, (int64_t)times.size(), times.data(), pos.data(), vel.data(), ok.data()),
              "ee_ships_evaluate_relative");
    }
    // CubicHermiteSpline knots (t, position, velocity) per ship
    std::vector<std::vector<std::array<double, 7>>> take_solution() {
        std::vector<int64_t> nk((size_t)n_), off((size_t)n_ + 1, 0);
        check(ee_ships_info(h_, nullptr, nullptr, nk.data(), nullptr, nullptr), "ee_ships_info");
        for (int64_t i = 0; i < n_; ++i) off[(size_t)i + 1] = off[(size_t)i] + nk[(size_t)i];
        std::vector<double> flat((size_t)off.back() * 7);
        check(ee_ships_take_knots(h_, off.data(), flat.data()), "ee_ships_take_knots");
        std::vector<std::vector<std::array<double, 7>>> out((size_t)n_);
        for (int64_t i = 0; i < n_; ++i)
            for (int64_t k = off[(size_t)i]; k < off[(size_t)i + 1]; ++k) {
                std::array<double, 7> kn;
                for (int c = 0; c < 7; ++c) kn[(size_t)c] = flat[(size_t)k * 7 + (size_t)c];
                out[(size_t)i].push_back(kn);
            }
        return out;
    }

  private:
    ee_ships* h_ = nullptr;
    int64_t n_;
};

}  // namespace ee
