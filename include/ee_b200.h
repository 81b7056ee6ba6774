/* ee_b200.h -- C ABI of the B200-native ephemeris engine (libee_b200.so).
 *
 * Drop-in boundary for ONE hot path of Canleskis/ephemeris-explorer: the fixed-step high-order stepper driving the
 * all-pairs Newtonian acceleration, and the massless-ship propagator sampling the resulting spline ephemeris.
 * The reference has no FFI of its own (it is a pure-Rust workspace); its seam is the set of generic traits the
 * Prediction Planner requires of a propagator (ephemeris_explorer/src/prediction.rs:31-37).  Each entry point below
 * names the reference interface it stands in for (file:line under the reference tree); a Rust shim maps them back
 * onto those traits (shim/).
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every array; vectors are AoS double[3] exactly like Vec<DVec3>.
 *   - units as the reference: km, km/s, km^3/s^2; time = f64 seconds since 1958-01-01 TAI (ftime/src/epoch.rs:3-7).
 *   - every function returns an ee_status.  0..5 mirror integration::StepError / *PropagatorError
 *     (integration/src/lib.rs:312-318, ephemeris/src/propagators/nbody.rs:43-47); >= 100 are engine errors whose
 *     text is available from ee_last_error().
 *   - a handle is not thread-safe but may move between threads (one stepping thread at a time, as
 *     prediction.rs:385-391 uses a propagator).
 *   - there is NO CPU fallback: without a CUDA device every create call fails with EE_ERR_CUDA.
 */
#ifndef EE_B200_H
#define EE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ee_nbody ee_nbody; /* NBodyPropagator<D, DVec3, M, SplineInterpolators<D>>  (nbody.rs:65-68)   */
typedef struct ee_ephem ee_ephem; /* Vec<UniformSpline<DVec3>> + mus, device resident       (trajectory.rs:412-417) */
typedef struct ee_ships ee_ships; /* a batch of SpacecraftPropagator<T, .., M, ..> (spacecraft.rs:415-426), M = any IntegrationMethod */

typedef enum ee_status {
    EE_OK = 0,
    EE_STEP_SIZE_UNDERFLOW = 1,    /* StepError::StepSizeUnderflow                      */
    EE_MAX_ITERATIONS_REACHED = 2, /* StepError::MaxIterationsReached                   */
    EE_BOUND_REACHED = 3,          /* StepError::BoundReached                           */
    EE_EVAL_FAILED = 4,            /* StepError::EvalFailed (ship left the ephemeris)   */
    EE_SOLOUT_EXIT = 5,            /* NBodyPropagatorError::Solout                      */
    EE_ERR_INVALID = 100,
    EE_ERR_CUDA = 101,
    EE_ERR_NCCL = 102,
    EE_ERR_UNSUPPORTED = 103
} ee_status;

/* integration::methods aliases (integration/src/methods.rs:37-40) */
typedef enum ee_method { EE_QUINLAN_TREMAINE_12 = 12, EE_STORMER_13 = 13, EE_BLANES_MOAN_14A = 14 } ee_method;

/* EE_MODE_PARITY reproduces the reference's floating-point evaluation order bit for bit (pair loop order of
 * nbody.rs:22-38, left-to-right linear combinations, no FMA contraction, IEEE sqrt/div).
 * EE_MODE_THROUGHPUT uses FMA, an rsqrt seed + cubic refinement and tiled summation (<= 1e-12 relative over short
 * runs; documented in DESIGN.md). */
typedef enum ee_mode { EE_MODE_PARITY = 0, EE_MODE_THROUGHPUT = 1 } ee_mode;

/* how shards exchange per acceleration evaluation when world > 1 */
typedef enum ee_exchange {
    EE_EXCHANGE_ALLREDUCE = 0, /* sources sharded, ncclAllReduce(sum) of 3N partial accelerations            */
    EE_EXCHANGE_ALLGATHER = 1  /* targets sharded, ncclAllGather of the new positions (rank-count invariant) */
} ee_exchange;

const char* ee_last_error(void);
int32_t ee_version(void);
/* The pair force lives in the un-vendored crate `particular` (0.8.0-dev @ d490707a, Cargo.lock:4278-4280): parity mode
 * follows its published scalar form dir * (mu / (n * sqrt(n))) (variant 0, default).  Variant 1 is the other plausible
 * reading -- one reciprocal, two products: dir * (mu * (1 / (n * sqrt(n)))).  Applies to handles created afterwards;
 * both variants are tested bit for bit against the oracle's twin switch. */
int32_t ee_set_pair_variant(int32_t variant);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t ee_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * n-body propagator
 * ---------------------------------------------------------------------------------------------------------- */

/* NBodyPropagator::new(direction, initial_time, positions, velocities, gravitational_parameters, solout)
 * (ephemeris/src/propagators/nbody.rs:93-121).  h_signed = direction.signed_delta() (negative = Backward,
 * propagators/mod.rs:80-82).  device = CUDA ordinal. */
int32_t ee_nbody_create(int64_t n, const double* positions, const double* velocities, const double* mus, double t0,
                        double h_signed, int32_t method, int32_t mode, int32_t device, ee_nbody** out);

/* Same, as rank `rank` of `world` processes (one per GPU) over NCCL.  unique_id = 128 bytes from ee_nccl_unique_id()
 * on rank 0, distributed by the host (bench.py uses torch.distributed).  Every rank passes the full body arrays. */
int32_t ee_nccl_unique_id(void* out128);
int32_t ee_nbody_create_sharded(int64_t n, const double* positions, const double* velocities, const double* mus,
                                double t0, double h_signed, int32_t method, int32_t mode, int32_t device, int32_t rank,
                                int32_t world, const void* unique_id128, int32_t exchange, ee_nbody** out);

/* NVLink peer path for a sharded handle (throughput mode, EE_EXCHANGE_ALLREDUCE layout, n >= 32768): instead of
 * ncclAllReduce + a separate epilogue, each rank finishes its slice of bodies by reading every rank's partial
 * accelerations over NVLink, runs the integrator epilogue and stores the new positions into all peers' rings, all in
 * one kernel, with flag barriers in peer memory (csrc/ee_sym.cuh).  ee_nbody_p2p_export writes 512 bytes of CUDA-IPC handles; the host
 * all-gathers the blobs (world x 512 bytes, rank order) and passes them to ee_nbody_p2p_connect on every rank.
 * The G partials are added in rank order (deterministic; <= 1e-12 from the 1-GPU run). */
int32_t ee_nbody_p2p_export(ee_nbody* h, void* blob512);
int32_t ee_nbody_p2p_connect(ee_nbody* h, const void* all_blobs);
/* Diagnosis of the peer path: with enable != 0 every following step is timed launch by launch with CUDA events (and
 * synchronised, so it is not the fast path).  The call returns the means since tracing was switched on, in ms:
 * [pair units + local reduce, flag barrier, slice finish (peer loads, epilogue, peer stores), flag barrier], then
 * switches tracing to `enable`.  Every rank must switch in the same step. */
int32_t ee_nbody_p2p_trace(ee_nbody* h, int32_t enable, double* mean_ms4, int64_t* steps);

/* SplineInterpolators::new(delta, [SplineInterpolator{ZERO, sample_period_b, PolyonmialInterpolator::new(pos_b),
 * LeastSquaresFit{degree_b}}]) + Integration::with_solout (nbody.rs:332-340, dynamics/celestial.rs:156-186,
 * integration/src/lib.rs:435-445).  Call before the first step. */
int32_t ee_nbody_set_solout(ee_nbody* h, double delta, const double* sample_periods, const int32_t* degrees);

/* n_steps x IncrementalPropagator::step (nbody.rs:200-207).  Asynchronous on the handle's stream; errors that
 * the reference would return from a step (underflow, bound) are detected on the host before launch. */
int32_t ee_nbody_step(ee_nbody* h, int64_t n_steps);
/* IncrementalPropagator::step_to (ephemeris/src/lib.rs:47-58): step until has_reached(epoch). */
int32_t ee_nbody_step_to(ee_nbody* h, double epoch);
/* Block until all queued steps are done (and surface asynchronous errors). */
int32_t ee_nbody_sync(ee_nbody* h);

/* problem.time / problem.state.{y,dy} (nbody.rs:150-153; integration/src/problem.rs:101-112).  Any pointer may
 * be NULL.  acc = the integrator's current_ddy.  Synchronises. */
int32_t ee_nbody_state(ee_nbody* h, double* time, double* positions, double* velocities, double* accelerations);
/* The same read without blocking the caller: the state after the steps queued so far is packed on the device and
 * copied into `positions` / `velocities` (n x 3 doubles each, may be NULL) on a separate copy stream while later steps
 * run; the arrays must stay valid and untouched until ee_nbody_state_wait returns.  Page-locked memory gives true
 * overlap (pageable memory is staged by the driver).  At most two reads are in flight; a third waits for the first.
 * Not available on target-sharded (EE_EXCHANGE_ALLGATHER) handles. */
int32_t ee_nbody_state_async(ee_nbody* h, double* time, double* positions, double* velocities);
int32_t ee_nbody_state_wait(ee_nbody* h);
/* NBodyPropagator::delta (nbody.rs:155-161) */
double ee_nbody_delta(const ee_nbody* h);
/* Integration::step_count (integration/src/lib.rs:461-464) */
int64_t ee_nbody_step_count(const ee_nbody* h);

/* DirectionalPropagator::time / has_reached on the spline solution (nbody.rs:226-234, :501-516). */
int32_t ee_nbody_solution_time(ee_nbody* h, double* epoch);
int32_t ee_nbody_has_reached(ee_nbody* h, double epoch, int32_t* reached);

/* Propagator::take_solution (nbody.rs:181-189) in two calls: sizes first, then the move.
 * n_poly[b] = polynomials in body b's UniformSpline.  coeffs = sum(n_poly) x 9 x 3 doubles, lowest order first,
 * zero padded; n_coef = sum(n_poly) coefficient counts after Polynomial::trim (trajectory.rs:387-395). */
int32_t ee_nbody_solution_sizes(ee_nbody* h, int64_t* n_poly);
int32_t ee_nbody_take_solution(ee_nbody* h, double* start, double* interval, double* coeffs, int32_t* n_coef);
/* Same solution, but kept on the device as an ephemeris table for ee_ships_* (no host round trip). */
int32_t ee_nbody_take_solution_ephem(ee_nbody* h, ee_ephem** out);

/* Checkpoint / resume (in the reference "the propagator IS the checkpoint": every snapshot sends
 * propagator.clone() back, prediction.rs:213-230, and `extend` resumes from it, :378).  The blob holds the complete
 * multistep state (time, step index, the ring of positions and accelerations, velocities) in HOST memory and, when a
 * solout is attached, its sampling schedule, pending samples and fitted polynomials; restore needs a handle created
 * with the same n / method / mode and replaces its solout by the blob's.  Sharded handles whose ranks all hold the
 * complete state (EE_EXCHANGE_ALLREDUCE layout, peer path included) snapshot locally and restore the same blob on
 * every rank.  ee_nbody_snapshot_size must be asked right before ee_nbody_snapshot (the size grows with the solution).
 * These two calls are also the host-buffer path bench.py times as `e2e`. */
int32_t ee_nbody_snapshot_size(const ee_nbody* h, int64_t* bytes);
int32_t ee_nbody_snapshot(ee_nbody* h, void* blob);
int32_t ee_nbody_restore(ee_nbody* h, const void* blob, int64_t blob_bytes); /* a short or foreign blob is refused */

/* K steps timed one by one on the device: before every step `flush_bytes` of scratch are overwritten (outside the
 * timed interval) to evict the L2, then the step is bracketed by CUDA events on the handle's stream.  total_ms is the
 * sum of the K step intervals.  Used by bench.py; semantics otherwise identical to ee_nbody_step. */
int32_t ee_nbody_step_timed(ee_nbody* h, int64_t n_steps, int64_t flush_bytes, double* total_ms);

/* Sustained FP64 FMA rate of the device measured with independent DFMA chains (2 flop per FMA), for the roofline
 * denominator of the all-pairs kernel. */
int32_t ee_fp64_fma_peak(int32_t device, double* tflops);

/* Clone (prediction.rs:224-229 clones the propagator at every snapshot): device-to-device, solout included, no scratch
 * allocated until the clone is stepped.  Cloning a sharded handle (EE_EXCHANGE_ALLREDUCE layout) gives an UNSHARDED
 * replica of the same state on the calling rank's GPU. */
int32_t ee_nbody_clone(ee_nbody* h, ee_nbody** out);
void ee_nbody_destroy(ee_nbody* h);

/* One acceleration evaluation, NewtonianGravity::eval (nbody.rs:16-39), for tests and kernel timing. */
int32_t ee_gravity_eval(int64_t n, const double* positions, const double* mus, int32_t mode, int32_t device,
                        double* accelerations);

/* per-kernel timing of the last ee_nbody_step call: total device ms and launches of the dominant kernel */
int32_t ee_nbody_last_timing(const ee_nbody* h, double* accel_kernel_ms, int64_t* accel_kernel_launches);

/* Host-side logic exposed for the CPU test-suite (no CUDA call inside):
 *  - ee_host_sampling_stride: steps between samples under the reference's exact rule `last_sample_time += delta;
 *    last_sample_time == sample_period` (nbody.rs:389-391); 0 = the accumulation never hits the period.
 *  - ee_host_pair_schedule: the pair-symmetric kernel's work list of rank `rank` of `world` for n bodies cut into I-tiles
 *    of `tile` bodies and j-chunks of 32: the canonical unit range [unit_lo, unit_hi) of units_total, and the item table
 *    (items4[k] = {tile row, first chunk, chunks, slot}; guided sizes = remaining / spread rounded down to a power of two,
 *    <= max_chunks; queue order = canonical order) with
 *    row_slot[n/tile + 1] = prefix of items per tile row.  Any output pointer may be NULL. */
int64_t ee_host_sampling_stride(double delta, double period);
/* The synthetic Plummer sphere of the bench configurations (nothing in the reference generates one): G = 1,
 * total mu = 1, mu_i = 1/n, scale radius 1; r = 1/sqrt(u^(-2/3) - 1) (r > 20 rejected); isotropic direction; speed
 * q sqrt(2) (1 + r^2)^(-1/4) with q by von Neumann rejection on q^2 (1 - q^2)^3.5; isotropic direction; centre of mass
 * position and velocity removed.  RNG: xoshiro256** seeded by four splitmix64 outputs of `seed`, uniform = (x >> 11) 2^-53.
 * One implementation for every host language, so every host regenerates the bench input bit for bit. */
int32_t ee_host_plummer(int64_t n, uint64_t seed, double* positions, double* velocities, double* mus);
int32_t ee_host_pair_schedule(int64_t n, int32_t tile, int32_t spread, int32_t world, int32_t rank, int32_t max_chunks,
                              int64_t* units_total, int64_t* unit_lo, int64_t* unit_hi, int64_t* n_items, int32_t* items4,
                              int64_t items_cap, int32_t* row_slot);

/* ------------------------------------------------------------------------------------------------------------
 * ephemeris table (piecewise-polynomial, uniform intervals)
 * ---------------------------------------------------------------------------------------------------------- */

/* Vec<UniformSpline<DVec3>> -> device.  Layout as ee_nbody_take_solution. */
int32_t ee_ephem_create(int64_t n_bodies, const double* mus, const double* start, const double* interval,
                        const int64_t* n_poly, const double* coeffs, const int32_t* n_coef, int32_t device,
                        ee_ephem** out);
/* UniformSpline::position / state_vector (trajectory.rs:449-471) batched: for each of n_times epochs and every body.
 * positions/velocities: n_times x n_bodies x 3; ok: n_times x n_bodies (0 = None).  velocities may be NULL. */
int32_t ee_ephem_evaluate(ee_ephem* e, int64_t n_times, const double* times, double* positions, double* velocities,
                          int32_t* ok);
int32_t ee_ephem_sizes(ee_ephem* e, int64_t* n_bodies, int64_t* n_poly);
int32_t ee_ephem_get(ee_ephem* e, double* mus, double* start, double* interval, double* coeffs, int32_t* n_coef);
void ee_ephem_destroy(ee_ephem* e);

/* LeastSquaresFit::interpolate (ephemeris_explorer/src/dynamics/celestial.rs:24-136) batched on the device:
 * n_fits x (9 samples of DVec3) at ts -> n_fits x 9 x 3 coefficients + counts. */
int32_t ee_lsq_fit(int64_t n_fits, const int32_t* degrees, const double* ts9, const double* samples, int32_t device,
                   double* coeffs, int32_t* n_coef);

/* ------------------------------------------------------------------------------------------------------------
 * massless ships
 * ---------------------------------------------------------------------------------------------------------- */

/* AdaptiveMethodParams<f64, AbsTol, f64> (integration/src/lib.rs:174-197; load/mod.rs:472-486) */
typedef struct ee_adaptive_params {
    double h_init, h_max;
    double tol_position, tol_velocity; /* AbsTol (dynamics/spacecraft.rs:609-613) */
    double fac_min, fac_max, fac;
    uint32_t n_max;
    /* how `err.powf(-1/k)` of the step-size controller (integration/src/runge_kutta/mod.rs:238) is evaluated:
     * EE_POW_GLIBC (0, default): glibc's pow, operation for operation -- what Rust's f64::powf resolves to on
     *   Linux/x86-64, i.e. the reference as built (csrc/ee_pow_glibc.h);
     * EE_POW_CORRECTLY_ROUNDED (1): the engine's libm-independent double-double pow (csrc/ee_pow.cuh). */
    uint32_t pow_mode;
    /* which AdaptiveRungeKutta method integrates the ships: the reference's IntegrationMethod
     * (ephemeris_explorer/src/flight_plan.rs:175-184; tableaux integration/src/methods.rs:92-1658).  0 = Verner87, the
     * method of every shipped flight plan.  Fine45 is the ERKNG (second-order, y'' = f(t, y, y')) form
     * (integration/src/runge_kutta/nystrom/explicit_generalized.rs:77-138), the others ERK
     * (integration/src/runge_kutta/explicit.rs:73-107). */
    uint32_t method;
} ee_adaptive_params;
enum { EE_POW_GLIBC = 0, EE_POW_CORRECTLY_ROUNDED = 1 };
enum {
    EE_SHIP_VERNER87 = 0,
    EE_SHIP_CASH_KARP45 = 1,
    EE_SHIP_DORMAND_PRINCE54 = 2,
    EE_SHIP_DORMAND_PRINCE87 = 3,
    EE_SHIP_FEHLBERG45 = 4,
    EE_SHIP_TSITOURAS75 = 5,
    EE_SHIP_VERNER98 = 6,
    EE_SHIP_FINE45 = 7
};

/* n_ships x SpacecraftPropagator::new(initial_time, initial_state, params, timeline, context, solout)
 * (ephemeris/src/propagators/spacecraft.rs:453-477) with M = Verner87, T = [StateVector;1].
 * states = n_ships x 6 (position, velocity).  Burns are flattened: burn_offsets[n_ships+1] indexes
 * burn_start/burn_end (epochs), burn_acc (x3, km/s^2 in the burn frame) and burn_ref (body index for a TNB frame
 * relative to that body, -1 = inertial; dynamics/spacecraft.rs:254-293).  The ephemeris must outlive the ships. */
int32_t ee_ships_create(ee_ephem* ephem, int64_t n_ships, const double* t0, const double* states,
                        const ee_adaptive_params* params, const int64_t* burn_offsets, const double* burn_start,
                        const double* burn_end, const double* burn_acc, const int32_t* burn_ref, ee_ships** out);
/* Each ship: IncrementalPropagator::step_to(t_end) (ephemeris/src/lib.rs:47-58) capped at max_steps accepted steps
 * per call.  A ship that errors keeps its status and stops, like prediction.rs:429-432 truncates a prediction.
 * A ship may also stop short with status OK -- the step cap, or (with analytics enabled) its transition/apsis lists
 * have to grow, which happens at the start of the next call: call again while ee_ships_info reports time < t_end. */
int32_t ee_ships_step_to(ee_ships* h, double t_end, int64_t max_steps);
/* per-ship: status (ee_status), problem.time, accepted steps so far, attempts n, rhs evaluations */
int32_t ee_ships_info(ee_ships* h, int32_t* status, double* time, int64_t* n_knots, uint32_t* n_attempts,
                      uint64_t* rhs_evals);
/* Propagator::take_solution for CubicHermiteSplineSolout (spacecraft.rs:645-695): knots = (t, pos, vel) x 7 doubles.
 * knot_offsets[n_ships+1] must come from ee_ships_info's n_knots (prefix sum). */
int32_t ee_ships_take_knots(ee_ships* h, const int64_t* knot_offsets, double* knots7);
/* Switch the ships' solution from CubicHermiteSplineSolout to the app's SpacecraftSolout
 * (ephemeris_explorer/src/dynamics/spacecraft.rs:448-586): besides the knots, every accepted step is searched on its
 * cubic-Hermite segment for sphere-of-influence crossings of every body (find_soi_crossing, bisection to 1 ms,
 * :91-161) and for apsides inside the spheres occupied during the step (find_apsis).  soi_radius[n_bodies] are the
 * bodies' SphereOfInfluence radii (INFINITY for the root body, :28-40).  Call before the first step (or right after
 * ee_ships_take_knots): it performs new_solution's `soi_at(now)` look-up. */
int32_t ee_ships_enable_analytics(ee_ships* h, const double* soi_radius);
/* per-ship number of SoiTransitions entries and Apsides entries held */
int32_t ee_ships_analytics_counts(ee_ships* h, int32_t* n_transitions, int32_t* n_apsides);
/* SoiTransitions = (time, body index) sorted by time; Apsides = (time, distance, body index, kind: 0 periapsis /
 * 1 apoapsis) sorted by time.  Offsets [n_ships+1] are prefix sums of ee_ships_analytics_counts.  Reads without
 * consuming; ee_ships_take_knots starts the new solution (transitions = [(now, soi_at(now))], no apsides). */
int32_t ee_ships_read_analytics(ee_ships* h, const int64_t* transition_offsets, double* transition_time,
                                int32_t* transition_body, const int64_t* apsis_offsets, double* apsis_time,
                                double* apsis_distance, int32_t* apsis_body, int32_t* apsis_kind);
/* RelativeTrajectory::state_vector (ephemeris/src/trajectory.rs:187-200, :315-335) batched over times -- the evaluation
 * the trajectory plotter makes for every candidate point (ephemeris_explorer/src/ui/world/plot.rs:326-334): trajectory(t)
 * minus reference(t), positions and velocities; ok[i] = 0 where either side is None.  The trajectory is body `body` of the
 * ephemeris (ee_ephem_evaluate_relative) or ship `ship`'s CubicHermiteSpline as held on the device since the last
 * ee_ships_take_knots (trajectory.rs:779-795); `reference` is a body index or -1 (no reference). */
int32_t ee_ephem_evaluate_relative(ee_ephem* e, int32_t body, int32_t reference, int64_t n_times, const double* times,
                                   double* pos, double* vel, int32_t* ok);
int32_t ee_ships_evaluate_relative(ee_ships* h, int64_t ship, int32_t reference, int64_t n_times, const double* times,
                                   double* pos, double* vel, int32_t* ok);
/* device milliseconds of the last ee_ships_step_to launch (CUDA events on the handle's stream) */
double ee_ships_last_ms(ee_ships* h);
void ee_ships_destroy(ee_ships* h);

#ifdef __cplusplus
}
#endif
#endif /* EE_B200_H */
