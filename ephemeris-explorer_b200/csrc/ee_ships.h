// ee_ships.h -- host-side handle of a batch of massless ships.
#pragma once
#include "ee_engine.h"

namespace ee {

struct Ships {
    Ephem* ephem;
    int64_t n;
    ee_adaptive_params params;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    DBuf<double> d_time, d_bound, d_state, d_next_h, d_seg_end, d_seg_acc, knots;
    DBuf<uint32_t> d_rk_i, d_natt;
    DBuf<int32_t> d_cur, d_status, d_seg_burn, d_seg_ref;
    DBuf<int64_t> d_nknots, d_seg_off;
    DBuf<unsigned long long> d_evals;
    int64_t kcap = 0, max_held = 1;
    DBuf<double> d_fsal_k;  // [n][6] last slope of an FSAL method between launches
    DBuf<double> scratch;   // position + polynomial caches of the kernel when they do not fit in shared memory (> ~100 bodies)
    // SpacecraftSolout analytics (ephemeris_explorer/src/dynamics/spacecraft.rs:448-586), optional
    bool analytics = false;
    int64_t tr_cap = 0, ap_cap = 0, max_tr = 1, max_ap = 0;
    DBuf<double> d_soi_r, d_tr_time, d_ap_time, d_ap_dist;
    DBuf<int32_t> d_tr_body, d_ap_body, d_ap_kind, d_ntr, d_nap;
    void enable_analytics(const double* soi_radius);
    void reset_analytics();
    void ensure_analytics_capacity();
    void analytics_counts(int32_t* n_tr, int32_t* n_ap);
    void read_analytics(const int64_t* tr_off, double* tr_time, int32_t* tr_body, const int64_t* ap_off, double* ap_time,
                        double* ap_distance, int32_t* ap_body, int32_t* ap_kind);

    Ships(Ephem* eph, int64_t n, const double* t0, const double* states, const ee_adaptive_params* p, const int64_t* burn_off,
          const double* bstart, const double* bend, const double* bacc, const int32_t* bref);
    ~Ships();
    Ships(const Ships&) = delete;
    void ensure_capacity(int64_t extra);
    void step_to(double t_end, int64_t max_steps);
    void info(int32_t* status, double* time, int64_t* n_knots, uint32_t* n_attempts, uint64_t* rhs_evals);
    void take_knots(const int64_t* offsets, double* out);
    void evaluate_relative(int64_t ship, int reference, int64_t nt, const double* times, double* pos, double* vel, int32_t* ok);
};

}  // namespace ee
