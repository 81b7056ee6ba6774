// planner_loop.cpp -- the Prediction Planner's call pattern driven through the public C ABI (include/ee_b200.hpp):
//   loop { propagator.step(); if sync.is_ready() || reached { take_solution(); clone(); } }      prediction.rs:408-446
// with Synchronisation::hertz(100) (planner.rs:93; load/mod.rs:675).  bench.py writes the system to a small binary file
// and reads one JSON line back.  The polynomials of every take_solution() are folded into one FNV-1a hash per body so
// that the result can be compared with the CPU oracle running the same loop (take timing differs, the polynomials do not).
//
//   planner_loop <input.bin> [device]
//   input.bin: int64 n; double t0, h, end, tick_seconds; double pos[n][3], vel[n][3], mu[n], period[n]; int32 degree[n]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "ee_b200.hpp"

static uint64_t fnv(uint64_t h, const void* p, size_t bytes) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < bytes; ++i) {
        h ^= c[i];
        h *= 1099511628211ull;
    }
    return h;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: planner_loop <input.bin> [device]\n");
        return 2;
    }
    const int device = argc > 2 ? std::atoi(argv[2]) : 0;
    std::ifstream f(argv[1], std::ios::binary);
    int64_t n = 0;
    double t0, h, end, tick;
    f.read((char*)&n, 8);
    f.read((char*)&t0, 8);
    f.read((char*)&h, 8);
    f.read((char*)&end, 8);
    f.read((char*)&tick, 8);
    std::vector<ee::Vec3> pos((size_t)n), vel((size_t)n);
    std::vector<double> mu((size_t)n), period((size_t)n);
    std::vector<int32_t> degree((size_t)n);
    f.read((char*)pos.data(), n * 24);
    f.read((char*)vel.data(), n * 24);
    f.read((char*)mu.data(), n * 8);
    f.read((char*)period.data(), n * 8);
    f.read((char*)degree.data(), n * 4);
    if (!f) {
        std::fprintf(stderr, "short input file\n");
        return 2;
    }
    try {
        using clk = std::chrono::steady_clock;
        ee::NBodyPropagator prop(ee::Forward{h}, t0, pos, vel, mu, EE_QUINLAN_TREMAINE_12, EE_MODE_PARITY, device);
        prop.with_solout(h, period, degree);
        prop.step(12);  // Blanes-Moan start-up + CUDA module load stay outside the timed loop (the oracle arm does the same)
        std::vector<uint64_t> hash((size_t)n, 1469598103934665603ull);
        int64_t steps = 12, ticks = 0, polys = 0;
        const auto tstart = clk::now();
        auto last = tstart;
        for (;;) {
            prop.step();
            ++steps;
            const bool reached = prop.has_reached(end);
            const auto now = clk::now();
            if (std::chrono::duration<double>(now - last).count() >= tick || reached) {
                auto sol = prop.take_solution();
                for (int64_t b = 0; b < n; ++b)
                    for (const auto& p : sol[(size_t)b].polynomials) {
                        const int32_t nc = (int32_t)p.coeffs.size();
                        hash[(size_t)b] = fnv(hash[(size_t)b], &nc, 4);
                        hash[(size_t)b] = fnv(hash[(size_t)b], p.coeffs.data(), p.coeffs.size() * 24);
                        ++polys;
                    }
                ee::NBodyPropagator snapshot(prop);  // PredictionResult::new: propagator.clone()
                (void)snapshot;
                ++ticks;
                last = clk::now();
                if (reached) break;
            }
        }
        const double secs = std::chrono::duration<double>(clk::now() - tstart).count();
        uint64_t all = 1469598103934665603ull;
        for (auto v : hash) all = fnv(all, &v, 8);
        std::printf("{\"steps\": %lld, \"seconds\": %.6f, \"steps_per_s\": %.1f, \"ticks\": %lld, \"polynomials\": %lld, "
                    "\"hash\": \"%016llx\", \"gpu_launches\": %llu}\n",
                    (long long)(steps - 12), secs, (double)(steps - 12) / secs, (long long)ticks, (long long)polys,
                    (unsigned long long)all, (unsigned long long)ee_launch_count());
    } catch (const std::exception& e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    return 0;
}
