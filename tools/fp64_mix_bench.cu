// fp64_mix_bench.cu -- developer microbenchmark (not part of the product): what limits the FP64 pipe on the rotation body
// of k_accel_sym?  Every variant runs the same 80 FP64 instructions per rotation (4 targets) and reports FP64 warp
// instructions per cycle per SM sub-partition (the pipe accepts one every 2 cycles: 0.5 = saturated).
//   V0  as the kernel: MUFU.RSQ64H seed, j data + j-side accumulators in shared memory
//   V1  no MUFU (integer-ALU seed of the same bit width), shared memory as V0
//   V2  MUFU, no shared memory traffic (j data in registers)
//   V3  no MUFU, no shared memory
//   V4  pure independent DFMA chains (the peak probe)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ double seed_mufu(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double seed_int(double x) {  // integer pipe only; numerically useless, same data flow
    const int hi = __double2hiint(x);
    return __hiloint2double(0x5fe6eb50 - (hi >> 1), 0);
}

template <int TI, bool MUFU, bool SMEM>
__global__ void __launch_bounds__(128, 3) k_mix(const double4* __restrict__ pm, double* __restrict__ out, int rot) {
    __shared__ double sx[4][32], sy[4][32], sz[4][32], sm[4][32], wa[4][3][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double xi[TI], yi[TI], zi[TI], mi[TI], ax[TI], ay[TI], az[TI];
#pragma unroll
    for (int t = 0; t < TI; ++t) {
        const double4 p = pm[(blockIdx.x * 128 + threadIdx.x) * TI + t];
        xi[t] = p.x; yi[t] = p.y; zi[t] = p.z; mi[t] = p.w;
        ax[t] = ay[t] = az[t] = 0.0;
    }
    const double4 q = pm[lane];
    sx[warp][lane] = q.x; sy[warp][lane] = q.y; sz[warp][lane] = q.z; sm[warp][lane] = q.w;
    wa[warp][0][lane] = wa[warp][1][lane] = wa[warp][2][lane] = 0.0;
    __syncwarp();
    double rx = q.x, ry = q.y, rz = q.z, rm = q.w, cx = 0, cy = 0, cz = 0;
    int jj = lane;
#pragma unroll 2
    for (int k = 0; k < rot; ++k) {
        double xj, yj, zj, mj, bx, by, bz;
        if (SMEM) {
            xj = sx[warp][jj]; yj = sy[warp][jj]; zj = sz[warp][jj]; mj = sm[warp][jj];
            bx = wa[warp][0][jj]; by = wa[warp][1][jj]; bz = wa[warp][2][jj];
        } else {
            xj = rx; yj = ry; zj = rz; mj = rm; bx = cx; by = cy; bz = cz;
        }
#pragma unroll
        for (int g = 0; g < TI; g += 2) {
            double dx[2], dy[2], dz[2], r2[2], y0[2], y2[2], e[2], pp[2], c[2], rc[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) { dx[u] = xj - xi[g + u]; dy[u] = yj - yi[g + u]; dz[u] = zj - zi[g + u]; }
#pragma unroll
            for (int u = 0; u < 2; ++u) r2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], dx[u] * dx[u]));
#pragma unroll
            for (int u = 0; u < 2; ++u) y0[u] = MUFU ? seed_mufu(r2[u]) : seed_int(r2[u]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                y2[u] = y0[u] * y0[u];
                e[u] = fma(-r2[u], y2[u], 1.0);
                c[u] = y2[u] * y0[u];
                pp[u] = fma(1.875, e[u], 1.5);
                pp[u] = e[u] * pp[u];
                rc[u] = fma(c[u], pp[u], c[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double si = mj * rc[u], sj = mi[g + u] * rc[u];
                ax[g + u] = fma(si, dx[u], ax[g + u]);
                ay[g + u] = fma(si, dy[u], ay[g + u]);
                az[g + u] = fma(si, dz[u], az[g + u]);
                bx = fma(-sj, dx[u], bx);
                by = fma(-sj, dy[u], by);
                bz = fma(-sj, dz[u], bz);
            }
        }
        if (SMEM) {
            wa[warp][0][jj] = bx; wa[warp][1][jj] = by; wa[warp][2][jj] = bz;
            __syncwarp();
        } else {
            cx = bx; cy = by; cz = bz;
        }
        jj = (jj + 1) & 31;
    }
    double s = cx + cy + cz + wa[warp][0][lane];
#pragma unroll
    for (int t = 0; t < TI; ++t) s += ax[t] + ay[t] + az[t];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

__global__ void __launch_bounds__(128, 3) k_dfma(double* out, int rot, double seed) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = seed + threadIdx.x + u;
    const double b = 0.999999, c = 1e-9;
    for (int k = 0; k < rot; ++k) {
#pragma unroll
        for (int r = 0; r < 10; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = fma(a[u], b, c);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += a[u];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

// operand-bandwidth probes: 8 independent chains per thread, 80 FP64 instructions per outer iteration
template <int KIND>
__global__ void __launch_bounds__(128, 3) k_ops(const double* __restrict__ in, double* out, int rot) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        a[u] = in[threadIdx.x + 128 * u];
        b[u] = in[threadIdx.x + 128 * (8 + u)];
        c[u] = in[threadIdx.x + 128 * (16 + u)];
    }
    for (int k = 0; k < rot; ++k) {
#pragma unroll
        for (int r = 0; r < 10; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (KIND == 0) a[u] = fma(a[u], b[u], c[u]);        // 3 distinct register operands
                if (KIND == 1) a[u] = fma(a[u], b[u], a[u]);        // 2 distinct
                if (KIND == 2) a[u] = a[u] + b[u];                  // DADD, 2 distinct
                if (KIND == 3) a[u] = a[u] * b[u];                  // DMUL, 2 distinct
                if (KIND == 4) a[u] = fma(a[u], b[(u + r) & 7], c[(u + 2 * r) & 7]);  // 3 distinct, rotating partners
                if (KIND == 5) a[u] = fma(b[0], c[u], a[u]);        // 3 distinct, one operand shared by consecutive instructions
                if (KIND == 6) a[u] = fma(b[u >> 2 << 2], c[u], a[u]);  // shared within groups of 4 (the kernel's triples are 3)
            }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += a[u];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

// rotation body variants: MODE 0 = as the kernel; 1 = no j-side accumulate (62 FP64 instructions per rotation, no
// read-modify-write of shared memory); 2 = as the kernel without __syncwarp; 3 = j data rotated through shuffles instead
// of shared-memory loads (j-side accumulators still in shared memory); 4 = as the kernel, 4 rotations unrolled
template <int MODE>
__global__ void __launch_bounds__(128, 3) k_body(const double4* __restrict__ pm, double* __restrict__ out, int rot) {
    __shared__ double sx[4][32], sy[4][32], sz[4][32], sm[4][32], wa[4][3][32];
    constexpr int TI = 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double xi[TI], yi[TI], zi[TI], mi[TI], ax[TI], ay[TI], az[TI];
#pragma unroll
    for (int t = 0; t < TI; ++t) {
        const double4 p = pm[(blockIdx.x * 128 + threadIdx.x) * TI + t];
        xi[t] = p.x; yi[t] = p.y; zi[t] = p.z; mi[t] = p.w;
        ax[t] = ay[t] = az[t] = 0.0;
    }
    const double4 q = pm[lane];
    sx[warp][lane] = q.x; sy[warp][lane] = q.y; sz[warp][lane] = q.z; sm[warp][lane] = q.w;
    wa[warp][0][lane] = wa[warp][1][lane] = wa[warp][2][lane] = 0.0;
    __syncwarp();
    double rx = q.x, ry = q.y, rz = q.z, rm = q.w;
    int jj = lane;
#pragma unroll(MODE == 4 ? 4 : 2)
    for (int k = 0; k < rot; ++k) {
        double xj, yj, zj, mj, bx = 0, by = 0, bz = 0;
        if (MODE == 3) {
            xj = rx; yj = ry; zj = rz; mj = rm;
        } else {
            xj = sx[warp][jj]; yj = sy[warp][jj]; zj = sz[warp][jj]; mj = sm[warp][jj];
        }
        if (MODE != 1) { bx = wa[warp][0][jj]; by = wa[warp][1][jj]; bz = wa[warp][2][jj]; }
#pragma unroll
        for (int g = 0; g < TI; g += 2) {
            double dx[2], dy[2], dz[2], r2[2], y0[2], y2[2], e[2], pp[2], c[2], rc[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) { dx[u] = xj - xi[g + u]; dy[u] = yj - yi[g + u]; dz[u] = zj - zi[g + u]; }
#pragma unroll
            for (int u = 0; u < 2; ++u) r2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], dx[u] * dx[u]));
#pragma unroll
            for (int u = 0; u < 2; ++u) y0[u] = seed_mufu(r2[u]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                y2[u] = y0[u] * y0[u];
                e[u] = fma(-r2[u], y2[u], 1.0);
                c[u] = y2[u] * y0[u];
                pp[u] = fma(1.875, e[u], 1.5);
                pp[u] = e[u] * pp[u];
                rc[u] = fma(c[u], pp[u], c[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double si = mj * rc[u];
                if (MODE == 5) {
                    asm volatile("fma.rn.f64 %0, %3, %4, %0;\n\tfma.rn.f64 %1, %3, %5, %1;\n\tfma.rn.f64 %2, %3, %6, %2;"
                                 : "+d"(ax[g + u]), "+d"(ay[g + u]), "+d"(az[g + u])
                                 : "d"(si), "d"(dx[u]), "d"(dy[u]), "d"(dz[u]));
                    const double sj = -(mi[g + u] * rc[u]);
                    asm volatile("fma.rn.f64 %0, %3, %4, %0;\n\tfma.rn.f64 %1, %3, %5, %1;\n\tfma.rn.f64 %2, %3, %6, %2;"
                                 : "+d"(bx), "+d"(by), "+d"(bz)
                                 : "d"(sj), "d"(dx[u]), "d"(dy[u]), "d"(dz[u]));
                    continue;
                }
                ax[g + u] = fma(si, dx[u], ax[g + u]);
                ay[g + u] = fma(si, dy[u], ay[g + u]);
                az[g + u] = fma(si, dz[u], az[g + u]);
                if (MODE != 1) {
                    const double sj = mi[g + u] * rc[u];
                    bx = fma(-sj, dx[u], bx);
                    by = fma(-sj, dy[u], by);
                    bz = fma(-sj, dz[u], bz);
                }
            }
        }
        if (MODE != 1) {
            wa[warp][0][jj] = bx; wa[warp][1][jj] = by; wa[warp][2][jj] = bz;
            if (MODE != 2) __syncwarp();
        }
        if (MODE == 3) {
            rx = __shfl_sync(0xffffffffu, rx, (lane + 1) & 31);
            ry = __shfl_sync(0xffffffffu, ry, (lane + 1) & 31);
            rz = __shfl_sync(0xffffffffu, rz, (lane + 1) & 31);
            rm = __shfl_sync(0xffffffffu, rm, (lane + 1) & 31);
        }
        jj = (jj + 1) & 31;
    }
    double s = wa[warp][0][lane] + rx;
#pragma unroll
    for (int t = 0; t < TI; ++t) s += ax[t] + ay[t] + az[t];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <class F>
double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, blocks = sms * 3, rot = 20000;
    const double clk = prop.clockRate * 1e3;  // Hz (max boost)
    double4* pm;
    double* out;
    CK(cudaMalloc(&pm, sizeof(double4) * blocks * 128 * 4));
    CK(cudaMalloc(&out, sizeof(double) * blocks * 128));
    double4* h = (double4*)malloc(sizeof(double4) * blocks * 128 * 4);
    srand(1);
    for (int i = 0; i < blocks * 128 * 4; ++i) h[i] = make_double4(rand() / 1e9, rand() / 1e9, rand() / 1e9, 1.0 / 65536);
    CK(cudaMemcpy(pm, h, sizeof(double4) * blocks * 128 * 4, cudaMemcpyHostToDevice));
    auto report = [&](const char* name, double ms, double per_rot = 80.0) {
        // FP64 warp instructions per SMSP: rot * per_rot per warp, 4 warps per CTA, 3 CTAs per SM, 4 SMSPs per SM
        const double inst_per_smsp = (double)rot * per_rot * 4.0 * 3.0 / 4.0;
        const double cycles = ms * 1e-3 * clk;
        printf("{\"variant\": \"%s\", \"ms\": %.4f, \"fp64_inst_per_cycle_per_smsp\": %.4f, \"pipe_util\": %.4f}\n", name, ms,
               inst_per_smsp / cycles, 2.0 * inst_per_smsp / cycles);
    };
    report("V0 mufu+smem", time_ms([&] { k_mix<4, true, true><<<blocks, 128>>>(pm, out, rot); }));
    report("V1 int-seed+smem", time_ms([&] { k_mix<4, false, true><<<blocks, 128>>>(pm, out, rot); }));
    double* din;
    CK(cudaMalloc(&din, sizeof(double) * 128 * 24));
    {
        double hin[128 * 24];
        for (int i = 0; i < 128 * 24; ++i) hin[i] = 0.999 + 1e-6 * (i % 97);
        CK(cudaMemcpy(din, hin, sizeof(hin), cudaMemcpyHostToDevice));
    }
    report("O0 dfma 3 distinct regs", time_ms([&] { k_ops<0><<<blocks, 128>>>(din, out, rot); }));
    report("O1 dfma 2 distinct regs", time_ms([&] { k_ops<1><<<blocks, 128>>>(din, out, rot); }));
    report("O2 dadd", time_ms([&] { k_ops<2><<<blocks, 128>>>(din, out, rot); }));
    report("O3 dmul", time_ms([&] { k_ops<3><<<blocks, 128>>>(din, out, rot); }));
    report("O4 dfma 3 distinct rotating", time_ms([&] { k_ops<4><<<blocks, 128>>>(din, out, rot); }));
    report("O5 dfma shared scalar (reuse)", time_ms([&] { k_ops<5><<<blocks, 128>>>(din, out, rot); }));
    report("O6 dfma scalar shared by 4 (reuse)", time_ms([&] { k_ops<6><<<blocks, 128>>>(din, out, rot); }));
    report("B0 body as kernel", time_ms([&] { k_body<0><<<blocks, 128>>>(pm, out, rot); }));
    report("B1 body, no j-side accumulate (62/rot)", time_ms([&] { k_body<1><<<blocks, 128>>>(pm, out, rot); }), 62.0);
    report("B2 body, no syncwarp", time_ms([&] { k_body<2><<<blocks, 128>>>(pm, out, rot); }));
    report("B3 body, j data by shuffle", time_ms([&] { k_body<3><<<blocks, 128>>>(pm, out, rot); }));
    report("B4 body, unroll 4", time_ms([&] { k_body<4><<<blocks, 128>>>(pm, out, rot); }));
    report("B5 body, accumulate triples as asm blocks", time_ms([&] { k_body<5><<<blocks, 128>>>(pm, out, rot); }));
    report("V4 dfma chains", time_ms([&] { k_dfma<<<blocks, 128>>>(out, rot, 1.0); }));
    printf("{\"clock_hz_assumed\": %.0f, \"sms\": %d}\n", clk, sms);
    return 0;
}
