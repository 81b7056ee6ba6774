#!/usr/bin/env python3
"""bench.py -- body-steps/s of the n-body hot path on B200, next to the reference's CPU path.

Workload (BASELINE.json configs[3], the configuration the metric and its 70 %-of-fp64-roofline target are quoted on):
a synthetic 65 536-body Plummer sphere, f64, QuinlanTremaine12 (Blanes-Moan start-up excluded from timing), h = 2^-10,
throughput mode.  A "step" is one integrator step = one all-pairs acceleration evaluation fused with the Cowell
velocity reconstruction and the predictor, i.e. 65 536 body-steps.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
  python bench.py --impl reference [--gpus N] --steps K ...       the reference's CPU algorithm (the C++ oracle port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import os as _os
# NCCL writes its INFO log to stdout by default, where the one JSON line goes: send it to stderr instead (the driver
# reads the rank count from it), unless the caller already chose a file.
if _os.environ.get("NCCL_DEBUG") and not _os.environ.get("NCCL_DEBUG_FILE"):
    _os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_BODIES = 65536
H_STEP = 2.0 ** -10
FLUSH_BYTES = 256 << 20  # > 126 MB L2
REF_BUDGET_S = 240.0  # CPU seconds the --impl reference arm may spend on warm-up + timed steps ("within a few minutes")
WORKLOAD = "plummer-65536 QuinlanTremaine12 steady-state step, h=2^-10 (BASELINE.json configs[3])"


def flops_per_body_step(n):
    # 20 flop per directed interaction + 236 flop of integrator linear algebra
    return 20.0 * (n - 1) + 236.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the C++ oracle (the reference's algorithm restated; the Rust reference cannot be built in this image).
# This is the ONLY part of bench.py that touches oracle/.
_ORACLE = None


def oracle_lib():
    """The oracle as a timing subject: a copy built ON this box with -O3 -march=native (`make -C oracle native`), falling
    back to the portable build the tests use.  Same source and arithmetic either way (no FMA contraction)."""
    global _ORACLE
    if _ORACLE is not None:
        return _ORACLE
    build = "portable (-O3, generic x86-64)"
    path = ROOT / "oracle" / "libee_oracle.so"
    try:
        r = subprocess.run(["make", "-C", str(ROOT / "oracle"), "native"], capture_output=True, text=True, timeout=300)
        nat = ROOT / "oracle" / "_native" / "libee_oracle_native.so"
        if r.returncode == 0 and nat.exists():
            path, build = nat, "-O3 -march=native, built on this box"
    except Exception:
        pass
    if not path.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "libee_oracle.so"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(str(path))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ora_gravity_eval_row_sample.restype = ctypes.c_int64
    lib.ora_gravity_eval_row_sample.argtypes = [ctypes.c_int64, dp, dp, ctypes.c_int64, ctypes.c_int64, dp]
    lib.ora_gravity_eval_row_sample_mt.restype = ctypes.c_int64
    lib.ora_gravity_eval_row_sample_mt.argtypes = [ctypes.c_int64, dp, dp, ctypes.c_int64, ctypes.c_int32, dp,
                                                   ctypes.POINTER(ctypes.c_int32)]
    lib.ora_nbody_create.restype = ctypes.c_void_p
    lib.ora_nbody_create.argtypes = [ctypes.c_int64, dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_int32]
    lib.ora_nbody_set_solout.argtypes = [ctypes.c_void_p, ctypes.c_double, dp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]
    lib.ora_nbody_destroy.argtypes = [ctypes.c_void_p]
    lib.ora_planner_loop.restype = ctypes.c_int32
    lib.ora_planner_loop.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_int64),
                                     ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), dp,
                                     ctypes.POINTER(ctypes.c_uint64)]
    _ORACLE = (lib, build)
    return _ORACLE


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def oracle_row_sample(pos, mu, rows_stride, row0=0):
    """Times the reference's pair loop on rows row0, row0+stride, ... of one evaluation; returns (seconds, pairs)."""
    lib, _ = oracle_lib()
    n = len(mu)
    out = np.zeros((n, 3))
    t0 = time.perf_counter()
    pairs = lib.ora_gravity_eval_row_sample(n, _dp(pos), _dp(mu), row0, rows_stride, _dp(out))
    return time.perf_counter() - t0, pairs


def oracle_all_cores(pos, mu, seconds=4.0):
    """Context figure only: the same symmetric pair loop on every host thread (native threads inside the oracle, private
    output arrays, reduction over threads included).  The reference itself runs this loop on ONE thread (nbody.rs:22-38
    inside one background task, prediction.rs:385), so the headline CPU arm stays single-threaded."""
    lib, _ = oracle_lib()
    n = len(mu)
    threads = os.cpu_count() or 1
    total_pairs = n * (n - 1) // 2
    t_probe, p_probe = oracle_row_sample(pos, mu, rows_stride=4096)
    stride = max(1, int(round(total_pairs / ((p_probe / t_probe) * seconds * threads))))  # ~`seconds` of wall time
    secs, used = ctypes.c_double(), ctypes.c_int32()
    pairs = lib.ora_gravity_eval_row_sample_mt(n, _dp(pos), _dp(mu), stride, threads, ctypes.byref(secs), ctypes.byref(used))
    return {"threads": used.value, "body_steps_per_s": n / (secs.value * total_pairs / pairs), "pairs_per_s": pairs / secs.value,
            "note": "symmetric pair loop, rows dealt round-robin to native threads, private outputs + reduction, extrapolated "
                    "by pair count from %.2f%% of the pairs" % (100.0 * pairs / total_pairs)}


def cpu_baseline(pos, mu, target_seconds=12.0):
    """The reference's algorithm (C++ oracle, 1 thread -- the reference's pair loop is serial, nbody.rs:22-38) on a
    bounded sample of the same workload: a strided subset of the rows of ONE acceleration evaluation, extrapolated by
    the exact pair count.  The O(24 N) multistep update (< 0.01 % of a step at this N) is not included."""
    _, build = oracle_lib()
    n = len(mu)
    total_pairs = n * (n - 1) // 2
    t_probe, p_probe = oracle_row_sample(pos, mu, rows_stride=4096)  # ~16 rows
    rate = p_probe / t_probe
    stride = max(1, int(round(total_pairs / (rate * target_seconds))))
    t, pairs = oracle_row_sample(pos, mu, rows_stride=stride)
    t_step = t * total_pairs / pairs
    return {"value": n / t_step, "unit": "body-steps/s", "cores": 1, "kind": "port", "build": build,
            "sample": "rows 0,%d,2*%d.. of the symmetric pair loop of one 65536-body evaluation (%.1f%% of its pairs, %.1f s), "
                      "extrapolated by pair count; multistep update (<0.01%%) excluded" % (stride, stride, 100.0 * pairs / total_pairs, t),
            "pairs_per_s": pairs / t, "host_cores_available": os.cpu_count(), "all_cores_context": oracle_all_cores(pos, mu)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import ephemeris_explorer_b200.synthetic as synthetic  # pure numpy
    _, build = oracle_lib()
    pos, vel, mu = synthetic.plummer(N_BODIES)
    n = N_BODIES
    total_pairs = n * (n - 1) // 2
    # Size the per-step sample so that warm-up + timed steps end within REF_BUDGET_S of CPU time.  Whenever the whole run
    # fits (the driver's --steps 20 --warmup 5 does: 25 x ~5 s) every step is ONE COMPLETE 65536-body evaluation of the
    # reference's pair loop and nothing is extrapolated; otherwise a step is every stride-th row, scaled by the pair count.
    # The first warm-up step is always one complete evaluation: it is also the measurement the sample size is chosen from
    # (a short probe over-estimates the full time by up to 2x).
    t_full, pairs = oracle_row_sample(pos, mu, rows_stride=1)
    assert pairs == total_pairs
    stride = max(1, int(np.ceil(t_full * (args.warmup + args.steps) / REF_BUDGET_S)))
    times, sampled = [], []
    for it in range(1, args.warmup + args.steps):
        t, pairs = oracle_row_sample(pos, mu, rows_stride=stride, row0=it % stride)
        if it >= args.warmup:
            sampled.append(t)
            times.append(t * total_pairs / pairs)
    t_step = float(np.mean(times))
    value = n / t_step
    if stride == 1:
        sample = ("each step = one complete 65536-body evaluation of the symmetric pair loop (100% of the pairs, nothing "
                  "extrapolated); single thread (the reference's loop is serial); the O(24 N) multistep update (< 0.01 %) excluded")
    else:
        sample = ("each step = rows r, r+%d, .. of one 65536-body evaluation's pair loop (%.2f%% of the pairs), extrapolated by "
                  "pair count; single thread (the reference's loop is serial)" % (stride, 100.0 * pairs / total_pairs))
    line = {
        "impl": "reference", "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "bodies": n, "parallelism": "cpu-1thread"},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": "port", "build": build, "sample": sample,
                         "extrapolated": stride != 1, "ms_per_sampled_step": float(np.mean(sampled)) * 1e3,
                         "all_cores_context": oracle_all_cores(pos, mu)},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def rel_err(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))


def run_ours(args, rank, world, local_rank):
    import torch
    import ephemeris_explorer_b200 as ee
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    from ephemeris_explorer_b200 import distributed as eed

    def max_over_ranks(x):
        return eed.max_over_ranks(dist, x, device="cuda")

    n = N_BODIES
    pos, vel, mu = ee.synthetic.plummer(n)
    exchange = ee.EXCHANGE_ALLGATHER if args.exchange == "allgather" else ee.EXCHANGE_ALLREDUCE  # p2p uses the allreduce layout
    uid = None
    if world > 1:
        uid = eed.broadcast_unique_id(dist, ee.nccl_unique_id() if rank == 0 else None, device="cuda")
    prop = ee.NBodyPropagator.new(ee.Forward(H_STEP), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT, device=local_rank,
                                  rank=rank, world=world, unique_id=uid, exchange=exchange)
    if world > 1 and args.exchange == "p2p":
        eed.connect_peers(dist, prop, device="cuda")  # NVLink peer path: CUDA-IPC handles travel over torch.distributed
    prop.step(12)  # Blanes-Moan start-up (12 calls, 289 evaluations): not part of the steady-state metric
    sampler = ClockSampler(local_rank)
    sampler.start()  # samples span warm-up, the timed steps and the e2e loop (the timed region alone is only ~0.1 s)
    prop.step_timed(max(args.warmup, 3), FLUSH_BYTES)
    prop.sync()
    barrier()
    launches0 = ee.lib.ee_launch_count()
    ms = prop.step_timed(args.steps, FLUSH_BYTES)  # sum of K per-step CUDA-event intervals on the launching stream
    prop.sync()
    launches = ee.lib.ee_launch_count() - launches0
    barrier()
    ms = max_over_ranks(ms)
    value = eed.whole_job_rate(n, world, args.steps, ms * 1e-3, sharded=True)
    ms_per_step = ms / args.steps

    # ---- end to end through the public API with HOST buffers, same at every N:
    #   H2D  ee_nbody_restore of the complete multistep state from pinned host memory on every rank (the reference's `extend`
    #        resumes from a host-side clone, prediction.rs:378) -- once, amortised over the K steps;
    #   per step  ee_nbody_step(1), then ee_nbody_state_async -> positions + velocities into pinned host memory on rank 0
    #        (the copy of step s overlaps the kernels of step s+1; two host buffers alternate).
    e2e = None
    if args.exchange != "allgather":
        blob = np.empty(prop.snapshot_size(), dtype=np.uint8)
        prop.snapshot(blob)
        pin = torch.empty(blob.nbytes, dtype=torch.uint8).pin_memory()
        hb = pin.numpy()
        hb[:] = blob
        host = [(torch.empty((n, 3), dtype=torch.float64).pin_memory(), torch.empty((n, 3), dtype=torch.float64).pin_memory())
                for _ in range(2)]
        bufs = [(a.numpy(), b.numpy()) for a, b in host]
        barrier()
        t0 = time.perf_counter()
        prop.restore(hb)
        if dist:
            dist.barrier()  # a rank may not store into a peer that is still restoring
        for s in range(args.steps):
            prop.step(1)
            if rank == 0:
                prop.state_async(*bufs[s & 1])
        if rank == 0:
            prop.state_wait()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n * args.steps / dt, "unit": "body-steps/s", "h2d_bytes_per_step": world * blob.nbytes / args.steps,
               "d2h_bytes_per_step": 2 * n * 24,
               "how": "restore(pinned host snapshot) on every rank once, then per step: step(1) + state_async()->pinned host "
                      "positions+velocities on rank 0 (overlaps the next step), state_wait() at the end; wall clock, max over ranks"}

    # ---- self-check of what was just timed: the state after all those steps against an independent run
    parity = None
    total_steps = prop.step_count()
    t_end, p_end, v_end = prop.state()
    if rank == 0:
        if world > 1:
            ref = ee.NBodyPropagator.new(ee.Forward(H_STEP), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT, device=local_rank)
            ref.step(total_steps)
            rt, rp, rv = ref.state()
            ref.close()
            parity = {"parity_rel": rel_err(p_end, rp), "parity_rel_velocity": rel_err(v_end, rv), "steps": int(total_steps),
                      "time_equal": bool(rt == t_end),
                      "against": "1-GPU throughput handle stepped the same %d steps on rank 0's GPU" % total_steps}
        if world == 1 and not args.no_parity_kernel:
            k = 20  # 12 start-up + 8 steady steps: the parity-mode kernel is bit-exact to the CPU oracle at any N
            a = ee.NBodyPropagator.new(ee.Forward(H_STEP), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT, device=local_rank)
            b = ee.NBodyPropagator.new(ee.Forward(H_STEP), 0.0, pos, vel, mu, mode=ee.MODE_PARITY, device=local_rank)
            a.step(k)
            b.step(k)
            parity = {"parity_rel": rel_err(a.state()[1], b.state()[1]), "steps": k,
                      "against": "parity-mode kernel (reference summation order, IEEE sqrt/div; bit-exact to the CPU oracle in "
                                 "tests/test_nbody_gpu.py), same %d steps" % k}
            a.close()
            b.close()
    breakdown = None
    if world > 1 and args.exchange == "p2p":  # where a sharded step spends its time (event-timed, synchronising: not the timed path)
        barrier()  # rank 0 has just run the self-check: without this the first traced barrier would measure that wait
        prop.p2p_trace(True)
        prop.step(2)   # discarded: the ranks fall into step with each other
        prop.sync()
        prop.p2p_trace(False)
        prop.p2p_trace(True)
        prop.step(8)
        prop.sync()
        ms4, ksteps = prop.p2p_trace(False)
        mx = [max_over_ranks(float(v)) for v in ms4]
        breakdown = {"steps": int(ksteps), "max_over_ranks_ms": {"pair_units_and_local_reduce": mx[0], "barrier_1": mx[1],
                                                                  "slice_finish_peer_loads_stores": mx[2], "barrier_2": mx[3]},
                     "rank0_ms": [float(v) for v in ms4],
                     "note": "CUDA events around the five launches of the peer step; barrier times include waiting for the slowest rank"}
    if dist:
        dist.barrier()

    clocks = sampler.stop()
    if rank == 0:
        peak = ee.fp64_fma_peak(local_rank)  # measured in this run with independent DFMA chains
        nominal = 148 * 64 * 2 * 1.965e9 / 1e12
        achieved = value * flops_per_body_step(n) / world / 1e12  # per GPU
        # DRAM bytes per launch from the round-2 `ncu --set full` capture (profiles/r02): see profiles/README.md
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": TRAFFIC_BYTES if world == 1 else None,
                    "traffic_step_total": TRAFFIC_BYTES + TRAFFIC_REDUCE_BYTES if world == 1 else None,
                    "traffic_source": "ncu --set full, round 2 (profiles/r02/ncu_sym_raw.csv): dram read+write per launch of the dominant "
                                      "kernel k_accel_sym; traffic_step_total adds k_sym_reduce, which reads the partial sums back "
                                      "(0.1 ms of a 3.0 ms step).  Algorithmic state traffic is 530 B x 65536 = 34.7 MB: the 10x excess is "
                                      "the price of evaluating every pair once (partials for both bodies), and the path stays FP64-pipe-bound",
                    "kernel": "k_accel_sym (98% of the step) + k_sym_reduce (2%); achieved uses the whole step time", "per": "GPU",
                    "peak_source": "measured in-run: 8 independent DFMA chains/thread, 2 flop/FMA (MEASURED_PEAKS.json has no fp64 figure)",
                    "nominal_peak": nominal, "frac_of_nominal": achieved / nominal, "flops_per_body_step": flops_per_body_step(n)}
        line = {
            "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bodies": n, "mode": "throughput",
                       "parallelism": "1gpu" if world == 1 else "%s-x%d" % (args.exchange, world),
                       "exchange": None if world == 1 else {"p2p": "pair units sharded; local reduce; NVLink peer loads of G partial accelerations + epilogue + peer stores in one kernel, flag barriers (no NCCL on the data path)", "allreduce": "pair units sharded; ncclAllReduce of 3N partial accelerations per evaluation", "allgather": "targets sharded; ncclAllGather of new positions"}[args.exchange],
                       "l2": "256 MiB flush written before every timed step (outside the event pair)",
                       "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if breakdown:
            line["p2p_breakdown"] = breakdown
        if parity:
            line["parity_rel"] = parity["parity_rel"]
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pos, mu)
        if world == 1 and not args.no_extras:
            line["extras"] = extras(ee, local_rank, peak)
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


# ncu --set full, round 2 session B (profiles/r02/ncu_sym_raw.csv), DRAM bytes per launch:
TRAFFIC_BYTES = 124.9e6        # k_accel_sym<4,256,2,16>: 2.2 MB read + 122.7 MB written (the i-/j-side partial sums)
TRAFFIC_REDUCE_BYTES = 221.0e6  # k_sym_reduce<1024,8,32>: 214.4 MB read (the partials + the multistep history) + 6.5 MB written


# ---------------------------------------------------------------------------------------------------------------------
def planner_loop_c2(ee, s, device, days=3650.0):
    """C2 through the drop-in call shape: the Prediction Planner's loop (step() one at a time, has_reached() every step,
    take_solution() + clone() at 100 Hz) -- ours from C++ host code over the C ABI (tools/planner_loop.cpp), the CPU arm
    by the oracle running the same loop.  Hashes of all fitted polynomials must agree (bit-exact splines)."""
    end = s.epoch + days * 86400.0
    tick = 0.01
    out = {}
    exe = ROOT / "tools" / "bin" / "planner_loop"
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(np.array([len(s.mu)], dtype=np.int64).tobytes())
        f.write(np.array([s.epoch, s.dt, end, tick], dtype=np.float64).tobytes())
        for a in (s.position, s.velocity, s.mu, s.sample_period):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(s.degree, dtype=np.int32).tobytes())
        path = f.name
    try:
        r = subprocess.run([str(exe), path, str(device)], capture_output=True, text=True, timeout=600)
        ours = json.loads(r.stdout.strip().splitlines()[-1])
    finally:
        os.unlink(path)
    out["gpu"] = ours
    lib, build = oracle_lib()
    pos, vel, mu = (np.ascontiguousarray(a, dtype=np.float64) for a in (s.position, s.velocity, s.mu))
    per = np.ascontiguousarray(s.sample_period, dtype=np.float64)
    deg = np.ascontiguousarray(s.degree, dtype=np.int32)
    h = lib.ora_nbody_create(len(mu), _dp(pos), _dp(vel), _dp(mu), s.epoch, s.dt, 12)
    lib.ora_nbody_set_solout(h, s.dt, _dp(per), deg.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 0)
    steps, ticks, polys, secs, hsh = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_uint64()
    st = lib.ora_planner_loop(h, end, tick, ctypes.byref(steps), ctypes.byref(ticks), ctypes.byref(polys), ctypes.byref(secs),
                              ctypes.byref(hsh))
    lib.ora_nbody_destroy(h)
    out["cpu_oracle"] = {"status": st, "steps": steps.value, "seconds": secs.value, "steps_per_s": steps.value / secs.value,
                         "ticks": ticks.value, "polynomials": polys.value, "hash": "%016x" % hsh.value, "cores": 1, "build": build}
    if "steps_per_s" in ours:
        out["speedup_vs_cpu_oracle"] = ours["steps_per_s"] / out["cpu_oracle"]["steps_per_s"]
        out["splines_bit_exact"] = bool(ours["hash"] == out["cpu_oracle"]["hash"] and ours["steps"] == steps.value)
    out["loop"] = ("step() x1, has_reached() every step, take_solution()+clone() every %.0f ms wall clock (Synchronisation::hertz(100)), "
                   "until the splines reach epoch + %.0f days; 12 start-up steps outside the clock" % (tick * 1e3, days))
    return out


def extras(ee, device, fp64_peak):
    """The other BASELINE.json configs, measured briefly (not the headline)."""
    out = {}
    try:
        s = ee.formats.load_system(ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json")
        # C2, one batched call (the kernel's own rate)
        prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY, device=device,
                                      solout=(s.dt, s.sample_period, s.degree))
        prop.step(12)
        prop.sync()
        t0 = time.perf_counter()
        k = 200000
        prop.step(k)
        prop.sync()
        dt = time.perf_counter() - t0
        out["C2_full_solar_system_32_parity"] = {"steps_per_s": k / dt, "body_steps_per_s": 32 * k / dt, "steps": k, "bound": "latency",
                                                 "note": "bit-exact parity mode, persistent single-CTA kernel, spline solout on, one step(200000) call"}
        prop.close()
        out["C2_planner_loop"] = planner_loop_c2(ee, s, device)
        # C3
        n3 = 4096
        p0, v0, mu = ee.synthetic.plummer(n3)
        prop = ee.NBodyPropagator.new(ee.Forward(H_STEP), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT, device=device)
        prop.step(12 + 3)
        prop.sync()
        ms = 1e30
        for _ in range(3):  # 64 steps enqueued back to back inside one event pair (what a step_to / batched step call does)
            prop.step(64)
            prop.sync()
            ms = min(ms, prop.last_timing()[0])
        ms_sync = prop.step_timed(64, 0)  # the same with a host synchronisation after every step
        bs = n3 * 64 / (ms * 1e-3)
        ach = bs * flops_per_body_step(n3) / 1e12
        out["C3_plummer_4096"] = {"body_steps_per_s": bs, "ms_per_step": ms / 64, "ms_per_step_host_sync_each": ms_sync / 64,
                                  "kernels": "pair-symmetric kernel, 4-warp CTAs x 512-body tiles + warp-per-body reduce, programmatic "
                                             "dependent launch (2 launches per step)",
                                  "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                                               "l2": "inputs L2-resident (1.6 MB of state), no flush",
                                               "note": "8.4e6 pairs are 9 us of FP64 pipe; every warp's share is one 128x32 block of pairs, so "
                                                       "the step is launch ramp + one block + reduce (profiles/README.md)"}}
        prop.close()
        out["C5_ships"] = ships_c5(ee, s, device)
    except Exception as exc:  # extras never break the headline line
        out["error"] = repr(exc)
    return out


def ships_c5(ee, s, device, ns=1024):
    """C5: 1 024 perturbed "Mars Transfer Ship" states coasting 1950-01-01 -> 1950-08-20 against the 2-year spline ephemeris
    of the 32-body system (built on the device).  Roofline: L2 (spline table look-ups), 6 360 algorithmic bytes per RHS."""
    eph_prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY, device=device,
                                      solout=(s.dt, s.sample_period, s.degree))
    eph_prop.step_to(s.epoch + 2 * 365 * 86400.0)
    eph = eph_prop.take_solution_ephemeris()
    ship = ee.formats.load_ship(ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json", s.names,
                                name="Mars Transfer Ship")
    rng = np.random.default_rng(20260924)
    states = np.tile(np.concatenate([ship.position, ship.velocity]), (ns, 1))
    states[:, :3] += 10.0 * rng.uniform(-1, 1, (ns, 3))
    states[:, 3:] += 0.010 * rng.uniform(-1, 1, (ns, 3))
    ships = ee.SpacecraftPropagator.new(ship.start, states, ee.default_adaptive_params(ship.tolerance, ship.tolerance), None, eph)
    t0 = time.perf_counter()
    ships.step_to(ship.end, max_steps=200000)
    wall = time.perf_counter() - t0
    info = ships.info()
    # the same batch with the app's SpacecraftSolout (SOI transitions + apsides searched on every accepted step)
    ana = ee.SpacecraftPropagator.new(ship.start, states, ee.default_adaptive_params(ship.tolerance, ship.tolerance), None, eph)
    ana.enable_analytics(ee.formats.soi_radii(s))
    ana_ms = 0.0
    while True:
        ana.step_to(ship.end, max_steps=200000)
        ana_ms += ana.last_ms()
        ai = ana.info()
        if np.all((ai["time"] >= ship.end) | (ai["status"] != 0)):
            break
    events = ana.analytics()
    ana_out = {"kernel_ms": ana_ms, "transitions": int(sum(len(t) for t, _ in events)), "apsides": int(sum(len(a) for _, a in events)),
               "knots_equal_to_plain_run": bool(np.array_equal(ai["n_knots"], info["n_knots"]))}
    ana.close()
    steps = int(info["n_knots"].sum() - ns)
    evals = int(info["rhs_evals"].sum())
    kms = ships.last_ms()
    bytes_per_rhs = 6360.0
    gbs = evals * bytes_per_rhs / (kms * 1e-3) / 1e9
    return {"ships": ns, "status_ok": int((info["status"] == 0).sum()), "accepted_steps": steps, "rhs_evals": evals,
            "kernel_ms": kms, "wall_s": wall, "ship_steps_per_s": steps / (kms * 1e-3), "rhs_evals_per_s": evals / (kms * 1e-3),
            "method": "Verner87", "with_analytics": ana_out,
            "roofline": {"bound": "latency", "achieved": gbs, "unit": "GB/s", "peak": None,
                         "note": "achieved = algorithmic 6360 B of spline coefficients per RHS (SURVEY 8d) x RHS evaluations / kernel "
                                 "time, the figure SURVEY 8d asks for; it is not what bounds the kernel: 1024 ships are 1.7 warps per "
                                 "scheduler, the coefficients come out of a per-warp shared-memory cache, and the time is the FP64 "
                                 "dependency chain of one right-hand side (ncu: 50% fixed-latency waits, FP64 pipe 14% busy; "
                                 "profiles/README.md), so no fraction of a bandwidth peak is claimed"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "allreduce", "allgather"],
                    help="multi-GPU exchange: NVLink peer path (default), NCCL all-reduce of partial accelerations, NCCL all-gather of positions")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-parity-kernel", action="store_true", help="skip the 1-GPU self-check against the parity-mode kernel (~15 s)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
