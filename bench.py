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
if "EE_NCCL_DEBUG" in _os.environ:
    _os.environ["NCCL_DEBUG"] = _os.environ["EE_NCCL_DEBUG"]
else:
    _os.environ.pop("NCCL_DEBUG", None)  # NCCL prints its version banner to stdout at VERSION/WARN/INFO: keep stdout to one JSON line
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_BODIES = 65536
H_STEP = 2.0 ** -10
FLUSH_BYTES = 256 << 20  # > 126 MB L2


def flops_per_body_step(n):
    # 20 flop per directed interaction + 236 flop of integrator linear algebra
    return 20.0 * (n - 1) + 236.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

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


def oracle_row_sample(pos, mu, rows_stride, row0=0):
    """Times the reference's pair loop on rows row0, row0+stride, ... of one evaluation; returns (seconds, pairs)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    lib = oracle.lib
    lib.ora_gravity_eval_row_sample.restype = ctypes.c_int64
    lib.ora_gravity_eval_row_sample.argtypes = [ctypes.c_int64, oracle._dp, oracle._dp, ctypes.c_int64, ctypes.c_int64, oracle._dp]
    n = len(mu)
    out = np.zeros((n, 3))
    t0 = time.perf_counter()
    pairs = lib.ora_gravity_eval_row_sample(n, oracle.p(pos), oracle.p(mu), row0, rows_stride, oracle.p(out))
    return time.perf_counter() - t0, pairs


def oracle_all_cores(pos, mu, seconds=4.0):
    """Context figure only: the same symmetric pair loop spread over every host thread (thread t takes rows t, t+T*s, ..
    into its own output array; ctypes releases the GIL).  The reference itself runs this loop on ONE thread
    (nbody.rs:22-38 inside one background task, prediction.rs:385), so the headline CPU arm stays single-threaded."""
    import concurrent.futures as cf
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    lib = oracle.lib
    lib.ora_gravity_eval_row_sample.restype = ctypes.c_int64
    lib.ora_gravity_eval_row_sample.argtypes = [ctypes.c_int64, oracle._dp, oracle._dp, ctypes.c_int64, ctypes.c_int64, oracle._dp]
    n = len(mu)
    threads = os.cpu_count() or 1
    total_pairs = n * (n - 1) // 2
    t_probe, p_probe = oracle_row_sample(pos, mu, rows_stride=4096)
    stride = max(1, int(round(total_pairs / ((p_probe / t_probe) * seconds))))  # per-thread sample of ~`seconds`
    outs = [np.zeros((n, 3)) for _ in range(threads)]

    def work(t):
        return lib.ora_gravity_eval_row_sample(n, oracle.p(pos), oracle.p(mu), t, threads * stride, oracle.p(outs[t]))

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:
        pairs = sum(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return {"threads": threads, "body_steps_per_s": n / (dt * total_pairs / pairs), "pairs_per_s": pairs / dt,
            "note": "symmetric pair loop, rows dealt round-robin to threads, private outputs (no final reduction timed)"}


def cpu_baseline(pos, mu, target_seconds=12.0):
    """The reference's algorithm (C++ oracle, 1 thread -- the reference's pair loop is serial, nbody.rs:22-38) on a
    bounded sample of the same workload: a strided subset of the rows of ONE acceleration evaluation, extrapolated by
    the exact pair count.  The O(24 N) multistep update (< 0.01 % of a step at this N) is not included."""
    n = len(mu)
    total_pairs = n * (n - 1) // 2
    t_probe, p_probe = oracle_row_sample(pos, mu, rows_stride=4096)  # ~16 rows
    rate = p_probe / t_probe
    stride = max(1, int(round(total_pairs / (rate * target_seconds))))
    t, pairs = oracle_row_sample(pos, mu, rows_stride=stride)
    t_step = t * total_pairs / pairs
    return {"value": n / t_step, "unit": "body-steps/s", "cores": 1, "kind": "port",
            "sample": "rows 0,%d,2*%d.. of the symmetric pair loop of one 65536-body evaluation (%.1f%% of its pairs, %.1f s), "
                      "extrapolated by pair count; multistep update (<0.01%%) excluded" % (stride, stride, 100.0 * pairs / total_pairs, t),
            "pairs_per_s": pairs / t, "host_cores_available": os.cpu_count(), "all_cores_context": oracle_all_cores(pos, mu)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import ephemeris_explorer_b200.synthetic as synthetic  # pure numpy
    pos, vel, mu = synthetic.plummer(N_BODIES)
    n = N_BODIES
    total_pairs = n * (n - 1) // 2
    # size the per-step sample so that steps+warmup end within a few minutes (~0.4 s per step)
    t_probe, p_probe = oracle_row_sample(pos, mu, rows_stride=4096)
    stride = max(1, int(round(total_pairs / ((p_probe / t_probe) * 0.4))))
    times = []
    pairs = 0
    for it in range(args.warmup + args.steps):
        t, pairs = oracle_row_sample(pos, mu, rows_stride=stride, row0=it % stride)
        if it >= args.warmup:
            times.append(t * total_pairs / pairs)
    t_step = float(np.mean(times))
    value = n / t_step
    sample = ("each step = rows r, r+%d, .. of one 65536-body evaluation's pair loop (%.2f%% of the pairs), extrapolated by "
              "pair count; single thread (the reference's loop is serial)" % (stride, 100.0 * pairs / total_pairs))
    line = {
        "impl": "reference", "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "plummer-65536 QuinlanTremaine12 steady-state step, h=2^-10", "bodies": n,
                   "parallelism": "cpu-1thread"},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": "port", "sample": sample,
                         "all_cores_context": oracle_all_cores(pos, mu)},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    # ---- end to end through the public API with host buffers
    e2e = None
    if world == 1:
        blob = np.empty(prop.snapshot_size(), dtype=np.uint8)
        prop.snapshot(blob)
        pin = torch.empty(blob.nbytes, dtype=torch.uint8).pin_memory()
        hb = pin.numpy()
        hb[:] = blob
        t_host, p_host, v_host = ctypes.c_double(), np.zeros((n, 3)), np.zeros((n, 3))
        barrier()
        t0 = time.perf_counter()
        prop.restore(hb)  # H2D: the whole multistep state from pinned host memory (the reference's `extend` from a clone)
        for _ in range(args.steps):
            prop.step(1)
            ee._lib.check(ee.lib.ee_nbody_state(prop._h, ctypes.byref(t_host), p_host.ctypes.data_as(ee._lib.c_double_p),
                                               v_host.ctypes.data_as(ee._lib.c_double_p), None), "state")  # D2H of the step's result
        barrier()
        dt = time.perf_counter() - t0
        e2e = {"value": n * args.steps / dt, "unit": "body-steps/s", "h2d_bytes_per_step": blob.nbytes / args.steps,
               "d2h_bytes_per_step": 2 * n * 24,
               "how": "restore(host snapshot) once, then per step: step(1) + state()->host positions+velocities; wall clock"}
    else:
        p_host, v_host = np.zeros((n, 3)), np.zeros((n, 3))
        t_host = ctypes.c_double()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prop.step(1)
            ee._lib.check(ee.lib.ee_nbody_state(prop._h, ctypes.byref(t_host), p_host.ctypes.data_as(ee._lib.c_double_p),
                                               v_host.ctypes.data_as(ee._lib.c_double_p), None), "state")
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n * args.steps / dt, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 2 * n * 24,
               "how": "per step: step(1) + state()->host (gathers velocities over NCCL); wall clock, max over ranks"}

    clocks = sampler.stop()
    if rank == 0:
        peak = ee.fp64_fma_peak(local_rank)  # measured in this run with independent DFMA chains
        nominal = 148 * 64 * 2 * 1.965e9 / 1e12
        achieved = value * flops_per_body_step(n) / world / 1e12  # per GPU
        # DRAM bytes per launch of the dominant kernel from the round-1 `ncu --set full` capture (profiles/r01/
        # ncu_k_accel_sym_n65536_summary.csv): k_accel_sym 2.4 MB read + 94.4 MB written (the partial sums); the follow-up
        # k_sym_reduce reads 185 MB.  Algorithmic state traffic is 530 B x 65536 = 34.7 MB per step; none of it limits an
        # FP64-bound 3.2 ms kernel.
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": 96.8e6 if world == 1 else None, "traffic_source": "ncu r01, k_accel_sym<512>, bytes per launch",
                    "kernel": "k_accel_sym (98% of the step) + k_sym_reduce (2%); achieved uses the whole step time", "per": "GPU", "peak_source": "measured in-run: 8 independent DFMA chains/thread, 2 flop/FMA "
                    "(MEASURED_PEAKS.json has no fp64 figure)", "nominal_peak": nominal, "frac_of_nominal": achieved / nominal,
                    "flops_per_body_step": flops_per_body_step(n)}
        line = {
            "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "plummer-65536 QuinlanTremaine12 steady-state step, h=2^-10 (BASELINE.json configs[3])",
                       "bodies": n, "mode": "throughput", "parallelism": "1gpu" if world == 1 else "%s-x%d" % (args.exchange, world),
                       "exchange": None if world == 1 else {"p2p": "pair items sharded; local reduce; NVLink peer loads of G partial accelerations + epilogue + peer stores in one kernel, flag barriers (no NCCL on the data path)", "allreduce": "pair items sharded; ncclAllReduce of 3N partial accelerations per evaluation", "allgather": "targets sharded; ncclAllGather of new positions"}[args.exchange],
                       "l2": "256 MiB flush written before every timed step (outside the event pair)",
                       "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pos, mu)
        if world == 1 and not args.no_extras:
            line["extras"] = extras(ee, local_rank)
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def extras(ee, device):
    """The other BASELINE.json configs, measured briefly (not the headline)."""
    out = {}
    try:
        s = ee.formats.load_system(ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json")
        prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY, device=device,
                                      solout=(s.dt, s.sample_period, s.degree))
        prop.step(12)
        prop.sync()
        t0 = time.perf_counter()
        k = 200000
        prop.step(k)
        prop.sync()
        dt = time.perf_counter() - t0
        out["full_solar_system_32_parity"] = {"steps_per_s": k / dt, "body_steps_per_s": 32 * k / dt, "steps": k,
                                              "note": "bit-exact parity mode, persistent single-CTA kernel, spline solout on"}
        p0, v0, mu = ee.synthetic.plummer(4096)
        prop = ee.NBodyPropagator.new(ee.Forward(H_STEP), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT, device=device)
        prop.step(12 + 3)
        ms = prop.step_timed(64, 0)
        out["plummer_4096"] = {"body_steps_per_s": 4096 * 64 / (ms * 1e-3), "ms_per_step": ms / 64}
    except Exception as exc:  # extras never break the headline line
        out["error"] = repr(exc)
    return out


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
