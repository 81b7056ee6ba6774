"""ctypes binding of libee_b200.so (the C ABI in include/ee_b200.h).

There is no CPU fallback: if the shared library cannot be loaded (or built) importing this module raises, and every
entry point fails with EE_ERR_CUDA when no CUDA device is present.
"""
import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libee_b200.so"

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)


class AdaptiveParams(C.Structure):
    """ee_adaptive_params == AdaptiveMethodParams<f64, AbsTol, f64> (integration/src/lib.rs:174-197)."""

    _fields_ = [
        ("h_init", C.c_double), ("h_max", C.c_double),
        ("tol_position", C.c_double), ("tol_velocity", C.c_double),
        ("fac_min", C.c_double), ("fac_max", C.c_double), ("fac", C.c_double),
        ("n_max", C.c_uint32), ("pow_mode", C.c_uint32), ("method", C.c_uint32),
    ]


# every symbol include/ee_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ee_last_error": (C.c_char_p, []),
    "ee_version": (C.c_int32, []),
    "ee_launch_count": (C.c_uint64, []),
    "ee_set_pair_variant": (C.c_int32, [C.c_int32]),
    "ee_host_sampling_stride": (C.c_int64, [C.c_double, C.c_double]),
    "ee_host_plummer": (C.c_int32, [C.c_int64, C.c_uint64, c_double_p, c_double_p, c_double_p]),
    "ee_host_pair_schedule": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i64_p, c_i64_p,
                                          c_i64_p, c_i64_p, c_i32_p, C.c_int64, c_i32_p]),
    "ee_nbody_create": (C.c_int32, [C.c_int64, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int32,
                                    C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ee_nccl_unique_id": (C.c_int32, [C.c_void_p]),
    "ee_nbody_create_sharded": (C.c_int32, [C.c_int64, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double,
                                            C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_int32, C.POINTER(C.c_void_p)]),
    "ee_nbody_p2p_export": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "ee_nbody_p2p_connect": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "ee_nbody_p2p_trace": (C.c_int32, [C.c_void_p, C.c_int32, c_double_p, c_i64_p]),
    "ee_nbody_set_solout": (C.c_int32, [C.c_void_p, C.c_double, c_double_p, c_i32_p]),
    "ee_nbody_step": (C.c_int32, [C.c_void_p, C.c_int64]),
    "ee_nbody_step_to": (C.c_int32, [C.c_void_p, C.c_double]),
    "ee_nbody_sync": (C.c_int32, [C.c_void_p]),
    "ee_nbody_state": (C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "ee_nbody_state_async": (C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "ee_nbody_state_wait": (C.c_int32, [C.c_void_p]),
    "ee_nbody_delta": (C.c_double, [C.c_void_p]),
    "ee_nbody_step_count": (C.c_int64, [C.c_void_p]),
    "ee_nbody_solution_time": (C.c_int32, [C.c_void_p, c_double_p]),
    "ee_nbody_has_reached": (C.c_int32, [C.c_void_p, C.c_double, c_i32_p]),
    "ee_nbody_solution_sizes": (C.c_int32, [C.c_void_p, c_i64_p]),
    "ee_nbody_take_solution": (C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_i32_p]),
    "ee_nbody_take_solution_ephem": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ee_nbody_clone": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ee_nbody_snapshot_size": (C.c_int32, [C.c_void_p, c_i64_p]),
    "ee_nbody_snapshot": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "ee_nbody_restore": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64]),
    "ee_nbody_step_timed": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, c_double_p]),
    "ee_fp64_fma_peak": (C.c_int32, [C.c_int32, c_double_p]),
    "ee_nbody_destroy": (None, [C.c_void_p]),
    "ee_gravity_eval": (C.c_int32, [C.c_int64, c_double_p, c_double_p, C.c_int32, C.c_int32, c_double_p]),
    "ee_nbody_last_timing": (C.c_int32, [C.c_void_p, c_double_p, c_i64_p]),
    "ee_ephem_create": (C.c_int32, [C.c_int64, c_double_p, c_double_p, c_double_p, c_i64_p, c_double_p, c_i32_p,
                                    C.c_int32, C.POINTER(C.c_void_p)]),
    "ee_ephem_evaluate": (C.c_int32, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_i32_p]),
    "ee_ephem_sizes": (C.c_int32, [C.c_void_p, c_i64_p, c_i64_p]),
    "ee_ephem_get": (C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_i32_p]),
    "ee_ephem_destroy": (None, [C.c_void_p]),
    "ee_lsq_fit": (C.c_int32, [C.c_int64, c_i32_p, c_double_p, c_double_p, C.c_int32, c_double_p, c_i32_p]),
    "ee_ships_create": (C.c_int32, [C.c_void_p, C.c_int64, c_double_p, c_double_p, C.POINTER(AdaptiveParams), c_i64_p,
                                    c_double_p, c_double_p, c_double_p, c_i32_p, C.POINTER(C.c_void_p)]),
    "ee_ships_step_to": (C.c_int32, [C.c_void_p, C.c_double, C.c_int64]),
    "ee_ships_info": (C.c_int32, [C.c_void_p, c_i32_p, c_double_p, c_i64_p, c_u32_p, c_u64_p]),
    "ee_ships_take_knots": (C.c_int32, [C.c_void_p, c_i64_p, c_double_p]),
    "ee_ships_enable_analytics": (C.c_int32, [C.c_void_p, c_double_p]),
    "ee_ships_analytics_counts": (C.c_int32, [C.c_void_p, c_i32_p, c_i32_p]),
    "ee_ships_read_analytics": (C.c_int32, [C.c_void_p, c_i64_p, c_double_p, c_i32_p, c_i64_p, c_double_p, c_double_p, c_i32_p,
                                            c_i32_p]),
    "ee_ephem_evaluate_relative": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, c_double_p, c_double_p, c_double_p, c_i32_p]),
    "ee_ships_evaluate_relative": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, c_double_p, c_double_p, c_double_p, c_i32_p]),
    "ee_ships_last_ms": (C.c_double, [C.c_void_p]),
    "ee_ships_destroy": (None, [C.c_void_p]),
}

STATUS_NAMES = {
    0: "ok", 1: "step size underflow", 2: "max iterations reached", 3: "integration bound reached",
    4: "failed to evaluate ODE", 5: "solout exit", 100: "invalid argument", 101: "CUDA error", 102: "NCCL error",
    103: "unsupported",
}


def load() -> C.CDLL:
    if not LIB_PATH.exists():
        if os.environ.get("EE_NO_AUTOBUILD"):
            raise ImportError("libee_b200.so is missing (run __graft_entry__.build())")
        from .build import build
        build()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class EngineError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib.ee_last_error().decode() if code >= 100 else STATUS_NAMES.get(code, "?")
        super().__init__("%s: status %d (%s)" % (where, code, msg))


def check(code: int, where: str) -> None:
    if code != 0:
        raise EngineError(code, where)
