"""ephemeris-explorer_b200: B200-native engine for ephemeris-explorer's gravitational-integration hot path.

The product is the CUDA library (csrc/ -> libee_b200.so, C ABI in include/ee_b200.h); this Python package is the thin
host-side mirror of the reference's propagator interface used by tests and bench.py.
"""
from . import formats, synthetic  # noqa: F401
from ._lib import AdaptiveParams, EngineError, lib  # noqa: F401
from .propagators import (  # noqa: F401
    BLANES_MOAN_14A, EXCHANGE_ALLGATHER, EXCHANGE_ALLREDUCE, MODE_PARITY, MODE_THROUGHPUT, POW_CORRECTLY_ROUNDED, POW_GLIBC, QUINLAN_TREMAINE_12, STORMER_13, Backward,
    ConstantThrust, CubicHermiteSpline, Ephemeris, Forward, LeastSquaresFit, NBodyPropagator, ShipStepError, SpacecraftPropagator,
    UniformSpline, default_adaptive_params, SHIP_METHOD_NAMES, VERNER87, CASH_KARP45, DORMAND_PRINCE54, DORMAND_PRINCE87,
    FEHLBERG45, TSITOURAS75, VERNER98, FINE45, fp64_fma_peak, gravity_eval, lsq_fit, nccl_unique_id, set_pair_variant,
)
