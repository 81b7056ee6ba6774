"""Deterministic synthetic inputs for the bench configs (nothing in the reference generates these).

Plummer sphere: G = 1, total mu = 1, mu_i = 1/N exactly, scale radius a = 1; radius r = a/sqrt(u^(-2/3) - 1)
(reject r > 20a); speed v = q*sqrt(2)*(1 + r^2)^(-1/4) with q from von-Neumann rejection on g(q) = q^2 (1-q^2)^3.5;
isotropic directions; centre-of-mass position and velocity removed.  RNG: xoshiro256** seeded by splitmix64(seed).
The generator itself is host code inside libee_b200.so (`ee_host_plummer`, csrc/ee_capi.cu) so that every host language
-- this module, the C++ mirror, the Rust shim -- regenerates the same bits.
"""
import ctypes as C

import numpy as np


def plummer(n: int, seed: int = 20260924):
    from ._lib import check, lib
    pos = np.zeros((n, 3))
    vel = np.zeros((n, 3))
    mu = np.zeros(n)
    dp = C.POINTER(C.c_double)
    check(lib.ee_host_plummer(int(n), int(seed), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp), mu.ctypes.data_as(dp)),
          "ee_host_plummer")
    return pos, vel, mu
