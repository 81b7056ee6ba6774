"""Deterministic synthetic inputs for the bench configs (nothing in the reference generates these).

Plummer sphere: G = 1, total mu = 1, mu_i = 1/N exactly, scale radius a = 1; radius r = a/sqrt(u^(-2/3) - 1)
(reject r > 20a); speed v = q*sqrt(2)*(1 + r^2)^(-1/4) with q from von-Neumann rejection on g(q) = q^2 (1-q^2)^3.5;
isotropic directions; centre-of-mass position and velocity removed.  RNG: numpy PCG64 seeded with 20260924.
"""
import numpy as np


def _iso(rng, n):
    z = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    s = np.sqrt(1.0 - z * z)
    return np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=1)


def plummer(n: int, seed: int = 20260924):
    rng = np.random.Generator(np.random.PCG64(seed))
    r = np.empty(0)
    while len(r) < n:
        u = rng.uniform(1e-12, 1.0, 2 * n)
        rr = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        r = np.concatenate([r, rr[rr <= 20.0]])
    r = r[:n]
    q = np.empty(0)
    while len(q) < n:
        x = rng.uniform(0.0, 1.0, 4 * n)
        y = rng.uniform(0.0, 0.1, 4 * n)
        q = np.concatenate([q, x[y < x * x * (1.0 - x * x) ** 3.5]])
    q = q[:n]
    v = q * np.sqrt(2.0) * (1.0 + r * r) ** (-0.25)
    pos = _iso(rng, n) * r[:, None]
    vel = _iso(rng, n) * v[:, None]
    mu = np.full(n, 1.0 / n)
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), mu
