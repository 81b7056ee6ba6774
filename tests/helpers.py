"""Shared test helpers (CPU side)."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SYSTEMS = ROOT / "tests" / "golden" / "systems"


def load_system(name):
    from ephemeris_explorer_b200 import formats
    return formats.load_system(SYSTEMS / (name + ".json"))


def energy(pos, vel, mu):
    """Specific 'energy' with G folded into mu: sum mu_i v_i^2/2 - sum_{i<j} mu_i mu_j / r_ij  (units of G*...)."""
    ke = 0.5 * np.sum(mu * np.sum(vel * vel, axis=1))
    pe = 0.0
    n = len(mu)
    for i in range(n):
        d = pos[i + 1:] - pos[i]
        r = np.sqrt(np.sum(d * d, axis=1))
        pe -= np.sum(mu[i] * mu[i + 1:] / r)
    return ke + pe


def kepler_two_body(mu1, mu2, a, e, t):
    """Analytic positions/velocities of two bodies on a bound relative orbit (periapsis at t=0 on +x), barycentric."""
    mu = mu1 + mu2
    n = np.sqrt(mu / a**3)
    M = n * t
    E = M
    for _ in range(100):
        E = E - (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
    x = a * (np.cos(E) - e)
    y = a * np.sqrt(1 - e * e) * np.sin(E)
    r = a * (1 - e * np.cos(E))
    vx = -a * a * n * np.sin(E) / r
    vy = a * a * n * np.sqrt(1 - e * e) * np.cos(E) / r
    rel = np.array([x, y, 0.0])
    relv = np.array([vx, vy, 0.0])
    p1, p2 = -mu2 / mu * rel, mu1 / mu * rel
    v1, v2 = -mu2 / mu * relv, mu1 / mu * relv
    return np.array([p1, p2]), np.array([v1, v2])


def rel_err(a, b):
    """max over bodies of |a - b| / |b| (the parity metric of SURVEY.md 8d)."""
    num = np.sqrt(np.sum((a - b) ** 2, axis=-1))
    den = np.sqrt(np.sum(b**2, axis=-1))
    return float(np.max(num / den))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(a.view(np.uint64) == b.view(np.uint64)))


# ephemeris/tests/spacecraft_propagation.rs:357-371 (hours) for the 10-body system
SHIP_TEST_PERIOD_HOURS = [72.0, 12.0, 60.0, 18.0, 6.0, 72.0, 150.0, 150.0, 150.0, 150.0]
SHIP_TEST_DEGREES = [6, 7, 7, 7, 6, 7, 7, 6, 6, 5]
