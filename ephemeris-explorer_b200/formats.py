"""Readers for the reference's on-disk formats (inputs to the harness).

  state.json      {name?, epoch, bodies:[{name, mu, position[3], velocity[3]}]}   load/solar_system/loaders.rs:223-236
  ephemeris.json  {dt, settings:{name:{degree,count}}}                           load/solar_system/loaders.rs:299-309
  ship json       {name, integrator, tolerance, start, end, position, velocity, burns[]}   load/solar_system/mod.rs:208-226
  epoch strings   "YYYY-MM-DD HH:MM:SS[.fff]" -> f64 seconds since 1958-01-01 TAI          ftime/src/epoch.rs:19-44, :155-217
  durations       "<int> <unit> ..." summed in integer milliseconds, then * 1e-3            ftime/src/duration.rs:279-345
"""
import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

import numpy as np


def _days_from_civil(y: int, m: int, d: int) -> int:
    # Howard Hinnant's algorithm, as ftime/src/epoch.rs:244-253
    y -= 1 if m <= 2 else 0
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    mp = m - 3 if m > 2 else m + 9
    doy = (153 * mp + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


def parse_epoch(s: str) -> float:
    date_s, time_s = s.split(" ", 1)
    year, month, day = (int(x) for x in date_s.split("-", 2))
    if "." in time_s:
        hms, frac = time_s.split(".", 1)
        if not frac or not frac.isdigit():
            raise ValueError("bad fractional seconds in %r" % s)
        digits = frac[:3]
        millis = int(digits) * 10 ** (3 - len(digits))
    else:
        hms, millis = time_s, 0
    hour, minute, second = (int(x) for x in hms.split(":", 2))
    if not (1 <= month <= 12) or hour > 23 or minute > 59 or second > 59:
        raise ValueError("epoch out of range: %r" % s)
    days = _days_from_civil(year, month, day) - _days_from_civil(1958, 1, 1)
    secs = days * 86400 + hour * 3600 + minute * 60 + second
    return float(secs) + float(millis) / 1000.0


_UNITS_MS = {}
for _names, _ms in (
    (("y", "yr", "yrs", "year", "years"), 31_557_600_000),  # ftime MS_PER_YEAR = 365.25 d
    (("d", "day", "days"), 86_400_000),
    (("h", "hr", "hrs", "hour", "hours"), 3_600_000),
    (("m", "min", "mins", "minute", "minutes"), 60_000),
    (("s", "sec", "secs", "second", "seconds"), 1_000),
    (("ms", "msec", "msecs", "millisecond", "milliseconds"), 1),
):
    for _n in _names:
        _UNITS_MS[_n] = _ms


def parse_duration(s: str) -> float:
    s = s.strip()
    if not s:
        raise ValueError("empty duration")
    sign = 1.0
    if s[0] == "+":
        s = s[1:].lstrip()
    elif s[0] == "-":
        sign = -1.0
        s = s[1:].lstrip()
    toks = s.split()
    total_ms = 0
    for num, unit in zip(toks[0::2], toks[1::2]):
        total_ms += int(num) * _UNITS_MS[unit.strip().lower()]
    return sign * (float(total_ms) * 1e-3)


@dataclass
class SolarSystem:
    name: str
    epoch: float
    names: List[str]
    mu: np.ndarray
    position: np.ndarray  # (n, 3) km
    velocity: np.ndarray  # (n, 3) km/s
    dt: Optional[float] = None  # ephemeris.json
    degree: Optional[np.ndarray] = None
    count: Optional[np.ndarray] = None

    @property
    def sample_period(self) -> np.ndarray:
        # load/mod.rs:325: sample_period = ephemerides.dt * interpolation.count as f64
        return self.dt * self.count.astype(np.float64)


def _load_fixture(path: Path) -> dict:
    fx = json.loads(path.read_text())
    if fx.get("schema") != "ee-fixture-1":
        raise ValueError("%s is not an ee-fixture-1 file" % path)
    return fx


def load_system(directory) -> SolarSystem:
    """Reads a system from the reference's directory layout (state.json + ephemeris.json) or from one columnar
    `ee-fixture-1` file (tests/golden/make_fixtures.py)."""
    d = Path(directory)
    if d.is_file():
        fx = _load_fixture(d)
        b = fx["bodies"]
        return SolarSystem(
            name=fx["name"], epoch=parse_epoch(fx["epoch"]), names=list(b["name"]),
            mu=np.array(b["mu"], dtype=np.float64), position=np.array(b["position"], dtype=np.float64),
            velocity=np.array(b["velocity"], dtype=np.float64), dt=parse_duration(fx["dt"]),
            degree=np.array(b["degree"], dtype=np.int32), count=np.array(b["count"], dtype=np.int64))
    st = json.loads((d / "state.json").read_text())
    bodies = st["bodies"]
    sys_ = SolarSystem(
        name=st.get("name", d.name),
        epoch=parse_epoch(st["epoch"]),
        names=[b["name"] for b in bodies],
        mu=np.array([b["mu"] for b in bodies], dtype=np.float64),
        position=np.array([b["position"] for b in bodies], dtype=np.float64),
        velocity=np.array([b["velocity"] for b in bodies], dtype=np.float64),
    )
    ep = d / "ephemeris.json"
    if ep.exists():
        e = json.loads(ep.read_text())
        sys_.dt = parse_duration(e["dt"])
        sys_.degree = np.array([e["settings"][n]["degree"] for n in sys_.names], dtype=np.int32)
        sys_.count = np.array([e["settings"][n]["count"] for n in sys_.names], dtype=np.int64)
    return sys_


def soi_radii(system: SolarSystem) -> np.ndarray:
    """SphereOfInfluence per body as the reference's loader derives them (ephemeris_explorer/src/load/mod.rs:283-306):
    bodies by descending mu; a body's parent candidates are the heavier bodies whose own sphere contains it; radius =
    distance * (mu / mu_parent)^(2/5) (dynamics/spacecraft.rs:34-39), the smallest candidate wins; no parent = INFINITY."""
    order = sorted(range(len(system.names)), key=lambda i: -system.mu[i])  # stable, like sort_by(total_cmp) on b vs a
    radius = {}
    for i in order:
        best = np.inf
        for j in list(radius.keys()):  # insertion order = the sorted order so far
            d = float(np.linalg.norm(system.position[i] - system.position[j]))
            if d < radius[j]:
                r = d * (system.mu[i] / system.mu[j]) ** (2.0 / 5.0)
                if r < best:
                    best = r
        radius[i] = best
    return np.array([radius[i] for i in range(len(system.names))], dtype=np.float64)


@dataclass
class Burn:
    start: float
    end: float
    acceleration: np.ndarray
    reference: int  # body index, -1 = inertial


@dataclass
class Ship:
    name: str
    integrator: str
    tolerance: float
    start: float
    end: float
    position: np.ndarray
    velocity: np.ndarray
    burns: List[Burn] = field(default_factory=list)


def load_ship(path, body_names: List[str], name: Optional[str] = None) -> Ship:
    """Reads a ship from the reference's ship JSON, or (with `name`) from the `ships` list of an ee-fixture-1 file."""
    if name is not None:
        j = next(s for s in _load_fixture(Path(path))["ships"] if s["name"] == name)
    else:
        j = json.loads(Path(path).read_text())
    burns = []
    for b in j.get("burns", []):
        st = parse_epoch(b["start"])
        ref = b.get("reference")
        burns.append(Burn(st, st + parse_duration(b["duration"]), np.array(b["acceleration"], dtype=np.float64),
                          body_names.index(ref) if ref is not None else -1))
    return Ship(j["name"], j.get("integrator", "Verner87"), float(j.get("tolerance", 1e-3)), parse_epoch(j["start"]),
                parse_epoch(j["end"]), np.array(j["position"], dtype=np.float64), np.array(j["velocity"], dtype=np.float64),
                burns)
