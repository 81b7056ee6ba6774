"""Readers and writers for the reference's on-disk formats (either side of the hot path).

  state.json      {name?, epoch, bodies:[{name, mu, position[3], velocity[3]}]}   load/solar_system/loaders.rs:223-236
  ephemeris.json  {dt, settings:{name:{degree,count}}}                           load/solar_system/loaders.rs:299-309
  ship json       {name, integrator, tolerance, start, end, position, velocity, burns[]}   load/solar_system/mod.rs:208-226
  epoch strings   "YYYY-MM-DD HH:MM:SS[.fff]" -> f64 seconds since 1958-01-01 TAI          ftime/src/epoch.rs:19-44, :155-217
  durations       "<int> <unit> ..." summed in integer milliseconds, then * 1e-3            ftime/src/duration.rs:279-345

Output side: `format_epoch` / `format_duration` are the `Display` forms serde writes (ftime/src/epoch.rs:219-249,
ftime/src/duration.rs:217-277); `export_state` is the app's "export solar system state at an epoch"
(ephemeris_explorer/src/ui/windows/export.rs:222-257) fed by the batched device evaluation of the spline ephemeris;
`save_system` / `save_ship` write the directory layout `load_system` / `load_ship` read.
"""
import json
import math
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


def _civil_from_days(z: int):
    # inverse of _days_from_civil, as ftime/src/epoch.rs:266-280 (z = days since 1970-01-01)
    z += 719468
    era = (z if z >= 0 else z - 146096) // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + 3 if mp < 10 else mp - 9
    return (y + (1 if m <= 2 else 0), m, d)


def _round_half_away(x: float) -> int:
    # f64::round
    return int(math.floor(x + 0.5)) if x >= 0.0 else -int(math.floor(-x + 0.5))


def format_epoch(t: float) -> str:
    """`Epoch::to_string` (ftime/src/epoch.rs:219-249): "YYYY-MM-DD HH:MM:SS.mmm" from f64 seconds since 1958-01-01."""
    secs = math.floor(t)
    millis = _round_half_away((t - float(secs)) * 1000.0)
    secs = int(secs)
    if millis == 1000:
        secs += 1
        millis = 0
    days_since_1958, sod = divmod(secs, 86400)  # Python's divmod floors, like floor_div_mod (:251-261)
    year, month, day = _civil_from_days(days_since_1958 - 4383)
    return "%04d-%02d-%02d %02d:%02d:%02d.%03d" % (year, month, day, sod // 3600, (sod % 3600) // 60, sod % 60, millis)


def format_duration(seconds: float) -> str:
    """`Duration::to_string` (ftime/src/duration.rs:217-277): "[-]<y> y <d> d <h> h <m> m <s> s <ms> ms", zero units
    skipped, "0 s" when everything is zero, Julian years."""
    sign = "-" if math.copysign(1.0, seconds) < 0.0 else ""
    t = abs(seconds)
    whole = math.trunc(t)
    ms = _round_half_away((t - float(whole)) * 1e3)
    secs = int(whole)
    if ms == 1000:
        ms = 0
        secs += 1
    parts = []
    for unit, size in (("y", 31_557_600), ("d", 86_400), ("h", 3_600), ("m", 60), ("s", 1)):
        q, secs = divmod(secs, size)
        if q > 0:
            parts.append("%d %s" % (q, unit))
    if ms > 0:
        parts.append("%d ms" % ms)
    if not parts:
        parts.append("0 s")
    return sign + " ".join(parts)


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


def _vec(v) -> list:
    return [float(x) for x in np.asarray(v, dtype=np.float64).reshape(3)]


def state_document(name: str, epoch: float, names, mus, positions, velocities) -> dict:
    """The JSON document of `export_solar_system` (ephemeris_explorer/src/ui/windows/export.rs:229-249) = the schema
    the state loader reads back (load/solar_system/loaders.rs:223-236).  Floats are written in their shortest
    round-trip form (Python's repr, like serde_json's ryu), so a reload returns the same f64 bits."""
    return {
        "name": name,
        "epoch": format_epoch(epoch),
        "bodies": [
            {"name": str(n), "mu": float(m), "position": _vec(p), "velocity": _vec(v)}
            for n, m, p, v in zip(names, mus, positions, velocities)
        ],
    }


def export_state(ephemeris, names, mus, epoch: float, name: str = "Solar System") -> Optional[dict]:
    """`ExportType::State { epoch }`: every body's `Trajectory::state_vector(epoch)` (trajectory.rs:449-471), evaluated
    in one batched device call on the spline table, assembled into a state.json document.  Like the reference's
    `collect::<Option<Vec<_>>>()`, the result is None when any body's trajectory does not cover the epoch."""
    pos, vel, ok = ephemeris.evaluate(np.array([epoch], dtype=np.float64), velocities=True)
    if not bool(np.all(ok)):
        return None
    return state_document(name, epoch, names, mus, pos[0], vel[0])


def save_system(directory, system: SolarSystem) -> None:
    """Writes `state.json` (+ `ephemeris.json` when the sampling settings are known) in the reference's layout."""
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    doc = state_document(system.name, system.epoch, system.names, system.mu, system.position, system.velocity)
    (d / "state.json").write_text(json.dumps(doc, indent=4))
    if system.dt is not None and system.degree is not None and system.count is not None:
        eph = {"dt": format_duration(system.dt),
               "settings": {n: {"degree": int(g), "count": int(c)} for n, g, c in zip(system.names, system.degree, system.count)}}
        (d / "ephemeris.json").write_text(json.dumps(eph, indent=4))


def save_ship(path, ship: Ship, body_names: List[str]) -> None:
    """Writes the ship JSON `load_ship` reads (load/solar_system/mod.rs:208-226)."""
    burns = []
    for b in ship.burns:
        burns.append({"start": format_epoch(b.start), "duration": format_duration(b.end - b.start),
                      "acceleration": _vec(b.acceleration), "reference": body_names[b.reference] if b.reference >= 0 else None})
    doc = {"name": ship.name, "integrator": ship.integrator, "tolerance": float(ship.tolerance),
           "start": format_epoch(ship.start), "end": format_epoch(ship.end), "position": _vec(ship.position),
           "velocity": _vec(ship.velocity), "burns": burns}
    Path(path).write_text(json.dumps(doc, indent=4))
