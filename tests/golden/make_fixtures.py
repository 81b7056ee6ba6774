#!/usr/bin/env python3
"""Regenerate tests/golden/systems/*.json from the reference's checked-in data files (run in the build container only).

The reference's `systems/*/{state.json,ephemeris.json,ships/*.json}` are the only offline fixtures it has
(SURVEY.md section 4).  They are data, not source.  Each system is folded into ONE columnar fixture file
(`ee-fixture-1`): body names / mu / positions / velocities / spline degree / sample count as parallel arrays, the
epoch and dt strings verbatim, ships as records.  Floats survive exactly (Python repr is the shortest round-trip
decimal).  ephemeris-explorer_b200/formats.py reads both this layout and the reference's own directory layout.
"""
import json
from pathlib import Path

SRC = Path("/root/reference/systems")
DST = Path(__file__).resolve().parent / "systems"


def main():
    DST.mkdir(parents=True, exist_ok=True)
    for sysdir in sorted(p for p in SRC.iterdir() if p.is_dir()):
        st = json.loads((sysdir / "state.json").read_text())
        ep = json.loads((sysdir / "ephemeris.json").read_text())
        names = [b["name"] for b in st["bodies"]]
        fx = {
            "schema": "ee-fixture-1",
            "source": "systems/%s" % sysdir.name,
            "name": st.get("name", sysdir.name),
            "epoch": st["epoch"],
            "dt": ep["dt"],
            "bodies": {
                "name": names,
                "mu": [b["mu"] for b in st["bodies"]],
                "position": [b["position"] for b in st["bodies"]],
                "velocity": [b["velocity"] for b in st["bodies"]],
                "degree": [ep["settings"][n]["degree"] for n in names],
                "count": [ep["settings"][n]["count"] for n in names],
            },
            "ships": [json.loads(s.read_text()) for s in sorted((sysdir / "ships").glob("*.json"))],
        }
        out = DST / (sysdir.name + ".json")
        out.write_text(json.dumps(fx, separators=(",", ":")))
        print("wrote", out)


if __name__ == "__main__":
    main()
