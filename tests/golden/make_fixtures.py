#!/usr/bin/env python3
"""Regenerate tests/golden/systems/ from the reference's checked-in data files (run in the build container only).

The reference's `systems/*/{state.json,ephemeris.json,ships/*.json}` are the only offline fixtures it has
(SURVEY.md section 4).  They are data, not source; they are re-serialised compactly here, keeping the reference's
schemas so ephemeris-explorer_b200/formats.py reads either copy.  Floats survive exactly: Python's repr is the
shortest round-trip decimal.
"""
import json
from pathlib import Path

SRC = Path("/root/reference/systems")
DST = Path(__file__).resolve().parent / "systems"


def main():
    for sysdir in sorted(p for p in SRC.iterdir() if p.is_dir()):
        out = DST / sysdir.name
        (out / "ships").mkdir(parents=True, exist_ok=True)
        for name in ("state.json", "ephemeris.json"):
            data = json.loads((sysdir / name).read_text())
            (out / name).write_text(json.dumps(data, separators=(",", ":")))
        for ship in sorted((sysdir / "ships").glob("*.json")):
            data = json.loads(ship.read_text())
            (out / "ships" / ship.name).write_text(json.dumps(data, separators=(",", ":")))
        print("wrote", out)


if __name__ == "__main__":
    main()
