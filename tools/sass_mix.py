#!/usr/bin/env python3
"""Static instruction mix of the library's kernels from `cuobjdump -sass` (no GPU needed).

For every kernel whose name contains one of the given substrings: the opcode histogram of the whole function and of each
loop (a backward branch and the addresses it spans), innermost loops first.  Used to check the per-pair instruction budget
the roofline discussion in DESIGN.md relies on (20 FP64 instructions per unordered pair in `k_accel_sym`) against the
code the compiler actually emitted.

    python tools/sass_mix.py k_accel_sym k_sym_reduce > profiles/r02/sass_mix_sym.txt
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "ephemeris-explorer_b200" / "libee_b200.so"
INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = INS.match(line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2), m.group(3)))
    if name:
        yield name, body


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    except OSError:
        return name


def base(op):
    return op.split(".")[0]


def mix(ins):
    c = Counter()
    for _, op, _ in ins:
        b = base(op)
        c[op if b == "MUFU" else b] += 1
    return c


def describe(c, total):
    fp64 = sum(c[k] for k in FP64)
    keys = ["DFMA", "DADD", "DMUL"] + sorted(k for k in c if k.startswith("MUFU")) + ["LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ATOMG", "RED"]
    parts = ["%s %d" % (k, c[k]) for k in keys if c[k]]
    return "%d instructions, %d FP64 (%.0f %%): %s" % (total, fp64, 100.0 * fp64 / max(total, 1), ", ".join(parts))


def loops(ins):
    addr = {a: i for i, (a, _, _) in enumerate(ins)}
    found = []
    for i, (a, op, args) in enumerate(ins):
        if base(op) != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)", args)
        if not m:
            continue
        t = int(m.group(1), 16)
        if t < a and t in addr:
            found.append((addr[t], i))
    found.sort(key=lambda p: p[1] - p[0])
    return found


def main():
    wanted = sys.argv[1:] or ["k_accel_sym"]
    for name, ins in functions():
        if not any(w in name for w in wanted):
            continue
        print("== %s" % demangle(name))
        print("   whole function: " + describe(mix(ins), len(ins)))
        for lo, hi in loops(ins):
            body = ins[lo:hi + 1]
            c = mix(body)
            if sum(c[k] for k in FP64) < 8:
                continue
            print("   loop 0x%04x..0x%04x: %s" % (ins[lo][0], ins[hi][0], describe(c, len(body))))
        print()


if __name__ == "__main__":
    main()
