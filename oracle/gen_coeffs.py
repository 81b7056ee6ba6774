#!/usr/bin/env python3
"""Generate the f64 coefficient tables used by the oracle and by the CUDA engine.

TEST/BUILD INFRASTRUCTURE.  This script READS the reference sources (integration/src/methods.rs and
integration/src/multistep/second_order/cowell.rs) where they lie under /root/reference, restates the
reference's `Ratio` -> f64 rule exactly, and writes derived hex-float tables.  It never copies source
text: only numbers leave this script, each already reduced to the one f64 the reference would use.

Rules restated (reference file:line):
  * `frac!(n, d)`      = Ratio::const_new -> gcd-normalised, denominator positive  (ratio.rs:45-49, :154-177)
  * `frac_f64!(x)`     = Ratio::from_f64: smallest p with x*10^p integral in f64   (ratio.rs:76-103)
  * `a.const_sub(b)`   = un-normalised lcm subtraction                             (ratio.rs:126-143)
  * `f64 * Ratio`      = x * (numer as f64 / denom as f64)                         (ratio.rs:221-228)
    (`as f64` from i128 is round-to-nearest-even == Python float(int)).
The generated header is committed; the GPU box never needs /root/reference.
"""
import math
import re
import sys
from pathlib import Path

REF = Path("/root/reference/integration/src")


def gcd(a, b):
    return math.gcd(abs(a), abs(b))


class Ratio:
    __slots__ = ("n", "d")

    def __init__(self, n, d):
        self.n, self.d = n, d

    @staticmethod
    def const_new(n, d):
        r = Ratio(n, d)
        if r.d == 0:
            return r
        if r.n == 0:
            r.d = 1
            return r
        if r.n == r.d:
            r.n = r.d = 1
            return r
        g = gcd(r.n, r.d)
        r.n //= g
        r.d //= g
        if r.d < 0:
            r.n, r.d = -r.n, -r.d
        assert -(2**127) <= r.n < 2**127 and r.d < 2**127, "does not fit i128"
        return r

    @staticmethod
    def from_f64(val):
        p = 0
        new_val = val
        while True:
            a = abs(new_val)
            # (a as u64) as f64 == a   (saturating truncation)
            t = min(int(a), 2**64 - 1)
            if float(t) == a:
                break
            p += 1
            new_val = val * float(10**p)
            assert not math.isinf(new_val)
        return Ratio.const_new(int(new_val), 10**p)

    def const_sub(self, rhs):
        if self.d == rhs.d:
            return Ratio(self.n - rhs.n, self.d)
        l = abs(self.d) * abs(rhs.d) // gcd(self.d, rhs.d)
        return Ratio(self.n * (l // self.d) - rhs.n * (l // rhs.d), l)

    def f64(self):
        # Rust integer division semantics are irrelevant: both casts are RNE, the divide is IEEE.
        return float(self.n) / float(self.d)


def block(src, name):
    """Text of `pub struct <name>;` ... up to the next `#[doc(hidden)]` (or end of module)."""
    m = re.search(r"pub struct %s;" % re.escape(name), src)
    assert m, name
    end = src.find("#[doc(hidden)]", m.end())
    return src[m.end(): end if end >= 0 else len(src)]


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def const_body(blk, name):
    """Right-hand side of `const <name>: ... = <rhs>;` at brace/bracket depth 0."""
    m = re.search(r"const %s\s*:[^=]*=" % name, blk)
    assert m, name
    i = m.end()
    depth = 0
    j = i
    while True:
        c = blk[j]
        if c in "[{(":
            depth += 1
        elif c in "]})":
            depth -= 1
        elif c == ";" and depth == 0:
            break
        j += 1
    return blk[i:j]


def parse_int(s):
    return int(s.replace("_", "").strip())


def parse_ratio_list(txt, env=None):
    """Parse a flat `&[ item, item, ... ]` of frac!/frac_f64!/X[i].const_sub(Y[i]) items."""
    txt = strip_comments(txt)
    out = []
    pos = 0
    item_re = re.compile(
        r"frac!\(\s*(-?[\d_]+)\s*,\s*(-?[\d_]+)\s*,?\s*\)"
        r"|frac_f64!\(\s*(-?[\d._eE+-]+)\s*\)"
        r"|(?:Self::)?(\w+)\[(\d+)\]\.const_sub\(\s*(?:Self::)?(\w+)\[(\d+)\]\s*\)"
        r"|Self::(\w+)\[(\d+)\]\[(\d+)\]"
    )
    for m in item_re.finditer(txt):
        if m.group(1) is not None:
            out.append(Ratio.const_new(parse_int(m.group(1)), parse_int(m.group(2))))
        elif m.group(3) is not None:
            out.append(Ratio.from_f64(float(m.group(3))))
        elif m.group(4) is not None:
            a = env[m.group(4)][int(m.group(5))]
            b = env[m.group(6)][int(m.group(7))]
            out.append(a.const_sub(b))
        else:  # `Self::A[6][0]`: a copy of another table's entry
            out.append(env[m.group(8)][int(m.group(9))][int(m.group(10))])
    return out


def parse_e(blk, B):
    """`const E` of an embedded method: either a flat list, or `{ const BH = [...]; [B[i].const_sub(BH[i]), ...] }`."""
    ebody = const_body(blk, "E")
    mbh = re.search(r"const BH\s*:[^=]*=", ebody)
    if not mbh:
        return parse_ratio_list(ebody)
    k = mbh.end()
    depth = 0
    j = k
    while True:
        c = ebody[j]
        if c in "[(":
            depth += 1
        elif c in "])":
            depth -= 1
        elif c == ";" and depth == 0:
            break
        j += 1
    BH = parse_ratio_list(ebody[k:j])
    assert len(BH) == len(B)
    return parse_ratio_list(ebody[j + 1:], env={"B": B, "BH": BH})


def parse_u16(blk, name):
    return parse_int(strip_comments(const_body(blk, name)))


def parse_bool(blk, name):
    return strip_comments(const_body(blk, name)).strip() == "true"


def tri(rows):
    """Strictly-lower-triangular rows packed: entry (s, j) at s*(s-1)/2 + j."""
    out = []
    for s, r in enumerate(rows):
        assert len(r) == s, (s, len(r))
        out.extend(x.f64() for x in r)
    return out


def parse_int_list(txt):
    txt = strip_comments(txt)
    return [parse_int(x) for x in re.findall(r"-?[\d_]+", txt[txt.index("["):])]


def split_rows(txt):
    """Split `&[ &[...], &[...], ... ]` into row texts."""
    txt = strip_comments(txt)
    i = txt.index("[") + 1
    rows = []
    depth = 0
    start = None
    for k in range(i, len(txt)):
        c = txt[k]
        if c == "[":
            if depth == 0:
                start = k
            depth += 1
        elif c == "]":
            depth -= 1
            if depth == 0:
                rows.append(txt[start:k + 1])
            elif depth < 0:
                break
    return rows


def hexf(x):
    return float(x).hex()


def emit_array(name, vals, cmt=""):
    s = "static const double %s[%d] = {%s\n" % (name, len(vals), (" // " + cmt) if cmt else "")
    for v in vals:
        s += "    %s, /* %.17g */\n" % (hexf(v), v)
    s += "};\n"
    return s


def main(out_paths):
    methods = (REF / "methods.rs").read_text()
    cowell = (REF / "multistep/second_order/cowell.rs").read_text()

    parts = []
    parts.append(
        "// GENERATED by oracle/gen_coeffs.py from the reference's coefficient tables -- do not edit.\n"
        "// Every entry is the single f64 the reference forms via `f64 * Ratio` = x * (numer as f64 / denom as f64)\n"
        "// (integration/src/ratio.rs:221-228).  Hex-float literals are exact.\n"
        "#pragma once\n\n"
    )

    # ---- BlanesMoan6B (methods.rs:1660-1687): SRKN kick/drift coefficients, FSAL
    bm = block(methods, "BlanesMoan6B")
    A = parse_ratio_list(const_body(bm, "A"))
    B = parse_ratio_list(const_body(bm, "B"))
    assert len(A) == 7 and len(B) == 7
    parts.append("#define EE_BM6B_STAGES 7\n")
    parts.append(emit_array("EE_BM6B_A", [r.f64() for r in A], "drift  a_s, integration/src/methods.rs:1667-1675"))
    parts.append(emit_array("EE_BM6B_B", [r.f64() for r in B], "kick   b_s, integration/src/methods.rs:1677-1685"))

    # ---- BlanesMoan14A (methods.rs:1730-1774): SRKN, 15 stages, FSAL (B[0] = 0), usable as the fixed-step method itself
    bm = block(methods, "BlanesMoan14A")
    A = parse_ratio_list(const_body(bm, "A"))
    B = parse_ratio_list(const_body(bm, "B"))
    assert len(A) == 15 and len(B) == 15
    parts.append("#define EE_BM14A_STAGES 15\n")
    parts.append(emit_array("EE_BM14A_A", [r.f64() for r in A], "drift  a_s, integration/src/methods.rs:1740-1756"))
    parts.append(emit_array("EE_BM14A_B", [r.f64() for r in B], "kick   b_s, integration/src/methods.rs:1758-1774"))

    # ---- QuinlanTremaine12 / Stormer13 (methods.rs:2005-2062)
    for nm, order, tag in (("QuinlanTremaine12", 12, "QT12"), ("Stormer13", 13, "ST13")):
        blk = block(methods, nm)
        alpha = parse_int_list(const_body(blk, "ALPHA"))
        beta_n = parse_int_list(const_body(blk, "BETA_N"))
        beta_d = parse_int(strip_comments(const_body(blk, "BETA_D")))
        assert len(alpha) == order + 1 and len(beta_n) == order + 1
        # `1.0 * Ratio::from_int(-ALPHA[j+1])` and `1.0 * Ratio::from_int(BETA_N[j+1])`, j = 0..order-1
        parts.append("#define EE_%s_ORDER %d\n" % (tag, order))
        parts.append(emit_array("EE_%s_NEG_ALPHA" % tag, [1.0 * (float(-a) / 1.0) for a in alpha[1:]],
                                "-alpha_{j+1}, second_order/mod.rs:106"))
        parts.append(emit_array("EE_%s_BETA" % tag, [1.0 * (float(b) / 1.0) for b in beta_n[1:]],
                                "beta_{j+1} numerators, second_order/mod.rs:107"))
        parts.append("static const double EE_%s_INV_BETA_D = %s; /* 1/%d, Ratio::from_recip */\n\n"
                     % (tag, hexf(1.0 / float(beta_d)), beta_d))

    # ---- Cowell velocity coefficients (cowell.rs:132-167)
    for order in (12, 13):
        m = re.search(r"for Cowell<%d>\s*\{(.*?)\n\}" % order, cowell, re.S)
        blk = m.group(1)
        bn = parse_int_list(const_body(blk, "BETA_N"))
        bd = parse_int(strip_comments(const_body(blk, "BETA_D")))
        assert len(bn) == order
        parts.append(emit_array("EE_COWELL%d_BETA" % order, [1.0 * (float(b) / 1.0) for b in bn],
                                "cowell.rs BETA_N (1.0 * Ratio::from_int)"))
        parts.append("static const double EE_COWELL%d_INV_BETA_D = %s; /* 1/%d */\n\n"
                     % (order, hexf(1.0 / float(bd)), bd))

    # ---- Verner87 (methods.rs:492-806): 13 stages, order 8(7)
    v = block(methods, "Verner87")
    rows = split_rows(const_body(v, "A"))
    assert len(rows) == 13
    Arows = [parse_ratio_list(r) for r in rows]
    for s, r in enumerate(Arows):
        assert len(r) == s, (s, len(r))
    Bv = parse_ratio_list(const_body(v, "B"))
    Cv = parse_ratio_list(const_body(v, "C"))
    assert len(Bv) == 13 and len(Cv) == 13
    ebody = const_body(v, "E")
    # inner `const BH: &[Ratio] = &[ ... ];` then the `&[ ... ]` of E proper
    mbh = re.search(r"const BH\s*:[^=]*=", ebody)
    k = mbh.end()
    depth = 0
    j = k
    while True:
        c = ebody[j]
        if c in "[(":
            depth += 1
        elif c in "])":
            depth -= 1
        elif c == ";" and depth == 0:
            break
        j += 1
    BH = parse_ratio_list(ebody[k:j])
    assert len(BH) == 13
    Ev = parse_ratio_list(ebody[j + 1:], env={"B": Bv, "BH": BH})
    assert len(Ev) == 13, len(Ev)
    flatA = []
    for s in range(13):
        row = [r.f64() for r in Arows[s]] + [0.0] * (13 - s)
        flatA.extend(row[:13])
    parts.append("#define EE_V87_STAGES 13\n#define EE_V87_ORDER 8\n#define EE_V87_ORDER_EMBEDDED 7\n")
    parts.append(emit_array("EE_V87_A", flatA, "row-major 13x13 (strictly lower part used), methods.rs:503-709"))
    parts.append(emit_array("EE_V87_B", [r.f64() for r in Bv], "methods.rs:711-734"))
    parts.append(emit_array("EE_V87_C", [r.f64() for r in Cv], "methods.rs:736-750"))
    parts.append(emit_array("EE_V87_E", [r.f64() for r in Ev], "E = B - Bhat, methods.rs:755-804"))

    # ---- every adaptive method a ship can select (ephemeris_explorer/src/flight_plan.rs:175-184), one generic layout:
    # A packed strictly-lower-triangular, B, C, E = B - Bhat; Fine45 (ERKNG) carries the pairs AP/AV, BP/BV, EP/EV.
    # Ids: 0 = Verner87 (the default of the shipped systems), then flight_plan.rs order.
    erk = [("Verner87", 0), ("CashKarp45", 1), ("DormandPrince54", 2), ("DormandPrince87", 3), ("Fehlberg45", 4),
           ("Tsitouras75", 5), ("Verner98", 6)]
    meta = {}
    for nm, mid in erk:
        blk = block(methods, nm)
        rows = [parse_ratio_list(r) for r in split_rows(const_body(blk, "A"))]
        Bm = parse_ratio_list(const_body(blk, "B"), env={"A": rows})
        Cm = parse_ratio_list(const_body(blk, "C"))
        Em = parse_e(blk, Bm)
        S = len(rows)
        assert len(Bm) == S and len(Cm) == S and len(Em) == S, (nm, S, len(Bm), len(Cm), len(Em))
        order, emb, fsal = parse_u16(blk, "ORDER"), parse_u16(blk, "ORDER_EMBEDDED"), parse_bool(blk, "FSAL")
        meta[mid] = (nm, S, fsal, min(order, emb), 0)
        parts.append("// %s: %d stages, order %d(%d), FSAL %s\n" % (nm, S, order, emb, "true" if fsal else "false"))
        parts.append(emit_array("EE_RK%d_A" % mid, tri(rows), "packed lower triangle, entry (s,j) at s(s-1)/2+j"))
        parts.append(emit_array("EE_RK%d_B" % mid, [r.f64() for r in Bm]))
        parts.append(emit_array("EE_RK%d_C" % mid, [r.f64() for r in Cm]))
        parts.append(emit_array("EE_RK%d_E" % mid, [r.f64() for r in Em], "E = B - Bhat"))
    blk = block(methods, "Fine45")
    AP = [parse_ratio_list(r) for r in split_rows(const_body(blk, "AP"))]
    AV = [parse_ratio_list(r) for r in split_rows(const_body(blk, "AV"))]
    env = {"AP": AP, "AV": AV}
    BP = parse_ratio_list(const_body(blk, "BP"), env=env)
    BV = parse_ratio_list(const_body(blk, "BV"), env=env)
    Cm = parse_ratio_list(const_body(blk, "C"))
    EP = parse_ratio_list(const_body(blk, "EP"))
    EV = parse_ratio_list(const_body(blk, "EV"))
    S = len(AP)
    assert S == 7 and all(len(x) == S for x in (AV, BP, BV, Cm, EP, EV))
    order, emb, fsal = parse_u16(blk, "ORDER"), parse_u16(blk, "ORDER_EMBEDDED"), parse_bool(blk, "FSAL")
    meta[7] = ("Fine45", S, fsal, min(order, emb), 1)
    parts.append("// Fine45 (ERKNG, y'' = f(t, y, y')): %d stages, order %d(%d), FSAL %s\n" % (S, order, emb, "true" if fsal else "false"))
    parts.append(emit_array("EE_RK7_A", tri(AP), "AP"))
    parts.append(emit_array("EE_RK7_A2", tri(AV), "AV"))
    parts.append(emit_array("EE_RK7_B", [r.f64() for r in BP], "BP"))
    parts.append(emit_array("EE_RK7_B2", [r.f64() for r in BV], "BV"))
    parts.append(emit_array("EE_RK7_C", [r.f64() for r in Cm]))
    parts.append(emit_array("EE_RK7_E", [r.f64() for r in EP], "EP"))
    parts.append(emit_array("EE_RK7_E2", [r.f64() for r in EV], "EV"))
    parts.append("#define EE_RK_METHODS 8\n#define EE_RK_MAX_STAGES 16\n")
    parts.append("struct EeRkTableau { const char* name; int stages, fsal, kord, kind; const double *a, *b, *c, *e, *a2, *b2, *e2; };\n")
    parts.append("// kord = min(ORDER, ORDER_EMBEDDED) = RKEmbedded::LOWER_ORDER; kind 0 = ERK (first order), 1 = ERKNG\n")
    parts.append("static const EeRkTableau EE_RK_TABLE[EE_RK_METHODS] = {\n")
    for mid in range(8):
        nm, S, fsal, kord, kind = meta[mid]
        assert S <= 16
        extra = "EE_RK7_A2, EE_RK7_B2, EE_RK7_E2" if kind else "nullptr, nullptr, nullptr"
        parts.append('    {"%s", %d, %d, %d, %d, EE_RK%d_A, EE_RK%d_B, EE_RK%d_C, EE_RK%d_E, %s},\n'
                     % (nm, S, 1 if fsal else 0, kord, kind, mid, mid, mid, mid, extra))
    parts.append("};\n")

    text = "".join(parts)
    for p in out_paths:
        Path(p).write_text(text)
        print("wrote", p, len(text), "bytes")


if __name__ == "__main__":
    here = Path(__file__).resolve().parent
    outs = sys.argv[1:] or [str(here / "ee_oracle_coeffs.h"),
                            str(here.parent / "ephemeris-explorer_b200" / "csrc" / "ee_coeffs.h")]
    main(outs)
