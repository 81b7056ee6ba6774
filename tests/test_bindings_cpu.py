"""The three bindings of the C ABI agree with include/ee_b200.h argument for argument.

The Rust shim (shim/src/lib.rs) cannot be compiled in this image (no rustc), so its `extern "C"` blocks and `repr(C)`
structs are checked textually against the header's prototypes: same symbol, same number of arguments, same scalar /
pointer type in every position, same return type.  The ctypes table (ephemeris-explorer_b200/_lib.py) is checked the
same way; a drift between the header and a binding would otherwise show up as a corrupted call on the GPU box only.
"""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ee_b200.h"
SHIM = ROOT / "shim" / "src" / "lib.rs"

# canonical spelling: scalar names as in Rust ("f64", "i32", ...), pointers as "*f64" (constness dropped: ctypes has
# none), any pointer to an opaque handle / void / struct as "*opaque", a pointer to such a pointer as "**opaque"
C_SCALARS = {
    "double": "f64", "int32_t": "i32", "int64_t": "i64", "uint32_t": "u32", "uint64_t": "u64", "char": "c_char",
    "void": "void",
}
OPAQUE_C = {"ee_nbody", "ee_ephem", "ee_ships", "void", "ee_adaptive_params"}
OPAQUE_RUST = {"EeNBody", "EeEphem", "EeShips", "c_void", "EeAdaptiveParams", "u8"}


def strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def canon_c(decl):
    """'const double* positions' -> '*f64';  'ee_nbody** out' -> '**opaque';  'double t0' -> 'f64'."""
    decl = decl.replace("const", " ").strip()
    stars = decl.count("*")
    words = decl.replace("*", " ").split()
    base = words[0]
    if stars and base in OPAQUE_C:
        return "*" * stars + "opaque"
    return "*" * stars + C_SCALARS[base]


def header_prototypes():
    text = strip_c_comments(HEADER.read_text())
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?\**)\s*\b(ee_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = [] if args in ("", "void") else [canon_c(a) for a in args.split(",")]
        protos[name] = (canon_c(ret + " r") if ret != "void" else "void", arglist)
    return protos


def canon_rust(ty):
    ty = ty.strip()
    stars = 0
    while ty.startswith("*"):
        ty = re.sub(r"^\*\s*(const|mut)\s*", "", ty)
        stars += 1
    if stars and ty in OPAQUE_RUST:
        return "*" * stars + "opaque"
    return "*" * stars + ty


def shim_prototypes():
    text = re.sub(r"//[^\n]*", "", SHIM.read_text())
    protos = {}
    for block in re.finditer(r'extern\s+"C"\s*\{(.*?)\n\}', text, flags=re.S):
        for m in re.finditer(r"fn\s+(ee_[a-z0-9_]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block.group(1), flags=re.S):
            name, args, ret = m.group(1), m.group(2), m.group(3)
            arglist = [canon_rust(a.split(":", 1)[1]) for a in args.split(",") if ":" in a]
            protos[name] = (canon_rust(ret) if ret else "void", arglist)
    return protos


CTYPES_SCALARS = {
    C.c_double: "f64", C.c_int32: "i32", C.c_int64: "i64", C.c_uint32: "u32", C.c_uint64: "u64", C.c_void_p: "*opaque",
    C.c_char_p: "*c_char", None: "void",
}


def canon_ctypes(t):
    if t in CTYPES_SCALARS:
        return CTYPES_SCALARS[t]
    inner = t._type_  # POINTER(x)
    if inner is C.c_void_p or issubclass(inner, C.Structure):
        return "**opaque" if inner is C.c_void_p else "*opaque"
    return "*" + CTYPES_SCALARS[inner]


def test_header_parses_completely():
    protos = header_prototypes()
    text = strip_c_comments(HEADER.read_text())
    names = set(re.findall(r"\b(ee_[a-z0-9_]+)\s*\(", text))
    assert names == set(protos), names ^ set(protos)
    assert protos["ee_nbody_create"] == ("i32", ["i64", "*f64", "*f64", "*f64", "f64", "f64", "i32", "i32", "i32", "**opaque"])
    assert protos["ee_last_error"] == ("*c_char", [])
    assert protos["ee_nbody_destroy"] == ("void", ["*opaque"])


def test_ctypes_table_matches_the_header_argument_for_argument():
    import ephemeris_explorer_b200 as ee
    protos = header_prototypes()
    assert set(ee._lib.SIGNATURES) == set(protos), set(ee._lib.SIGNATURES) ^ set(protos)
    for name, (res, args) in ee._lib.SIGNATURES.items():
        got = (canon_ctypes(res), [canon_ctypes(a) for a in args])
        want = protos[name]
        # ctypes has no typed opaque pointers: a `void*` argument stands for any single-level pointer to a handle/blob
        assert got[0] == want[0], (name, got[0], want[0])
        assert len(got[1]) == len(want[1]), (name, got[1], want[1])
        for i, (g, w) in enumerate(zip(got[1], want[1])):
            assert g == w, "%s argument %d: ctypes %s, header %s" % (name, i, g, w)


def test_rust_shim_extern_blocks_match_the_header_argument_for_argument():
    protos = header_prototypes()
    shim = shim_prototypes()
    assert len(shim) >= 20
    for name, got in shim.items():
        assert name in protos, "the shim binds %s, which the header does not declare" % name
        want = protos[name]
        assert got[0] == want[0], (name, got[0], want[0])
        assert len(got[1]) == len(want[1]), (name, got[1], want[1])
        for i, (g, w) in enumerate(zip(got[1], want[1])):
            assert g == w, "%s argument %d: shim %s, header %s" % (name, i, g, w)
    # the trait surface the Planner needs (prediction.rs:31-37) is bound
    for need in ("ee_nbody_create", "ee_nbody_set_solout", "ee_nbody_step", "ee_nbody_take_solution", "ee_nbody_clone",
                 "ee_nbody_solution_time", "ee_nbody_has_reached", "ee_nbody_destroy", "ee_ships_create",
                 "ee_ships_step_to", "ee_ships_take_knots"):
        assert need in shim, need


def struct_fields_c(name):
    text = strip_c_comments(HEADER.read_text())
    body = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ty, rest = stmt.split(None, 1)
        for f in rest.split(","):
            fields.append((f.strip(), C_SCALARS[ty]))
    return fields


def test_adaptive_params_struct_layout_is_the_same_in_all_three_bindings():
    import ephemeris_explorer_b200 as ee
    want = struct_fields_c("ee_adaptive_params")
    assert [f for f, _ in want][:3] == ["h_init", "h_max", "tol_position"]
    got_py = [(n, CTYPES_SCALARS[t]) for n, t in ee._lib.AdaptiveParams._fields_]
    assert got_py == want
    text = re.sub(r"//[^\n]*", "", SHIM.read_text())
    m = re.search(r"#\[repr\(C\)\]\s*(?:#\[[^\]]*\]\s*)*pub\s+struct\s+EeAdaptiveParams\s*\{(.*?)\}", text, flags=re.S)
    assert m, "EeAdaptiveParams must be repr(C)"
    got_rs = [(a.split(":")[0].replace("pub", "").strip(), a.split(":")[1].strip()) for a in m.group(1).split(",") if ":" in a]
    assert got_rs == want
    assert C.sizeof(ee._lib.AdaptiveParams) == 7 * 8 + 3 * 4 + 4  # 7 doubles + 3 u32, padded to 8


def test_status_codes_agree():
    """ee_status (header) == StepError order of integration/src/lib.rs:312-318 as the shim and the Python mirror map it."""
    import ephemeris_explorer_b200 as ee
    text = strip_c_comments(HEADER.read_text())
    codes = {k: int(v) for k, v in re.findall(r"\b(EE_[A-Z_]+)\s*=\s*(\d+)", text)}
    assert codes["EE_OK"] == 0
    for k, v in (("EE_ERR_INVALID", 100), ("EE_ERR_CUDA", 101), ("EE_ERR_NCCL", 102), ("EE_ERR_UNSUPPORTED", 103)):
        assert codes[k] == v and v in ee._lib.STATUS_NAMES
    shim = SHIM.read_text()
    for code, variant in ((1, "StepSizeUnderflow"), (2, "MaxIterationsReached"), (3, "BoundReached"), (4, "EvalFailed")):
        assert re.search(r"\b%d\s*=>\s*Err\(CudaPropagatorError::%s\)" % (code, variant), shim), variant
        assert code in ee._lib.STATUS_NAMES
