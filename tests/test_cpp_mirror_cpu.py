"""The header-only C++ mirror of the reference's trait surface compiles and links against libee_b200.so (no GPU run)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SRC = r'''
#include "ee_b200.hpp"
#include <cstdio>
int main() {
    std::vector<ee::Vec3> p{{0, 0, 0}, {1, 0, 0}}, v{{0, 0, 0}, {0, 1, 0}};
    std::vector<double> mu{1.0, 1e-3};
    try {
        ee::NBodyPropagator prop(ee::Forward{0.01}, 0.0, p, v, mu);
        prop.with_solout(0.01, {0.01, 0.02}, {6, 6});
        prop.step(20);
        auto sol = prop.take_solution();
        std::printf("steps ok, polys %zu %zu\n", sol[0].polynomials.size(), sol[1].polynomials.size());
    } catch (const ee::EngineError& e) {
        std::printf("engine error %d\n", e.code);  // expected on a box without a GPU: 101
        return e.code == EE_ERR_CUDA ? 0 : 1;
    }
    return 0;
}
'''


def test_cpp_mirror_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    import ephemeris_explorer_b200 as ee  # builds the library if missing
    src = tmp_path / "mirror.cpp"
    src.write_text(SRC)
    exe = tmp_path / "mirror"
    libdir = ee._lib.LIB_PATH.parent
    subprocess.check_call(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(libdir),
                           "-lee_b200", "-Wl,-rpath," + str(libdir)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "engine error 101" in out.stdout or "steps ok" in out.stdout


OPS_SRC = r"""
#include "ee_b200.hpp"
#include <cstdio>
static ee::UniformSpline make(int n, double start, double interval = 8.0) {
    ee::UniformSpline u;
    u.start = start;
    u.interval = interval;
    for (int i = 0; i < n; ++i) u.polynomials.push_back(ee::Polynomial{{ee::Vec3{(double)i, 0, 0}}});
    return u;
}
static void show(const char* what, const ee::UniformSpline& u) {
    std::printf("%s %.17g %.17g", what, u.start, u.end());
    for (auto& p : u.polynomials) std::printf(" %g", p.coeffs[0][0]);
    std::printf("\n");
}
int main() {
    ee::UniformSpline u = make(6, 100.0);
    std::printf("idx %lld %lld %lld %lld %lld %lld\n", (long long)u.get_index(100.0), (long long)u.get_index(108.0),
                (long long)u.get_index(148.0), (long long)u.get_index_exclusive(108.0), (long long)u.get_index_exclusive(148.0),
                (long long)u.get_index_exclusive(148.001));
    std::printf("contains %d %d %d %d\n", (int)u.contains(100.0), (int)u.contains(148.0), (int)u.contains(99.0), (int)u.contains(-0.0 + 100.0));
    u.clear_after(116.0);
    show("clear_after", u);
    u = make(6, 100.0);
    u.clear_before(116.0);
    show("clear_before", u);
    u.push_front(ee::Polynomial{{ee::Vec3{9, 0, 0}}});
    show("push_front", u);
    ee::UniformSpline a = make(2, 100.0);
    a.append(make(3, 116.0));
    a.prepend(make(2, 84.0));
    show("joined", a);
    ee::UniformSpline w = make(4, 100.0);
    w.merge_forward(make(3, 116.0));   // drops polynomials 2, 3, appends 0, 1, 2
    show("merge_forward", w);
    ee::UniformSpline b = make(4, 100.0);
    b.merge_backward(make(3, 92.0));   // new piece ends at 116: drops polynomials 0, 1, prepends 0, 1, 2
    show("merge_backward", b);
    bool threw = false;
    try { a.append(make(1, 999.0)); } catch (const std::logic_error&) { threw = true; }
    std::printf("threw %d\n", (int)threw);
    std::vector<ee::Knot> l{{0.0, 0}, {1.0, 1}, {2.5, 2}, {4.0, 3}}, r{{2.5, 20}, {3.0, 21}};
    ee::join(l, r);
    std::printf("join");
    for (auto& k : l) std::printf(" %g:%g", k[0], k[1]);
    std::printf("\n");
    return 0;
}
"""


def test_cpp_trajectory_containers_agree_with_the_python_mirror(tmp_path):
    """UniformSpline's container operations and the ship `join` in include/ee_b200.hpp (host code, no GPU) against the same
    calls on the Python mirror, whose semantics tests/test_trajectory_ops_cpu.py pins on oracle solutions."""
    import numpy as np
    import ephemeris_explorer_b200 as ee
    src = tmp_path / "ops.cpp"
    src.write_text(OPS_SRC)
    exe = tmp_path / "ops"
    libdir = ee._lib.LIB_PATH.parent
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(libdir),
                           "-lee_b200", "-Wl,-rpath," + str(libdir)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr

    def make(n, start, interval=8.0):
        return ee.UniformSpline(start, interval, [np.array([[float(i), 0.0, 0.0]]) for i in range(n)])

    def show(what, u):
        return "%s %s %s" % (what, "%.17g" % u.start, "%.17g" % u.end()) + "".join(" %g" % p[0, 0] for p in u.polynomials)

    def idx(v):
        return -1 if v is None else v

    lines = []
    u = make(6, 100.0)
    lines.append("idx %d %d %d %d %d %d" % (idx(u.get_index(100.0)), idx(u.get_index(108.0)), idx(u.get_index(148.0)),
                                            idx(u.get_index_exclusive(108.0)), idx(u.get_index_exclusive(148.0)),
                                            idx(u.get_index_exclusive(148.001))))
    lines.append("contains %d %d %d %d" % (u.contains(100.0), u.contains(148.0), u.contains(99.0), u.contains(-0.0 + 100.0)))
    u.clear_after(116.0)
    lines.append(show("clear_after", u))
    u = make(6, 100.0)
    u.clear_before(116.0)
    lines.append(show("clear_before", u))
    u.push_front(np.array([[9.0, 0.0, 0.0]]))
    lines.append(show("push_front", u))
    a = make(2, 100.0)
    a.append(make(3, 116.0))
    a.prepend(make(2, 84.0))
    lines.append(show("joined", a))
    w = make(4, 100.0)
    w.merge_forward(make(3, 116.0))
    lines.append(show("merge_forward", w))
    b = make(4, 100.0)
    b.merge_backward(make(3, 92.0))
    lines.append(show("merge_backward", b))
    lines.append("threw 1")
    k = ee.CubicHermiteSpline(np.array([[0.0, 0] + [0] * 5, [1.0, 1] + [0] * 5, [2.5, 2] + [0] * 5, [4.0, 3] + [0] * 5], dtype=float))
    k.join(ee.CubicHermiteSpline(np.array([[2.5, 20] + [0] * 5, [3.0, 21] + [0] * 5], dtype=float)))
    lines.append("join" + "".join(" %g:%g" % (r[0], r[1]) for r in k.knots))
    assert out.stdout.strip().splitlines() == lines, out.stdout
    assert lines[6] == "merge_forward 100 140 0 1 0 1 2" and lines[7] == "merge_backward 92 132 0 1 2 2 3"
