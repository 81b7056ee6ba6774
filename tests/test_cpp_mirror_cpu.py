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
