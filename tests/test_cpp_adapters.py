"""Part 1 of include/momentum_b200_adapters.hpp compiles as plain C++17 against the C-ABI and links the product library."""
import os
import subprocess
import tempfile

import pytest

from momentum_b200 import solver as ms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <momentum_b200_adapters.hpp>
int main() {
  using namespace momentum_b200;
  // createTestCharacter(3): Y chain, 9 + 1 model parameters (character_helpers.cpp:106-149)
  std::vector<int32_t> parents{-1, 0, 1};
  std::vector<float> off{0,0,0, 0,1,0, 0,1,0}, pre{0,0,0,1, 0,0,0,1, 0,0,0,1};
  std::vector<int32_t> outer(22, 0), inner; std::vector<float> vals, offs(21, 0.f);
  int rows[] = {0,1,2,3,4,5,6, 10, 12, 19, 17};
  int cols[] = {0,1,2,3,4,5,6, 7, 8, 8, 9};
  float cf[] = {1,1,1,1,1,1,1, 1, .5f, .5f, 1};
  for (int r = 0; r < 21; ++r) { for (int k = 0; k < 11; ++k) if (rows[k] == r) { inner.push_back(cols[k]); vals.push_back(cf[k]); } outer[r + 1] = int(inner.size()); }
  try {
    Character ch(0, parents, off, pre, 10, outer, inner, vals, offs);
    BatchedSkeletonSolverFunction fn(ch, 4);
    GaussNewtonSolverOptions o; o.maxIterations = 6; o.minIterations = 6; o.regularization = 1e-7f; o.useBlockJtJ = true;
    int pos = fn.addPositionErrorFunction(1.f, {2}, {0.f, 1.f, 0.f}, {1.f});
    int pl = fn.addPlaneErrorFunction(1e-4f, {2}, {0.f, 0.5f, 0.f}, {1.f}, /*above=*/true);
    int mp = fn.addModelParametersErrorFunction(1e-3f, std::vector<float>(10, 1.f));
    // every block first, then the per-instance data (the record layout is fixed by the block list)
    std::vector<float> tg(4 * 3, 0.f); for (int b = 0; b < 4; ++b) { tg[3*b] = 0.3f * b; tg[3*b+1] = 2.5f; }
    fn.setTargets(pos, tg);
    std::vector<float> planes(4 * 4, 0.f); for (int b = 0; b < 4; ++b) { planes[4*b+1] = 1.f; planes[4*b+3] = -1.f; }
    fn.setTargets(pl, planes);
    fn.setTargets(mp, std::vector<float>(4 * 10, 0.f));
    BatchedGaussNewtonSolver solver(o, &fn);
    std::vector<float> params(4 * 10, 0.f);
    auto r = solver.solve(params);
    std::printf("solved: err0=%g it0=%d\n", r.errors[0], r.iterations[0]);
  } catch (const std::runtime_error& e) {
    std::printf("runtime_error: %s\n", e.what());
    return 3;
  }
  return 0;
}
"""


def _build_and_run():
    import __graft_entry__ as g

    g.build()
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        exe = os.path.join(d, "t")
        open(src, "w").write(SRC)
        lib_dir = os.path.dirname(ms.DEFAULT_LIB)
        subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", lib_dir, "-lmomentum_b200",
                               f"-Wl,-rpath,{lib_dir}"])
        return subprocess.run([exe], capture_output=True, text=True)


def test_adapter_header_compiles_and_fails_loudly_without_gpu():
    p = _build_and_run()
    if ms.load_library().mb2_device_count() > 0:
        assert p.returncode == 0 and "solved" in p.stdout, p.stdout + p.stderr
    else:
        assert p.returncode == 3 and "no usable sm_100 CUDA device" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_adapter_part1_solves_on_the_gpu():
    """The C++ host side (Part 1 of the adapters header) drives a real solve on the B200 under `pytest -m gpu`."""
    p = _build_and_run()
    assert p.returncode == 0 and "solved" in p.stdout, p.stdout + p.stderr
