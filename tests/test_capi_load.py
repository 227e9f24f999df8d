"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what include/cgmres_b200.h
declares, reports the reference's problem sizes, and fails loudly (never falls back) without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cgmres_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgmres_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_table_agree(built):
    from cgmres_cpp_b200._lib import SIGNATURES

    assert declared_symbols() == sorted(SIGNATURES)


def test_header_enums_match_the_python_mirror(built):
    """Model ids, build modes, exit codes and plant integrators: the numbers in include/cgmres_b200.h are the
    numbers cgmres_cpp_b200 uses."""
    import cgmres_cpp_b200 as cg

    src = open(os.path.join(ROOT, "include", "cgmres_b200.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(CGMRES_B200_[A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    assert enums["CGMRES_B200_MODEL_MASS_SPRING_DAMPER"] == cg.MSD
    assert enums["CGMRES_B200_MODEL_ARM_TYPE_INVERTED_PENDULUM"] == cg.ARM
    assert enums["CGMRES_B200_MODEL_SEMIACTIVE_DAMPER"] == cg.SEMIACTIVE
    assert enums["CGMRES_B200_MODE_EXACT"] == cg.MODE_EXACT
    assert enums["CGMRES_B200_MODE_FAST"] == cg.MODE_FAST
    assert enums["CGMRES_B200_MODE_ONCHIP_EXACT"] == cg.MODE_ONCHIP_EXACT
    assert enums["CGMRES_B200_MODE_PIPELINED_EXACT"] == cg.MODE_PIPELINED_EXACT
    assert (enums["CGMRES_B200_EXIT_FULL"], enums["CGMRES_B200_EXIT_CONVERGED"], enums["CGMRES_B200_EXIT_RHO0"],
            enums["CGMRES_B200_EXIT_BREAKDOWN"]) == (cg.EXIT_FULL, cg.EXIT_CONVERGED, cg.EXIT_RHO0, cg.EXIT_BREAKDOWN)


def test_library_exports_every_declared_symbol(built):
    from cgmres_cpp_b200._lib import LIB_PATH

    l = ctypes.CDLL(LIB_PATH)
    for name in declared_symbols():
        assert hasattr(l, name), name


def test_model_tables_match_reference(built, oracle_port):
    import cgmres_cpp_b200 as cg

    for m in (cg.MSD, cg.ARM, cg.SEMIACTIVE):
        d, o = cg.model_dims(m), oracle_port.dims(m)
        assert (d.dim_x, d.dim_u, d.dim_p, d.dv, d.k_max, d.control_input) == (o.dim_x, o.dim_u, o.dim_p, o.dv,
                                                                               o.k_max, o.n_ctrl)
        assert cg.model_params(m) == oracle_port.params(m)
    assert cg.model_name(cg.MSD) == "mass_spring_damper"
    with pytest.raises(cg.CgmresB200Error):
        cg.model_dims(17)


def test_no_cpu_fallback(built):
    import cgmres_cpp_b200 as cg

    if cg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cg.CgmresB200Error, match="no usable CUDA device"):
        cg.BatchedCgmres(cg.MSD, 8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cgmres_cpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, (dirpath, f)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), (dirpath, f)


def test_cpp_dropin_header_compiles_with_host_compiler(built, tmp_path):
    """include/cgmres.hpp + the per-example model.hpp/simulator.hpp names compile as plain C++ (no nvcc) and link
    against the C-ABI library: what a maintainer of the reference's main.cpp would build."""
    import subprocess

    src = tmp_path / "main.cpp"
    src.write_text("""
#include "cgmres.hpp"
#include "mass_spring_damper/model.hpp"
#include "mass_spring_damper/simulator.hpp"
#include "multiple_controller/model2.hpp"
#include "multiple_controller/simulator1.hpp"
static_assert(Cgmres<Model>::dim_x == 4 && Cgmres<Model>::dim_u == 6 && Cgmres<Model>::dv == 50, "msd dims");
static_assert(Cgmres<Model2>::dim_u == 3 && Cgmres<Model2>::dv == 25, "arm dims");
static_assert(Simulator::t_end == 20 && Simulator1::t_end == 10, "t_end");
int main() {
  double x[4] = {2, 2, 0, 0}, u[6] = {0, 0, 10, 10, 5e-4, 5e-4}, d[4];
  Simulator::dxdt(d, x, u);              // host-side use of the same functor the kernels inline
  if (d[2] != -2.0 + 2.0) return 2;      // -(k1*k2)/m1*x0 + k2/m1*x1 = -2 + 2
  try { Cgmres<Model> c; c.control(u, x); } catch (const std::exception&) { return 0; }  // no GPU here: must throw
  return 0;
}
""")
    exe = tmp_path / "main"
    lib_dir = os.path.join(ROOT, "cgmres_cpp_b200")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-L", lib_dir,
                        "-lcgmres_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_host_plant_step_matches_reference_plant(built, oracle_port):
    """cgmres_b200_plant_step_host is plain host code (the Simulator functor the kernels also inline): it must
    reproduce the reference's x += Simulator::dxdt(x,u)*dt bit for bit (<example>/main.cpp:74-76)."""
    import numpy as np

    import cgmres_cpp_b200 as cg
    from oracle import pyoracle as po

    rng = np.random.default_rng(7)
    for m in (cg.MSD, cg.SEMIACTIVE, cg.ARM):
        d = cg.model_dims(m)
        x = rng.normal(size=(50, d.dim_x))
        u = rng.normal(size=(50, d.dim_u))
        want = x.copy()
        for i in range(50):
            oracle_port.plant_step(m, want[i], u[i])
        got = x.copy()
        cg.plant_step_host(m, got, u)
        assert np.array_equal(got, want), po.MODEL_NAMES[m]


def test_portable_trig_host_equals_oracle_restatement_and_is_within_1ulp_of_glibc(built):
    """The arm model's sin/cos (include/cgmres_b200/portable_trig.hpp, compiled for the host inside the library) must
    equal the oracle's separate C restatement (oracle/portable_trig.h) bit for bit, and both must stay within
    1 ulp of glibc -- the only difference between the GPU's exact modes and the reference on the arm model."""
    import ctypes as C

    import numpy as np

    from cgmres_cpp_b200._lib import lib
    from oracle import pyoracle as po

    pt, gl = po.load("port_ptrig"), po.load("port")
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-0.8, 0.8, 4000), rng.uniform(-7, 7, 4000), rng.uniform(-400, 400, 4000),
                         rng.uniform(-8e5, 8e5, 2000), [0.0, 3.14159265358979, 0.785398163397448, -0.785398163397449]])
    worst = 0
    for x in xs:
        s, c = C.c_double(), C.c_double()
        lib().cgmres_b200_portable_sincos(float(x), C.byref(s), C.byref(c))
        assert (s.value, c.value) == pt.sincos(x), x
        gs, gc = gl.sincos(x)
        for a, b in ((s.value, gs), (c.value, gc)):
            worst = max(worst, abs(int(np.float64(a).view(np.int64)) - int(np.float64(b).view(np.int64))))
    assert worst <= 1, worst
