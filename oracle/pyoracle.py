"""ctypes front-end to the two CPU oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/cgmres_oracle.h).  The product package
cgmres_cpp_b200 never does.

    port = load("port")        # oracle/_build/libcgmres_oracle.so  (C restatement)
    pt   = load("port_ptrig")  # the same with the arm model's sin/cos from oracle/portable_trig.h
    ref  = load("reference")   # oracle/_ref/libcgmres_ref.so        (unmodified reference headers)

Both expose the same methods; arrays use the reference's own layouts
(x[n][dim_x], u[n][dim_u], U[n][dv*dim_u], ptau[n][(dv+1)*dim_p]).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MSD, ARM, SEMIACTIVE = 0, 1, 2
MODEL_NAMES = {MSD: "mass_spring_damper", ARM: "arm_type_inverted_pendulum", SEMIACTIVE: "semiactive_damper"}
EXIT_NAMES = ("full", "converged", "rho0_below_tol", "breakdown")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


@dataclass(frozen=True)
class Dims:
    dim_x: int
    dim_u: int
    dim_p: int
    dv: int
    k_max: int
    n_ctrl: int

    @property
    def L(self) -> int:
        return self.dim_u * self.dv


class Oracle:
    """One oracle library (kind 'port' or 'reference')."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.kind = kind
        self.path = path
        self._lib = C.CDLL(path)
        self._p = prefix
        f = self._fn
        f("model_dims", C.c_int, [C.c_int, C.POINTER(C.c_int)])
        f("model_params", C.c_int, [C.c_int, _dp])
        f("create", C.c_void_p, [C.c_int])
        f("destroy", None, [C.c_void_p])
        f("set_ptau", None, [C.c_void_p, _dp])
        f("set_ptau_repeat", None, [C.c_void_p, _dp])
        f("init_u0", None, [C.c_void_p, _dp])
        f("init_u0_newton", None, [C.c_void_p, _dp, _dp, _dp, C.c_int])
        f("control", None, [C.c_void_p, _dp, _dp])
        f("get_dtau", C.c_double, [C.c_void_p, C.c_double])
        f("get_state", None, [C.c_void_p, _dp, _dp, _dp])
        f("set_state", None, [C.c_void_p, _dp, _dp, _dp])
        f("last_status", C.c_int, [C.c_void_p])
        f("plant_step", None, [C.c_int, _dp, _dp])
        f("run_closed_loop", C.c_int,
          [C.c_int, C.c_int64, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int,
           _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, C.c_int])
        f("last_loop_seconds", C.c_double, [])
        if kind != "reference":
            f("sincos", None, [C.c_double, _dp, _dp])

    def sincos(self, x: float):
        """(sin x, cos x) as this oracle build evaluates them inside the arm model."""
        s, c = C.c_double(), C.c_double()
        self._sincos(float(x), C.byref(s), C.byref(c))
        return s.value, c.value

    def _fn(self, name, restype, argtypes):
        fn = getattr(self._lib, self._p + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, "_" + name, fn)

    # ---- static info -------------------------------------------------------
    def dims(self, model: int) -> Dims:
        d = (C.c_int * 6)()
        if self._model_dims(model, d) != 0:
            raise ValueError(f"unknown model {model}")
        return Dims(*[int(v) for v in d])

    def params(self, model: int) -> dict:
        p = np.zeros(6)
        self._model_params(model, _d(p))
        return dict(zip(("dt", "h", "zeta", "Tf", "alpha", "tol"), p.tolist()))

    # ---- single controller -------------------------------------------------
    def controller(self, model: int) -> "OracleController":
        return OracleController(self, model)

    def plant_step(self, model: int, x: np.ndarray, u: np.ndarray) -> None:
        """x <- x + Simulator::dxdt(x,u)*dt in place (1-D arrays)."""
        self._plant_step(model, _d(x), _d(u))

    # ---- batch closed loop -------------------------------------------------
    def run_closed_loop(self, model, x0, p, u0, n_steps, *, p_full=False, newton_iters=10,
                        rec_stride=0, n_threads=1, want_U=False, want_traj=True):
        """Runs n instances for n_steps closed-loop steps.  Returns a dict of arrays."""
        dm = self.dims(model)
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, dm.dim_x)
        n = x0.shape[0]
        u0 = np.ascontiguousarray(np.broadcast_to(np.asarray(u0, dtype=np.float64), (n, dm.dim_u)))
        if dm.dim_p > 0:
            plen = dm.dim_p * (dm.dv + 1) if p_full else dm.dim_p
            p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64), (n, plen)))
        else:
            p = None
        nrec = n_steps // rec_stride if rec_stride > 0 else 0
        out = {
            "x_fin": np.zeros((n, dm.dim_x)), "u_fin": np.zeros((n, dm.dim_u)),
            "exit_hist": np.zeros((n, 4), dtype=np.int32), "ctl_seconds": np.zeros(n),
        }
        if nrec and want_traj:
            out["x_traj"] = np.zeros((nrec, n, dm.dim_x))
            out["u_traj"] = np.zeros((nrec, n, dm.dim_u))
        if want_U:
            out["U_fin"] = np.zeros((n, dm.L))
            out["dUdt_fin"] = np.zeros((n, dm.L))
        rc = self._run_closed_loop(
            model, n, _d(x0), _d(p), int(p_full), _d(u0), newton_iters, n_steps, rec_stride,
            _d(out.get("x_traj")), _d(out.get("u_traj")), _d(out["x_fin"]), _d(out["u_fin"]),
            _d(out.get("U_fin")), _d(out.get("dUdt_fin")),
            out["exit_hist"].ctypes.data_as(_ip), _d(out["ctl_seconds"]), n_threads)
        if rc != 0:
            raise RuntimeError("run_closed_loop: bad arguments")
        # wall time of the step loops alone (max over the worker threads): what a throughput figure should use
        out["loop_seconds"] = float(self._last_loop_seconds())
        return out


class OracleController:
    """Mirror of one reference `Cgmres<Model>` object (include/cgmres.hpp:8-207)."""

    def __init__(self, lib: Oracle, model: int):
        self.lib, self.model = lib, model
        self.dims = lib.dims(model)
        self._h = lib._create(model)
        if not self._h:
            raise ValueError(f"unknown model {model}")

    def close(self):
        if self._h:
            self.lib._destroy(self._h)
            self._h = None

    __del__ = close

    def set_ptau(self, ptau):
        a = np.ascontiguousarray(ptau, dtype=np.float64)
        assert a.size == self.dims.dim_p * (self.dims.dv + 1)
        self.lib._set_ptau(self._h, _d(a))

    def set_ptau_repeat(self, p):
        a = np.ascontiguousarray(p, dtype=np.float64)
        assert a.size == self.dims.dim_p
        self.lib._set_ptau_repeat(self._h, _d(a))

    def init_u0(self, u0):
        a = np.ascontiguousarray(u0, dtype=np.float64)
        self.lib._init_u0(self._h, _d(a))

    def init_u0_newton(self, u0, x0, p0, n_loop=10):
        """Mutates and returns u0 like the reference (include/cgmres.hpp:61-76)."""
        u0 = np.array(u0, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        p0 = np.ascontiguousarray(p0 if self.dims.dim_p else np.zeros(1), dtype=np.float64)
        self.lib._init_u0_newton(self._h, _d(u0), _d(x0), _d(p0), n_loop)
        return u0

    def control(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.zeros(self.dims.dim_u)
        self.lib._control(self._h, _d(u), _d(x))
        return u

    def get_dtau(self, t):
        return float(self.lib._get_dtau(self._h, t))

    def get_state(self):
        t = np.zeros(1)
        U = np.zeros(self.dims.L)
        dUdt = np.zeros(self.dims.L)
        self.lib._get_state(self._h, _d(t), _d(U), _d(dUdt))
        return float(t[0]), U, dUdt

    def set_state(self, t=None, U=None, dUdt=None):
        ta = None if t is None else np.array([t], dtype=np.float64)
        Ua = None if U is None else np.ascontiguousarray(U, dtype=np.float64)
        da = None if dUdt is None else np.ascontiguousarray(dUdt, dtype=np.float64)
        self.lib._set_state(self._h, _d(ta), _d(Ua), _d(da))

    def last_status(self):
        """(exit_code, columns_used) of the last control(); (-1, -1) for the reference library."""
        s = int(self.lib._last_status(self._h))
        return (-1, -1) if s < 0 else (s & 0xFF, s >> 8)


_PATHS = {
    "port": (os.path.join(HERE, "_build", "libcgmres_oracle.so"), "oracle_"),
    "port_ptrig": (os.path.join(HERE, "_build", "libcgmres_oracle_ptrig.so"), "oraclept_"),
    "reference": (os.path.join(HERE, "_ref", "libcgmres_ref.so"), "ref_"),
}
_CACHE: dict = {}


def build(quiet: bool = True) -> None:
    """(Re)build the oracles with oracle/Makefile; the reference library only where its sources exist."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=None)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind][0])


def load(kind: str = "port") -> Oracle:
    if kind not in _CACHE:
        path, prefix = _PATHS[kind]
        if not os.path.exists(path):
            if kind in ("port", "port_ptrig"):
                build()
            else:
                raise FileNotFoundError(f"{path} missing: run `make -C oracle` where /root/reference exists")
        _CACHE[kind] = Oracle(path, prefix, kind)
    return _CACHE[kind]


def best() -> Oracle:
    """The strongest oracle present: the compiled reference if it travelled, else the C port."""
    return load("reference") if available("reference") else load("port")


# ---- workloads: the shipped initial conditions and the seeded synthetic batches live in the neutral module
# cgmres_cpp_b200/workloads.py (pure numpy; bench.py's measured arm must not import anything from oracle/) ----
from cgmres_cpp_b200.workloads import SHIPPED, synthetic_batch  # noqa: E402,F401
