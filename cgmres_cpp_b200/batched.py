"""Host-side mirror of the reference's `Cgmres<Model>` (include/cgmres.hpp:8-207), batched.

Same method names, argument meaning and side effects as the reference class, with a
leading instance dimension on every array:

    reference (one controller)                 here (n controllers on one B200)
    -----------------------------------------  -------------------------------------------
    Cgmres<Model> c;                           c = BatchedCgmres(MSD, n, device=0)
    c.set_ptau(pt)      pt[(dv+1)*dim_p]       c.set_ptau(pt)       pt[n][(dv+1)*dim_p]
    c.set_ptau_repeat(p)                       c.set_ptau_repeat(p) p[n][dim_p]
    c.init_u0(u)                               c.init_u0(u)         u[n][dim_u] (or [dim_u])
    c.init_u0_newton(u, x, p, 10)  mutates u   u = c.init_u0_newton(u, x, p, 10)  returns refined u
    c.control(u, x)                            u = c.control(x)     x[n][dim_x] -> u[n][dim_u]
    main.cpp loop: control + Euler plant step  c.step_closed_loop(k)  state stays in HBM

Everything numerical happens in libcgmres_b200.so (hand-written sm_100a kernels) through
the C ABI of include/cgmres_b200.h; numpy here only owns host buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import CgmresB200Error, check, lib

MSD, ARM, SEMIACTIVE = 0, 1, 2
MODE_EXACT, MODE_FAST, MODE_ONCHIP_EXACT, MODE_PIPELINED_EXACT = 0, 1, 2, 3
EXIT_FULL, EXIT_CONVERGED, EXIT_RHO0, EXIT_BREAKDOWN = 0, 1, 2, 3


@dataclass(frozen=True)
class ModelDims:
    dim_x: int
    dim_u: int
    dim_p: int
    dv: int
    k_max: int
    control_input: int

    @property
    def L(self) -> int:
        return self.dim_u * self.dv


def model_dims(model: int) -> ModelDims:
    d = (C.c_int * 6)()
    check(lib().cgmres_b200_model_dims(model, d))
    return ModelDims(*[int(v) for v in d])


def model_params(model: int) -> dict:
    p = (C.c_double * 6)()
    check(lib().cgmres_b200_model_params(model, p))
    return dict(zip(("dt", "h", "zeta", "Tf", "alpha", "tol"), [float(v) for v in p]))


def model_name(model: int) -> str:
    s = lib().cgmres_b200_model_name(model)
    if s is None:
        raise CgmresB200Error(f"unknown model {model}")
    return s.decode()


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class BatchedCgmres:
    """n independent C/GMRES controllers of one model, resident on one GPU."""

    def __init__(self, model: int, n: int, device: int = 0, mode: int = MODE_EXACT):
        self._h = None
        self.dims = model_dims(model)
        self.params = model_params(model)
        self.model, self.n, self.device, self.mode = model, int(n), device, mode
        h = C.c_void_p()
        check(lib().cgmres_b200_create(model, self.n, device, mode, C.byref(h)))
        self._h = h
        # the reference's public constants (cgmres.hpp:179-188)
        self.dim_x, self.dim_u, self.dim_p, self.dv = self.dims.dim_x, self.dims.dim_u, self.dims.dim_p, self.dims.dv
        self.dt, self.h, self.zeta = self.params["dt"], self.params["h"], self.params["zeta"]
        self.Tf, self.alpha = self.params["Tf"], self.params["alpha"]

    # -- lifetime ------------------------------------------------------------
    def close(self):
        if self._h is not None:
            lib().cgmres_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- reference methods -----------------------------------------------------
    def get_dtau(self, t: float) -> float:
        return float(lib().cgmres_b200_get_dtau(self._h, float(t)))

    def set_ptau(self, ptau):
        a = _f64(ptau, (self.n, (self.dv + 1) * self.dim_p))
        check(lib().cgmres_b200_set_ptau(self._h, _ptr(a)))

    def set_ptau_repeat(self, p):
        a = _f64(p, (self.n, self.dim_p))
        check(lib().cgmres_b200_set_ptau_repeat(self._h, _ptr(a)))

    def init_u0(self, u0):
        a = _f64(u0, (self.n, self.dim_u))
        check(lib().cgmres_b200_init_u0(self._h, _ptr(a)))

    def init_u0_newton(self, u0, x0, p0=None, n_loop: int = 10):
        u = _f64(u0, (self.n, self.dim_u)).copy()
        x = _f64(x0, (self.n, self.dim_x))
        p = _f64(p0 if self.dim_p else np.zeros((self.n, 0)), (self.n, self.dim_p))
        check(lib().cgmres_b200_init_u0_newton(self._h, _ptr(u), _ptr(x), _ptr(p), int(n_loop)))
        return u

    def control(self, x, out=None):
        x = _f64(x, (self.n, self.dim_x))
        if out is None:
            u = np.empty((self.n, self.dim_u))
        else:  # the C ABI writes n*dim_u doubles straight into this buffer
            u = out
            if not (isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous
                    and u.flags.writeable and u.shape == (self.n, self.dim_u)):
                raise ValueError(f"out must be a writable C-contiguous float64 array of shape ({self.n}, {self.dim_u})")
        check(lib().cgmres_b200_control(self._h, _ptr(u), _ptr(x)))
        return u

    def control_raw(self, u_ptr: int, x_ptr: int):
        """control() on caller-owned host buffers given as raw addresses (e.g. pinned torch tensors)."""
        check(lib().cgmres_b200_control(self._h, C.c_void_p(u_ptr), C.c_void_p(x_ptr)))

    def control_dev(self, u_dev_ptr: int, x_dev_ptr: int):
        check(lib().cgmres_b200_control_dev(self._h, C.c_void_p(u_dev_ptr), C.c_void_p(x_dev_ptr)))

    # -- device-resident closed loop -------------------------------------------
    def set_x(self, x):
        a = _f64(x, (self.n, self.dim_x))
        check(lib().cgmres_b200_set_x(self._h, _ptr(a)))

    def get_x(self):
        x = np.empty((self.n, self.dim_x))
        check(lib().cgmres_b200_get_x(self._h, _ptr(x)))
        return x

    def get_u(self):
        u = np.empty((self.n, self.dim_u))
        check(lib().cgmres_b200_get_u(self._h, _ptr(u)))
        return u

    def step_closed_loop(self, n_steps: int = 1):
        check(lib().cgmres_b200_step_closed_loop(self._h, int(n_steps)))

    def step_closed_loop_log(self, n_steps: int):
        """n_steps closed-loop steps with the trajectory recorded on the device: returns
        (x_log[n_steps][n][dim_x], u_log[n_steps][n][dim_u]) -- the rows the reference's mains print per step."""
        xl = np.empty((int(n_steps), self.n, self.dim_x))
        ul = np.empty((int(n_steps), self.n, self.dim_u))
        check(lib().cgmres_b200_step_closed_loop_log(self._h, int(n_steps), _ptr(xl), _ptr(ul)))
        return xl, ul

    def set_t(self, t):
        """Per-instance controller clocks t[n] (None: back to the batch-uniform clock)."""
        a = None if t is None else _f64(t, (self.n,))
        check(lib().cgmres_b200_set_t(self._h, _ptr(a)))

    def get_t(self):
        t = np.empty(self.n)
        check(lib().cgmres_b200_get_t(self._h, _ptr(t)))
        return t

    def set_plant_integrator(self, name: str):
        """'euler' (the reference's plant step, parity default) or 'rk4' (an extension: the reference has no RK4 to compare with)."""
        check(lib().cgmres_b200_set_plant_integrator(self._h, {"euler": 1, "rk4": 2}[name]))

    def synchronize(self):
        check(lib().cgmres_b200_synchronize(self._h))

    def set_stream(self, cuda_stream: int | None):
        check(lib().cgmres_b200_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # -- checkpoint / diagnostics -------------------------------------------------
    def get_state(self, want_U=True, want_dUdt=True):
        t = C.c_double()
        U = np.empty((self.n, self.dims.L)) if want_U else None
        dUdt = np.empty((self.n, self.dims.L)) if want_dUdt else None
        check(lib().cgmres_b200_get_state(self._h, C.byref(t), _ptr(U), _ptr(dUdt)))
        return float(t.value), U, dUdt

    def set_state(self, t=None, U=None, dUdt=None):
        tt = None if t is None else C.byref(C.c_double(float(t)))
        Ua = None if U is None else _f64(U, (self.n, self.dims.L))
        da = None if dUdt is None else _f64(dUdt, (self.n, self.dims.L))
        check(lib().cgmres_b200_set_state(self._h, tt, _ptr(Ua), _ptr(da)))

    def get_status(self):
        """(exit_code[n], columns_used[n]) of the last update."""
        s = np.empty(self.n, dtype=np.int32)
        check(lib().cgmres_b200_get_status(self._h, _ptr(s)))
        return s & 0xFF, s >> 8


def plant_step_host(model: int, x: np.ndarray, u: np.ndarray) -> None:
    """x[n][dim_x] += Simulator::dxdt(x,u)*dt in place on the host (the plant step of the reference's main loop)."""
    assert x.dtype == np.float64 and u.dtype == np.float64 and x.flags.c_contiguous and u.flags.c_contiguous
    check(lib().cgmres_b200_plant_step_host(model, x.shape[0], C.c_void_p(x.ctypes.data), C.c_void_p(u.ctypes.data)))


def launch_count() -> int:
    return int(lib().cgmres_b200_launch_count())


def device_count() -> int:
    return int(lib().cgmres_b200_device_count())


def measure_fp64_peak(device: int = 0, use_fma: bool = True) -> float:
    """Measured FP64 vector peak in TFLOP/s (DFMA = 2 flop; use_fma=False: separate DMUL+DADD)."""
    v = C.c_double()
    check(lib().cgmres_b200_measure_fp64_peak(device, int(use_fma), C.byref(v), None))
    return float(v.value)


def measure_fp64_latency(device: int = 0) -> dict:
    """Dependent-issue latency of DFMA / DADD / DMUL in SM cycles."""
    out = {}
    for op, name in enumerate(("dfma", "dadd", "dmul")):
        v = C.c_double()
        check(lib().cgmres_b200_measure_fp64_latency(device, op, C.byref(v)))
        out[name] = float(v.value)
    return out
