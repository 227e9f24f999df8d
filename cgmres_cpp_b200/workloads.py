"""Workload definitions shared by the benchmark, the tests and the profiling drivers: the reference's shipped
initial conditions and the seeded synthetic batches of SURVEY.md section 8(d).  Pure numpy, no dependency on the
library or on the test infrastructure."""
from __future__ import annotations

import numpy as np

MSD, ARM, SEMIACTIVE = 0, 1, 2

# the shipped initial conditions (known-answer programs, <example>/main.cpp)
SHIPPED = {
    # mass_spring_damper/main.cpp:35-55
    MSD: dict(x0=[2.0, 2.0, 0.0, 0.0], u0=[0.0, 0.0, 10.0, 10.0, 5e-4, 5e-4], p=[1.0, -1.0], steps=20001),
    # arm_type_inverted_pendulum/main.cpp:35-52
    ARM: dict(x0=[3.14159265358979, 3.14159265358979, 0.0, 0.0], u0=[0.0, 3.0, 0.01],
              p=[3.14159265358979 / 4.0, 0.0], steps=10001),
    # semiactive_damper/main.cpp:35-40
    SEMIACTIVE: dict(x0=[2.0, 0.0], u0=[0.028393761456740, 0.166095020295846, 0.030103250483332], p=[], steps=20001),
}


def synthetic_batch(model: int, n: int, seed: int = 12345):
    """Seeded synthetic batch of SURVEY.md section 8(d): returns x0[n][dim_x], p[n][dim_p], u0[dim_u].

    The distributions are the ones the survey measured as well conditioned
    (no breakdowns, closed-loop FMA/reordering drift below 1e-6).
    """
    rng = np.random.Generator(np.random.PCG64(seed + 1000003 * model))
    if model == MSD:
        x0 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n),
                       rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n)], axis=1)
        p = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)], axis=1)
    elif model == ARM:
        s, pi = 0.05, 3.14159265358979
        x0 = np.stack([pi + rng.uniform(-s, s, n), pi + rng.uniform(-s, s, n),
                       rng.uniform(-s / 2, s / 2, n), rng.uniform(-s / 2, s / 2, n)], axis=1)
        p = np.stack([pi / 4.0 + rng.uniform(-s, s, n), np.zeros(n)], axis=1)
    elif model == SEMIACTIVE:
        x0 = np.stack([rng.uniform(1.5, 2.5, n), rng.uniform(-0.1, 0.1, n)], axis=1)
        p = np.zeros((n, 0))
    else:
        raise ValueError(model)
    return np.ascontiguousarray(x0), np.ascontiguousarray(p), np.array(SHIPPED[model]["u0"], dtype=np.float64)
