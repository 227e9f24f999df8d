"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcgmres_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container only; the fixtures are committed.

    python tests/golden/make_golden.py

Per model (the reference's known-answer programs are its four example mains):
  shipped_*   the shipped initial condition of <example>/main.cpp, x and u after steps 100,200,...,2000,
              plus the full final state of the shipped run length (20001 / 10001 / 20001 steps)
  batch_*     8 seeded synthetic instances (SURVEY.md 8d distributions), 1000 steps, x/u every 100 steps and
              the complete controller state {U, dUdt} after 1000 steps
  tf_*        teacher-forcing snapshots of the shipped run: state {t,U,dUdt,x} before step s and U,dUdt,u
              after it, for s in (0, 1, 2, 10, 100, 999)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TF_STEPS = (0, 1, 2, 10, 100, 999)


def main():
    ref = po.load("reference")
    for model in (po.MSD, po.ARM, po.SEMIACTIVE):
        s = po.SHIPPED[model]
        dm = ref.dims(model)
        p = [s["p"]] if s["p"] else None
        g = {}
        a = ref.run_closed_loop(model, [s["x0"]], p, s["u0"], 2000, rec_stride=100)
        g["shipped_x_traj"], g["shipped_u_traj"] = a["x_traj"][:, 0], a["u_traj"][:, 0]
        a = ref.run_closed_loop(model, [s["x0"]], p, s["u0"], s["steps"], want_U=True)
        g["shipped_x_fin"], g["shipped_u_fin"] = a["x_fin"][0], a["u_fin"][0]
        g["shipped_U_fin"], g["shipped_dUdt_fin"] = a["U_fin"][0], a["dUdt_fin"][0]
        g["shipped_steps"] = np.array(s["steps"])

        x0, pb, u0 = po.synthetic_batch(model, 8)
        a = ref.run_closed_loop(model, x0, pb, u0, 1000, rec_stride=100, want_U=True)
        g["batch_x0"], g["batch_p"], g["batch_u0"] = x0, pb, u0
        g["batch_x_traj"], g["batch_u_traj"] = a["x_traj"], a["u_traj"]
        g["batch_U_fin"], g["batch_dUdt_fin"] = a["U_fin"], a["dUdt_fin"]

        # teacher forcing snapshots along the shipped run
        c = ref.controller(model)
        x = np.array(s["x0"], dtype=np.float64)
        if dm.dim_p:
            c.set_ptau_repeat(s["p"])
        c.init_u0(s["u0"])
        u_newton = c.init_u0_newton(s["u0"], x, s["p"] if dm.dim_p else [0.0], 10)
        g["newton_u0"] = u_newton
        snaps = {k: [] for k in ("t", "U", "dUdt", "x", "U_after", "dUdt_after", "u_after")}
        for step in range(max(TF_STEPS) + 1):
            if step in TF_STEPS:
                t, U, dUdt = c.get_state()
                snaps["t"].append(t), snaps["U"].append(U), snaps["dUdt"].append(dUdt), snaps["x"].append(x.copy())
            u = c.control(x)
            if step in TF_STEPS:
                _, U, dUdt = c.get_state()
                snaps["U_after"].append(U), snaps["dUdt_after"].append(dUdt), snaps["u_after"].append(u)
            ref.plant_step(model, x, u)
        for k, v in snaps.items():
            g["tf_" + k] = np.array(v)
        g["tf_steps"] = np.array(TF_STEPS)
        path = os.path.join(OUT, f"{po.MODEL_NAMES[model]}.npz")
        np.savez_compressed(path, **g)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
