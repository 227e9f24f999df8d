// ref_harness.cpp -- compiles the UNMODIFIED reference headers (found at build
// time under $(REF)=/root/reference, never copied into this repository) into
// oracle/_ref/libcgmres_ref.so, exporting the interface of cgmres_oracle.h with
// the prefix "ref_".
//
// TEST INFRASTRUCTURE ONLY (see cgmres_oracle.h).  It is the strongest oracle we
// have: the reference's own Cgmres<Model>::control (include/cgmres.hpp:78-110),
// its own models and its own plant equations.  Used to (a) pin the C restatement,
// (b) generate tests/golden/, (c) serve as the "reference" CPU baseline in bench.py.
//
// Flags: g++ -O3 -Wall (the reference's CMakeLists.txt:40), no -march=native, no
// -ffast-math.
//
// Two deliberate harness-side choices, both outside the reference sources:
//  * `#define private public` around the reference includes so that t / U / dUdt
//    (include/cgmres.hpp:195-197) can be read and written for checkpoints and
//    teacher-forced tests.
//  * dUdt is zeroed right after construction: the reference never initialises it
//    (include/cgmres.hpp:14) and only works on a fresh, zero heap.
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <thread>
#include <vector>

#define private public
#define protected public
#include "cgmres.hpp"
namespace ref_msd {
#include "mass_spring_damper/model.hpp"
#include "mass_spring_damper/simulator.hpp"
}  // namespace ref_msd
namespace ref_arm {
#include "arm_type_inverted_pendulum/model.hpp"
#include "arm_type_inverted_pendulum/simulator.hpp"
}  // namespace ref_arm
namespace ref_sad {
#include "semiactive_damper/model.hpp"
#include "semiactive_damper/simulator.hpp"
}  // namespace ref_sad
#undef private
#undef protected

#define ORACLE_PREFIX ref_
#include "cgmres_oracle.h"

namespace {

struct CtlBase {
  virtual ~CtlBase() {}
  virtual void set_ptau(const double*) = 0;
  virtual void set_ptau_repeat(const double*) = 0;
  virtual void init_u0(const double*) = 0;
  virtual void init_u0_newton(double*, const double*, const double*, int) = 0;
  virtual void control(double*, const double*) = 0;
  virtual double get_dtau(double) = 0;
  virtual void get_state(double*, double*, double*) = 0;
  virtual void set_state(const double*, const double*, const double*) = 0;
};

template <class Model>
struct Ctl : CtlBase {
  Cgmres<Model> c;
  static constexpr int L = Model::dim_u * Model::dv;
  Ctl() { memset(c.dUdt, 0, sizeof(double) * L); }
  void set_ptau(const double* p) override { c.set_ptau(p); }
  void set_ptau_repeat(const double* p) override { c.set_ptau_repeat(p); }
  void init_u0(const double* u) override { c.init_u0(u); }
  void init_u0_newton(double* u, const double* x, const double* p, int n) override {
    c.init_u0_newton(u, x, p, (uint16_t)n);
  }
  void control(double* u, const double* x) override { c.control(u, x); }
  double get_dtau(double t) override { return c.get_dtau(t); }
  void get_state(double* t, double* U, double* dUdt) override {
    if (t) *t = c.t;
    if (U) memcpy(U, c.U, sizeof(double) * L);
    if (dUdt) memcpy(dUdt, c.dUdt, sizeof(double) * L);
  }
  void set_state(const double* t, const double* U, const double* dUdt) override {
    if (t) c.t = *t;
    if (U) memcpy(c.U, U, sizeof(double) * L);
    if (dUdt) memcpy(c.dUdt, dUdt, sizeof(double) * L);
  }
};

// x = x + dxdt*dt exactly as <example>/main.cpp:74-76 writes it (mul then add from matrix.hpp)
template <class Sim>
void plant(double* x, const double* u) {
  double d[8];
  Sim::dxdt(d, x, u);
  mul(d, d, Sim::dt, Sim::dim_x);
  add(x, x, d, Sim::dim_x);
}

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct Job {
  int model, p_full, newton_iters, n_steps, rec_stride;
  int64_t n;
  const double *x0, *p, *u0;
  double *x_traj, *u_traj, *x_fin, *u_fin, *U_fin, *dUdt_fin, *ctl_seconds;
  double* loop_seconds;  // [n_threads]: wall time each thread spent inside its closed-loop step loops
};
double g_last_loop_seconds = 0.0;

template <class Model, class Sim>
void run_slice(const Job& jb, int tid, int nth) {
  constexpr int nx = Model::dim_x, nu = Model::dim_u, np = Model::dim_p, L = Model::dim_u * Model::dv;
  const int plen = jb.p_full ? np * (Model::dv + 1) : np;
  for (int64_t n = tid; n < jb.n; n += nth) {
    double x[8], u[8], p0[8] = {0};
    Ctl<Model> ctl;
    const double* pn = jb.p ? jb.p + (size_t)plen * n : nullptr;
    if (np > 0) {
      if (jb.p_full)
        ctl.set_ptau(pn);
      else
        ctl.set_ptau_repeat(pn);
      for (int j = 0; j < np; j++) p0[j] = pn[j];
    }
    for (int j = 0; j < nx; j++) x[j] = jb.x0[(size_t)nx * n + j];
    for (int j = 0; j < nu; j++) u[j] = jb.u0[(size_t)nu * n + j];
    ctl.init_u0(u);
    ctl.init_u0_newton(u, x, p0, jb.newton_iters);
    double acc = 0.0;
    const double loop0 = now_s();
    for (int s = 0; s < jb.n_steps; s++) {
      const double t0 = now_s();
      ctl.control(u, x);
      acc += now_s() - t0;
      plant<Sim>(x, u);
      if (jb.rec_stride > 0 && (s + 1) % jb.rec_stride == 0) {
        const size_t r = (size_t)((s + 1) / jb.rec_stride - 1);
        if (jb.x_traj) memcpy(jb.x_traj + (r * jb.n + n) * nx, x, sizeof(double) * nx);
        if (jb.u_traj) memcpy(jb.u_traj + (r * jb.n + n) * nu, u, sizeof(double) * nu);
      }
    }
    if (jb.loop_seconds) jb.loop_seconds[tid] += now_s() - loop0;
    if (jb.x_fin) memcpy(jb.x_fin + (size_t)nx * n, x, sizeof(double) * nx);
    if (jb.u_fin) memcpy(jb.u_fin + (size_t)nu * n, u, sizeof(double) * nu);
    if (jb.U_fin) memcpy(jb.U_fin + (size_t)L * n, ctl.c.U, sizeof(double) * L);
    if (jb.dUdt_fin) memcpy(jb.dUdt_fin + (size_t)L * n, ctl.c.dUdt, sizeof(double) * L);
    if (jb.ctl_seconds) jb.ctl_seconds[n] = acc;
  }
}

void run_slice_any(const Job& jb, int tid, int nth) {
  switch (jb.model) {
    case ORACLE_MODEL_MSD: run_slice<ref_msd::Model, ref_msd::Simulator>(jb, tid, nth); break;
    case ORACLE_MODEL_ARM: run_slice<ref_arm::Model, ref_arm::Simulator>(jb, tid, nth); break;
    case ORACLE_MODEL_SEMIACTIVE: run_slice<ref_sad::Model, ref_sad::Simulator>(jb, tid, nth); break;
  }
}

template <class Model>
void fill_dims(int* d) {
  d[0] = Model::dim_x;
  d[1] = Model::dim_u;
  d[2] = Model::dim_p;
  d[3] = Model::dv;
  d[4] = Model::k_max;
  d[5] = Model::control_input;
}
template <class Model>
void fill_params(double* p) {
  p[0] = Model::dt;
  p[1] = Model::h;
  p[2] = Model::zeta;
  p[3] = Model::Tf;
  p[4] = Model::alpha;
  p[5] = Model::tol;
}

}  // namespace

extern "C" {

int ref_model_dims(int model, int* dims) {
  switch (model) {
    case ORACLE_MODEL_MSD: fill_dims<ref_msd::Model>(dims); return 0;
    case ORACLE_MODEL_ARM: fill_dims<ref_arm::Model>(dims); return 0;
    case ORACLE_MODEL_SEMIACTIVE: fill_dims<ref_sad::Model>(dims); return 0;
  }
  return -1;
}
int ref_model_params(int model, double* par) {
  switch (model) {
    case ORACLE_MODEL_MSD: fill_params<ref_msd::Model>(par); return 0;
    case ORACLE_MODEL_ARM: fill_params<ref_arm::Model>(par); return 0;
    case ORACLE_MODEL_SEMIACTIVE: fill_params<ref_sad::Model>(par); return 0;
  }
  return -1;
}

void* ref_create(int model) {
  switch (model) {
    case ORACLE_MODEL_MSD: return new Ctl<ref_msd::Model>();
    case ORACLE_MODEL_ARM: return new Ctl<ref_arm::Model>();
    case ORACLE_MODEL_SEMIACTIVE: return new Ctl<ref_sad::Model>();
  }
  return nullptr;
}
void ref_destroy(void* h) { delete (CtlBase*)h; }
void ref_set_ptau(void* h, const double* p) { ((CtlBase*)h)->set_ptau(p); }
void ref_set_ptau_repeat(void* h, const double* p) { ((CtlBase*)h)->set_ptau_repeat(p); }
void ref_init_u0(void* h, const double* u0) { ((CtlBase*)h)->init_u0(u0); }
void ref_init_u0_newton(void* h, double* u0, const double* x0, const double* p0, int n_loop) {
  ((CtlBase*)h)->init_u0_newton(u0, x0, p0, n_loop);
}
void ref_control(void* h, double* u, const double* x) { ((CtlBase*)h)->control(u, x); }
double ref_get_dtau(void* h, double t) { return ((CtlBase*)h)->get_dtau(t); }
void ref_get_state(void* h, double* t, double* U, double* dUdt) { ((CtlBase*)h)->get_state(t, U, dUdt); }
void ref_set_state(void* h, const double* t, const double* U, const double* dUdt) {
  ((CtlBase*)h)->set_state(t, U, dUdt);
}
// the reference exposes no exit-path information (it only printf()s "Breakdown", include/gmres.hpp:64)
int ref_last_status(void*) { return -1; }

void ref_plant_step(int model, double* x, const double* u) {
  switch (model) {
    case ORACLE_MODEL_MSD: plant<ref_msd::Simulator>(x, u); break;
    case ORACLE_MODEL_ARM: plant<ref_arm::Simulator>(x, u); break;
    case ORACLE_MODEL_SEMIACTIVE: plant<ref_sad::Simulator>(x, u); break;
  }
}

int ref_run_closed_loop(int model, int64_t n, const double* x0, const double* p, int p_full, const double* u0,
                        int newton_iters, int n_steps, int rec_stride, double* x_traj, double* u_traj,
                        double* x_fin, double* u_fin, double* U_fin, double* dUdt_fin, int32_t* exit_hist,
                        double* ctl_seconds, int n_threads) {
  int dims[6];
  if (ref_model_dims(model, dims) != 0 || n < 0 || !x0 || !u0 || (dims[2] > 0 && !p) || n_steps < 0) return -1;
  if (exit_hist) memset(exit_hist, 0xff, sizeof(int32_t) * 4 * (size_t)n);  // -1: not available
  if (n_threads < 1) n_threads = 1;
  std::vector<double> loop_s((size_t)n_threads, 0.0);
  Job jb{model, p_full, newton_iters, n_steps, rec_stride, n, x0, p, u0,
         x_traj, u_traj, x_fin, u_fin, U_fin, dUdt_fin, ctl_seconds, loop_s.data()};
  if (n_threads == 1) {
    run_slice_any(jb, 0, 1);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(run_slice_any, std::cref(jb), t, n_threads);
    for (auto& t : th) t.join();
  }
  g_last_loop_seconds = 0.0;
  for (double v : loop_s) g_last_loop_seconds = v > g_last_loop_seconds ? v : g_last_loop_seconds;
  return 0;
}

// max over the worker threads of the wall time spent inside the closed-loop step loops of the last
// run_closed_loop (controller construction, init_u0_newton and thread start-up excluded)
double ref_last_loop_seconds(void) { return g_last_loop_seconds; }

}  // extern "C"
