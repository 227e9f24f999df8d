// cgmres.hpp -- header-only drop-in for the reference's `Cgmres<Model>` (include/cgmres.hpp:8-207),
// backed by the B200 library through the C ABI of cgmres_b200.h.
//
//   reference                                   this header
//   ------------------------------------------  -----------------------------------------------------------
//   Cgmres<Model> c;            one controller  Cgmres<Model> c;  or  Cgmres<Model> c(n, device, mode);
//   c.set_ptau / set_ptau_repeat / init_u0 /    same names, same argument meaning; every array carries a
//   init_u0_newton / control / get_dtau         leading instance dimension (n == 1: the reference's arrays)
//   main.cpp: control() + Euler plant step      c.step_closed_loop(k): the loop body stays on the device
//
// `Model` is one of the problem classes of cgmres_b200/models.hpp (the per-example model.hpp headers in this
// include tree export them under the reference's names Model / Model1 / Model2).  The kernels are compiled
// per model into libcgmres_b200.so; a new problem is added by writing its functor next to the shipped ones and
// instantiating the kernels for it ("adding a model").
//
// Error behaviour: the reference has none (it printf()s "Breakdown" and carries on).  Numerical exit paths are
// reported by status(); API / CUDA failures throw std::runtime_error with cgmres_b200_last_error().
#pragma once
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "cgmres_b200.h"
#include "cgmres_b200/models.hpp"
#include "gmres.hpp"

namespace cgmres_b200 {
// Which compiled kernel serves a host-side Model class.  The kernels are compiled per problem into
// libcgmres_b200.so, so the template argument only has to NAME one of them:
//  * the functors of cgmres_b200/models.hpp (and the per-example headers of this include tree that export them
//    as Model / Model1 / Model2) map directly;
//  * any other class with the reference's Model contract -- in particular the reference's own
//    <example>/model.hpp, which an unmodified main.cpp picks up from its own directory -- is identified by its
//    compile-time sizes and parameters and then PROBED: dxdt, dPhidx, dHdx, dHdu are evaluated on the host at two
//    fixed points and must agree with the shipped functor (to 1e-12 relative: the arm functor's sin/cos differ
//    from libm's in the last bit).  A class that matches nothing throws instead of running the wrong problem.
template <class Ours, class Theirs>
inline bool same_problem(void) {
  if (Ours::dim_x != Theirs::dim_x || Ours::dim_u != Theirs::dim_u || Ours::dim_p != Theirs::dim_p ||
      Ours::dv != Theirs::dv || Ours::k_max != Theirs::k_max)
    return false;
  const double pa[6] = {Ours::dt, Ours::h, Ours::zeta, Ours::Tf, Ours::alpha, Ours::tol};
  const double pb[6] = {Theirs::dt, Theirs::h, Theirs::zeta, Theirs::Tf, Theirs::alpha, Theirs::tol};
  for (int i = 0; i < 6; i++)
    if (pa[i] != pb[i]) return false;
  constexpr int nx = Ours::dim_x, nu = Ours::dim_u, np = Ours::dim_p > 0 ? Ours::dim_p : 1;
  for (int probe = 0; probe < 2; probe++) {
    double x[nx], u[nu], p[np], lmd[nx], a[nx > nu ? nx : nu], b[nx > nu ? nx : nu];
    for (int i = 0; i < nx; i++) x[i] = 0.3 + 0.17 * i - 0.45 * probe, lmd[i] = -0.2 + 0.11 * i + 0.3 * probe;
    for (int i = 0; i < nu; i++) u[i] = 0.05 + 0.23 * i - 0.1 * probe;
    for (int i = 0; i < np; i++) p[i] = 0.4 - 0.3 * i;
    auto agree = [&](int cnt) {
      for (int i = 0; i < cnt; i++) {
        const double scale = (a[i] < 0 ? -a[i] : a[i]) + 1.0;
        const double d = a[i] - b[i];
        if ((d < 0 ? -d : d) > 1e-12 * scale) return false;
      }
      return true;
    };
    Ours::dxdt(a, x, u, p), Theirs::dxdt(b, x, u, p);
    if (!agree(nx)) return false;
    Ours::dPhidx(a, x, p), Theirs::dPhidx(b, x, p);
    if (!agree(nx)) return false;
    Ours::dHdx(a, x, u, p, lmd), Theirs::dHdx(b, x, u, p, lmd);
    if (!agree(nx)) return false;
    Ours::dHdu(a, x, u, p, lmd), Theirs::dHdu(b, x, u, p, lmd);
    if (!agree(nu)) return false;
  }
  return true;
}

template <class Model>
struct ModelId {
  static int value(void) {
    if (same_problem<MassSpringDamperModel, Model>()) return CGMRES_B200_MODEL_MASS_SPRING_DAMPER;
    if (same_problem<ArmPendulumModel, Model>()) return CGMRES_B200_MODEL_ARM_TYPE_INVERTED_PENDULUM;
    if (same_problem<SemiactiveDamperModel, Model>()) return CGMRES_B200_MODEL_SEMIACTIVE_DAMPER;
    throw std::runtime_error(
        "Cgmres<Model>: no kernel in libcgmres_b200.so was compiled for this Model (add its functor to "
        "include/cgmres_b200/models.hpp and instantiate the kernels for it)");
  }
};
template <>
struct ModelId<MassSpringDamperModel> {
  static int value(void) { return CGMRES_B200_MODEL_MASS_SPRING_DAMPER; }
};
template <>
struct ModelId<ArmPendulumModel> {
  static int value(void) { return CGMRES_B200_MODEL_ARM_TYPE_INVERTED_PENDULUM; }
};
template <>
struct ModelId<SemiactiveDamperModel> {
  static int value(void) { return CGMRES_B200_MODEL_SEMIACTIVE_DAMPER; }
};
inline void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + cgmres_b200_last_error());
}
}  // namespace cgmres_b200

template <class Model>
class Cgmres : public Gmres {
 public:
  // n_instances independent controllers on `device`; the default is the reference's single object.
  explicit Cgmres(int64_t n_instances = 1, int device = 0, int mode = CGMRES_B200_MODE_ONCHIP_EXACT)
      : Gmres(len, Model::k_max, Model::tol), n_(n_instances), h_(nullptr) {
    cgmres_b200::check(cgmres_b200_create(cgmres_b200::ModelId<Model>::value(), n_instances, device, mode, &h_),
                       "cgmres_b200_create");
  }
  ~Cgmres(void) { cgmres_b200_destroy(h_); }

  int64_t size(void) const { return n_; }

  double get_dtau(const double t) const { return cgmres_b200_get_dtau(h_, t); }

  // ptau_buf[n][(dv+1)*dim_p] = [ p(t), p(t+dtau), ..., p(t+dv*dtau) ] per instance
  void set_ptau(const double* ptau_buf) { cgmres_b200::check(cgmres_b200_set_ptau(h_, ptau_buf), "set_ptau"); }
  // p_buf[n][dim_p], repeated over the horizon
  void set_ptau_repeat(const double* p_buf) {
    cgmres_b200::check(cgmres_b200_set_ptau_repeat(h_, p_buf), "set_ptau_repeat");
  }
  // u0[n][dim_u] copied to every stage of U
  void init_u0(const double* u0) { cgmres_b200::check(cgmres_b200_init_u0(h_, u0), "init_u0"); }
  // Newton refinement of u0 (mutated, like the reference) followed by init_u0
  void init_u0_newton(double* u0, const double* x0, const double* p0, const uint16_t n_loop) {
    cgmres_b200::check(cgmres_b200_init_u0_newton(h_, u0, x0, p0, n_loop), "init_u0_newton");
  }
  // one control update for every instance: x[n][dim_x] in, u[n][dim_u] out (host arrays)
  void control(double* u, const double* x) { cgmres_b200::check(cgmres_b200_control(h_, u, x), "control"); }

  // ---- batched extras (no counterpart in the reference) ----
  void set_x(const double* x) { cgmres_b200::check(cgmres_b200_set_x(h_, x), "set_x"); }
  void get_x(double* x) const { cgmres_b200::check(cgmres_b200_get_x(h_, x), "get_x"); }
  void get_u(double* u) const { cgmres_b200::check(cgmres_b200_get_u(h_, u), "get_u"); }
  // n_steps x { control ; x += Simulator::dxdt(x,u)*dt } with the state resident in HBM
  void step_closed_loop(int n_steps) {
    cgmres_b200::check(cgmres_b200_step_closed_loop(h_, n_steps), "step_closed_loop");
  }
  // the same with the trajectory recorded on the device: x_log[n_steps][n][dim_x], u_log[n_steps][n][dim_u]
  void step_closed_loop_log(int n_steps, double* x_log, double* u_log) {
    cgmres_b200::check(cgmres_b200_step_closed_loop_log(h_, n_steps, x_log, u_log), "step_closed_loop_log");
  }
  // per-instance controller clocks t[n] (nullptr: back to the batch-uniform clock)
  void set_t(const double* t) { cgmres_b200::check(cgmres_b200_set_t(h_, t), "set_t"); }
  void get_t(double* t) const { cgmres_b200::check(cgmres_b200_get_t(h_, t), "get_t"); }
  // plant integrator of step_closed_loop: CGMRES_B200_PLANT_EULER (reference) or CGMRES_B200_PLANT_RK4
  void set_plant_integrator(int integrator) {
    cgmres_b200::check(cgmres_b200_set_plant_integrator(h_, integrator), "set_plant_integrator");
  }
  void synchronize(void) { cgmres_b200::check(cgmres_b200_synchronize(h_), "synchronize"); }
  void get_state(double* t, double* U, double* dUdt) const {
    cgmres_b200::check(cgmres_b200_get_state(h_, t, U, dUdt), "get_state");
  }
  void set_state(const double* t, const double* U, const double* dUdt) {
    cgmres_b200::check(cgmres_b200_set_state(h_, t, U, dUdt), "set_state");
  }
  // exit path of each instance's last update (CGMRES_B200_EXIT_* | columns << 8)
  std::vector<int32_t> status(void) const {
    std::vector<int32_t> s((size_t)n_);
    if (n_ > 0) cgmres_b200::check(cgmres_b200_get_status(h_, s.data()), "get_status");
    return s;
  }
  cgmres_b200_handle handle(void) const { return h_; }

 public:
  // Parameters (cgmres.hpp:179-188 of the reference)
  static constexpr uint16_t dim_x = Model::dim_x;
  static constexpr uint16_t dim_u = Model::dim_u;
  static constexpr uint16_t dim_p = Model::dim_p;

  static constexpr double dt = Model::dt;
  static constexpr double h = Model::h;
  static constexpr double zeta = Model::zeta;
  static constexpr uint16_t dv = Model::dv;
  static constexpr double Tf = Model::Tf;
  static constexpr double alpha = Model::alpha;

 private:
  static constexpr uint16_t len = dim_u * dv;
  int64_t n_;
  cgmres_b200_handle h_;

  Cgmres(const Cgmres&);
  Cgmres& operator=(const Cgmres&);
};
