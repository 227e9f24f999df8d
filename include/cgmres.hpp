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
template <class Model>
struct ModelId;
template <>
struct ModelId<MassSpringDamperModel> {
  static constexpr int value = CGMRES_B200_MODEL_MASS_SPRING_DAMPER;
};
template <>
struct ModelId<ArmPendulumModel> {
  static constexpr int value = CGMRES_B200_MODEL_ARM_TYPE_INVERTED_PENDULUM;
};
template <>
struct ModelId<SemiactiveDamperModel> {
  static constexpr int value = CGMRES_B200_MODEL_SEMIACTIVE_DAMPER;
};
inline void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + cgmres_b200_last_error());
}
}  // namespace cgmres_b200

template <class Model>
class Cgmres : public Gmres {
 public:
  // n_instances independent controllers on `device`; the default is the reference's single object.
  explicit Cgmres(int64_t n_instances = 1, int device = 0, int mode = CGMRES_B200_MODE_EXACT)
      : Gmres(len, Model::k_max, Model::tol), n_(n_instances), h_(nullptr) {
    cgmres_b200::check(cgmres_b200_create(cgmres_b200::ModelId<Model>::value, n_instances, device, mode, &h_),
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
