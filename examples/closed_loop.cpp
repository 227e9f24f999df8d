// closed_loop.cpp -- the reference's example programs (<example>/main.cpp) as BATCHED runs on the B200 library.
//
//   closed_loop <mass_spring_damper|arm_type_inverted_pendulum|semiactive_damper|multiple_controller>
//               [n_instances=1] [mode: host|device = host] [steps = shipped count]
//
// (The reference's own main.cpp files compile unmodified against include/ -- see tests/test_dropin_mains.py; this
//  program is what they look like with n controllers instead of one.)
// Every example runs its shipped initial condition, replicated n times, and writes instance 0's trajectory in the
// reference's format: one "%f"-formatted line per step into <example>_x.txt / <example>_u.txt.
//   host   : the loop of main.cpp -- u = control(x) through host buffers, forward-Euler plant step on the host;
//            multiple_controller steps its two controllers alternately inside ONE loop, like
//            multiple_controller/main.cpp:104-118
//   device : step_closed_loop_log(): the loop runs on the GPU, the trajectory is recorded in device memory and copied
//            out chunk-wise (no host round trip per step); multiple_controller runs its two handles concurrently
// Build: g++ -O2 -std=c++17 -Iinclude examples/closed_loop.cpp -Lcgmres_cpp_b200 -lcgmres_b200 -o closed_loop
#include <stdio.h>
#include <string.h>
#include <sys/time.h>

#include <memory>
#include <string>
#include <vector>

#include "cgmres.hpp"

static double wall(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + (double)tv.tv_usec * 1e-6;
}

// one example = n replicated controllers of one problem plus its output files
template <class Model, class Simulator>
struct Loop {
  static constexpr int nx = Model::dim_x, nu = Model::dim_u, np = Model::dim_p;
  int64_t n;
  std::vector<double> x, u, p;
  Cgmres<Model> controller;
  FILE *fx = nullptr, *fu = nullptr;
  double t_control = 0;

  Loop(const char* tag, int64_t n_, const std::vector<double>& x0, const std::vector<double>& u0,
       const std::vector<double>& p0)
      : n(n_), x(n_ * nx), u(n_ * nu), p(n_ * (np > 0 ? np : 1)), controller(n_) {
    for (int64_t i = 0; i < n; i++) {
      for (int j = 0; j < nx; j++) x[i * nx + j] = x0[j];
      for (int j = 0; j < nu; j++) u[i * nu + j] = u0[j];
      for (int j = 0; j < np; j++) p[i * np + j] = p0[j];
    }
    if (np > 0) controller.set_ptau_repeat(p.data());
    controller.init_u0(u.data());
    controller.init_u0_newton(u.data(), x.data(), p.data(), 10);
    fx = fopen((std::string(tag) + "_x.txt").c_str(), "w");
    fu = fopen((std::string(tag) + "_u.txt").c_str(), "w");
    if (!fx || !fu) throw std::runtime_error("cannot open the output files");
  }
  ~Loop() {
    if (fx) fclose(fx);
    if (fu) fclose(fu);
  }

  void print_row(int i, const double* xr, const double* ur) {  // <example>/main.cpp:78-87, instance 0
    fprintf(fx, "%f", Simulator::dt * i);
    fprintf(fu, "%f", Simulator::dt * i);
    for (int j = 0; j < nx; j++) fprintf(fx, "\t%f", xr[j]);
    for (int j = 0; j < nu; j++) fprintf(fu, "\t%f", ur[j]);
    fprintf(fx, "\n");
    fprintf(fu, "\n");
  }

  // one iteration of the reference's loop body through host buffers (main.cpp:68-87)
  void host_step(int i) {
    const double t0 = wall();
    controller.control(u.data(), x.data());
    t_control += wall() - t0;
    double d[nx];
    for (int64_t k = 0; k < n; k++) {  // x = x + dxdt * dt
      Simulator::dxdt(d, &x[k * nx], &u[k * nu]);
      for (int j = 0; j < nx; j++) d[j] = d[j] * Simulator::dt;
      for (int j = 0; j < nx; j++) x[k * nx + j] = x[k * nx + j] + d[j];
    }
    print_row(i, x.data(), u.data());
  }

  // steps [first, first + count) on the device, trajectory logged in device memory
  void device_steps(int first, int count) {
    std::vector<double> xl((size_t)count * n * nx), ul((size_t)count * n * nu);
    const double t0 = wall();
    controller.step_closed_loop_log(count, xl.data(), ul.data());
    t_control += wall() - t0;
    for (int s = 0; s < count; s++) print_row(first + s, &xl[(size_t)s * n * nx], &ul[(size_t)s * n * nu]);
  }
};

template <class A, class B>
static double run_together(A* a, int steps_a, B* b, int steps_b, bool on_device) {
  const int steps = steps_a > steps_b ? steps_a : steps_b;
  if (on_device) {
    if (a) a->controller.set_x(a->x.data());
    if (b) b->controller.set_x(b->x.data());
    const int chunk = 1000;
    for (int s = 0; s < steps; s += chunk) {
      if (a && s < steps_a) a->device_steps(s, steps_a - s < chunk ? steps_a - s : chunk);
      if (b && s < steps_b) b->device_steps(s, steps_b - s < chunk ? steps_b - s : chunk);
    }
  } else {
    for (int i = 0; i < steps; i++) {  // both controllers inside one loop (multiple_controller/main.cpp:104-118)
      if (a && i < steps_a) a->host_step(i);
      if (b && i < steps_b) b->host_step(i);
    }
  }
  return (a ? a->t_control : 0.0) + (b ? b->t_control : 0.0);
}

int main(int argc, char** argv) {
  const std::string ex = argc > 1 ? argv[1] : "mass_spring_damper";
  const int64_t n = argc > 2 ? atoll(argv[2]) : 1;
  const bool dev = argc > 3 && strcmp(argv[3], "device") == 0;
  const int steps_arg = argc > 4 ? atoi(argv[4]) : -1;
  using namespace cgmres_b200;
  using Msd = Loop<MassSpringDamperModel, MassSpringDamperSimulator>;
  using Arm = Loop<ArmPendulumModel, ArmPendulumSimulator>;
  using Sad = Loop<SemiactiveDamperModel, SemiactiveDamperSimulator>;
  const double pi = 3.14159265358979;
  const bool both = ex == "multiple_controller";
  double t = 0;
  try {
    std::unique_ptr<Msd> msd;
    std::unique_ptr<Arm> arm;
    std::unique_ptr<Sad> sad;
    if (ex == "mass_spring_damper" || both)
      msd.reset(new Msd(both ? "multiple_controller_1" : "mass_spring_damper", n, {2.0, 2.0, 0.0, 0.0},
                        {0.0, 0.0, 10.0, 10.0, 5e-4, 5e-4}, {1.0, -1.0}));
    if (ex == "arm_type_inverted_pendulum" || both)
      arm.reset(new Arm(both ? "multiple_controller_2" : "arm_type_inverted_pendulum", n, {pi, pi, 0.0, 0.0},
                        {0.0, 3.0, 0.01}, {pi / 4.0, 0.0}));
    if (ex == "semiactive_damper")
      sad.reset(new Sad("semiactive_damper", n, {2.0, 0.0}, {0.028393761456740, 0.166095020295846, 0.030103250483332},
                        {}));
    const int s_msd = steps_arg >= 0 ? steps_arg : (both ? 10001 : 20001);
    const int s_arm = steps_arg >= 0 ? steps_arg : 10001;
    const int s_sad = steps_arg >= 0 ? steps_arg : 20001;
    if (sad)
      t = run_together<Sad, Sad>(sad.get(), s_sad, nullptr, 0, dev);
    else
      t = run_together<Msd, Arm>(msd.get(), msd ? s_msd : 0, arm.get(), arm ? s_arm : 0, dev);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  printf("Elapsed time = %f\n", t);
  return 0;
}
