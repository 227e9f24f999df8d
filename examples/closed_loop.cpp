// closed_loop.cpp -- the four example programs of the reference (<example>/main.cpp) on the B200 library.
//
//   closed_loop <mass_spring_damper|arm_type_inverted_pendulum|semiactive_damper|multiple_controller>
//               [n_instances=1] [mode: host|device = host] [steps = shipped count]
//
// Runs the shipped initial condition (replicated n times) through the same loop as the reference's main():
// control(u,x), forward-Euler plant step, one "%f"-formatted line per step into <example>_x.txt / _u.txt
// (instance 0), so the text files can be diffed against the reference's own output.
//   host   : u = control(x) through host buffers and the plant step on the host, exactly like main.cpp
//   device : step_closed_loop(1) + get_x/get_u (state never leaves HBM except for logging)
// Build: g++ -O2 -std=c++17 -Iinclude examples/closed_loop.cpp -Lcgmres_cpp_b200 -lcgmres_b200 -o closed_loop
#include <stdio.h>
#include <string.h>
#include <sys/time.h>

#include <string>
#include <vector>

#include "cgmres.hpp"

static double wall(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + (double)tv.tv_usec * 1e-6;
}

template <class Model, class Simulator>
static double run(const char* tag, int64_t n, bool on_device, int steps, const std::vector<double>& x0,
                  const std::vector<double>& u0, const std::vector<double>& p0) {
  constexpr int nx = Model::dim_x, nu = Model::dim_u, np = Model::dim_p;
  std::vector<double> x(n * nx), u(n * nu), p(n * (np > 0 ? np : 1)), d(nx);
  for (int64_t i = 0; i < n; i++) {
    for (int j = 0; j < nx; j++) x[i * nx + j] = x0[j];
    for (int j = 0; j < nu; j++) u[i * nu + j] = u0[j];
    for (int j = 0; j < np; j++) p[i * np + j] = p0[j];
  }
  Cgmres<Model> controller(n);
  if (np > 0) controller.set_ptau_repeat(p.data());
  controller.init_u0(u.data());
  controller.init_u0_newton(u.data(), x.data(), p.data(), 10);
  if (on_device) controller.set_x(x.data());

  FILE* fx = fopen((std::string(tag) + "_x.txt").c_str(), "w");
  FILE* fu = fopen((std::string(tag) + "_u.txt").c_str(), "w");
  if (!fx || !fu) return -1.0;
  double t_all = 0;
  for (int i = 0; i < steps; i++) {
    const double t0 = wall();
    if (on_device) {
      controller.step_closed_loop(1);
      controller.synchronize();
    } else {
      controller.control(u.data(), x.data());
    }
    t_all += wall() - t0;
    if (on_device) {
      controller.get_x(x.data());
      controller.get_u(u.data());
    } else {
      for (int64_t k = 0; k < n; k++) {  // x = x + dxdt * dt
        Simulator::dxdt(d.data(), &x[k * nx], &u[k * nu]);
        for (int j = 0; j < nx; j++) d[j] = d[j] * Simulator::dt;
        for (int j = 0; j < nx; j++) x[k * nx + j] = x[k * nx + j] + d[j];
      }
    }
    fprintf(fx, "%f", Simulator::dt * i);
    fprintf(fu, "%f", Simulator::dt * i);
    for (int j = 0; j < nx; j++) fprintf(fx, "\t%f", x[j]);
    for (int j = 0; j < nu; j++) fprintf(fu, "\t%f", u[j]);
    fprintf(fx, "\n");
    fprintf(fu, "\n");
  }
  fclose(fx);
  fclose(fu);
  return t_all;
}

int main(int argc, char** argv) {
  const std::string ex = argc > 1 ? argv[1] : "mass_spring_damper";
  const int64_t n = argc > 2 ? atoll(argv[2]) : 1;
  const bool dev = argc > 3 && strcmp(argv[3], "device") == 0;
  const int steps_arg = argc > 4 ? atoi(argv[4]) : -1;
  using namespace cgmres_b200;
  const double pi = 3.14159265358979;
  double t = 0;
  try {
    if (ex == "mass_spring_damper" || ex == "multiple_controller") {
      const char* tag = ex == "multiple_controller" ? "multiple_controller_1" : "mass_spring_damper";
      const int steps = steps_arg >= 0 ? steps_arg : (ex == "multiple_controller" ? 10001 : 20001);
      t += run<MassSpringDamperModel, MassSpringDamperSimulator>(tag, n, dev, steps, {2.0, 2.0, 0.0, 0.0},
                                                                  {0.0, 0.0, 10.0, 10.0, 5e-4, 5e-4}, {1.0, -1.0});
    }
    if (ex == "arm_type_inverted_pendulum" || ex == "multiple_controller") {
      const char* tag = ex == "multiple_controller" ? "multiple_controller_2" : "arm_type_inverted_pendulum";
      const int steps = steps_arg >= 0 ? steps_arg : 10001;
      t += run<ArmPendulumModel, ArmPendulumSimulator>(tag, n, dev, steps, {pi, pi, 0.0, 0.0}, {0.0, 3.0, 0.01},
                                                       {pi / 4.0, 0.0});
    }
    if (ex == "semiactive_damper") {
      const int steps = steps_arg >= 0 ? steps_arg : 20001;
      t += run<SemiactiveDamperModel, SemiactiveDamperSimulator>(
          "semiactive_damper", n, dev, steps, {2.0, 0.0}, {0.028393761456740, 0.166095020295846, 0.030103250483332}, {});
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  printf("Elapsed time = %f\n", t);
  return 0;
}
