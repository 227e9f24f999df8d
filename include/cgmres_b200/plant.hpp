// plant.hpp -- one plant step x <- Phi(x, u) with the input held over the sampling period, host + device.
//
//   PLANT_EULER  x += Simulator::dxdt(x,u)*dt, the integrator of every reference example
//                (<example>/main.cpp:74-76: mul(dxdt, dxdt, dt); add(x, x, dxdt)); the parity default.
//   PLANT_RK4    classical 4th-order Runge-Kutta on the same Simulator::dxdt (BASELINE.json's north star asks for an
//                RK4 plant on the device; the reference itself never integrates with RK4, so this option has no
//                reference counterpart: the tests check it against an independent numpy RK4 of the same plant).
#pragma once
#include "cgmres_b200/models.hpp"

namespace cgmres_b200 {

enum { PLANT_NONE = 0, PLANT_EULER = 1, PLANT_RK4 = 2 };

template <class Sim>
CGMRES_HD void plant_euler(double* x, const double* u) {
  constexpr int nx = Sim::dim_x;
  double f[nx];
  Sim::dxdt(f, x, u);
#pragma unroll
  for (int j = 0; j < nx; j++) {
    double m = f[j] * Sim::dt;  // two roundings, like the reference's mul + add
    x[j] = x[j] + m;
  }
}

template <class Sim>
CGMRES_HD void plant_rk4(double* x, const double* u) {
  constexpr int nx = Sim::dim_x;
  constexpr double dt = Sim::dt;
  double k1[nx], k2[nx], k3[nx], k4[nx], y[nx];
  Sim::dxdt(k1, x, u);
#pragma unroll
  for (int j = 0; j < nx; j++) y[j] = x[j] + (0.5 * dt) * k1[j];
  Sim::dxdt(k2, y, u);
#pragma unroll
  for (int j = 0; j < nx; j++) y[j] = x[j] + (0.5 * dt) * k2[j];
  Sim::dxdt(k3, y, u);
#pragma unroll
  for (int j = 0; j < nx; j++) y[j] = x[j] + dt * k3[j];
  Sim::dxdt(k4, y, u);
#pragma unroll
  for (int j = 0; j < nx; j++) x[j] = x[j] + (dt / 6.0) * (k1[j] + 2.0 * k2[j] + 2.0 * k3[j] + k4[j]);
}

template <class Sim>
CGMRES_HD void plant_step(int integrator, double* x, const double* u) {
  if (integrator == PLANT_RK4)
    plant_rk4<Sim>(x, u);
  else
    plant_euler<Sim>(x, u);
}

}  // namespace cgmres_b200
