// models.hpp -- the three problem definitions of the reference, as host+device functors.
//
// Contract (identical to the reference's duck-typed Model / Simulator classes,
// e.g. mass_spring_damper/model.hpp:4-125 and simulator.hpp:4-22): static
// constexpr sizes and solver parameters plus the static functions
//     dxdt(ret,x,u,p)  dPhidx(ret,x,p)  dHdx(ret,x,u,p,lmd)  dHdu(ret,x,u,p,lmd)  ddHduu(ret,x,u,p,lmd)
// each additionally CGMRES_HD (__host__ __device__) so the sm_100a kernels inline them.
// Input box constraints use the reference's dummy-input + multiplier formulation
// (u = [controls, dummies, multipliers]).
//
// The expression trees are kept in the reference's evaluation order (C++ parses
// a*b*c as (a*b)*c and a+b-c as (a+b)-c) because the exact build mode has to be
// bit-identical to the reference compiled without FMA contraction.
#pragma once
#include <math.h>
#include <stdint.h>

#include "cgmres_b200/portable_trig.hpp"

#if defined(__CUDACC__)
#define CGMRES_HD __host__ __device__ __forceinline__
#else
#define CGMRES_HD inline
#endif

namespace cgmres_b200 {

// Solver constants shared by every shipped example (<example>/model.hpp:19-34).
struct SolverDefaults {
  static constexpr double dt = 0.001;     // sampling period [s]
  static constexpr double h = 0.002;      // forward-difference step [s]
  static constexpr double zeta = 1000.0;  // stabilisation gain
  static constexpr double alpha = 0.5;    // horizon rise rate
  static constexpr double tol = 1e-6;     // GMRES tolerance
  static constexpr uint16_t k_max = 5;    // GMRES iterations
};

// Horizon step of the C/GMRES continuation, cgmres.hpp:32-34: dtau(t) = Tf*(1 - exp(-alpha*t))/dv.
template <class M>
CGMRES_HD double horizon_dtau(double t) {
  return M::Tf * (1 - exp(-M::alpha * t)) / (double)M::dv;
}

// ---------------------------------------------------------------------------
// Two-mass spring/damper chain, two bounded forces.
// reference: mass_spring_damper/model.hpp (and multiple_controller/model1.hpp)
// ---------------------------------------------------------------------------
struct MassSpringDamperPlantConstants {
  static constexpr double m1 = 1.0, m2 = 1.0, d1 = 1.0, d2 = 1.0, k1 = 1.0, k2 = 1.0;  // model.hpp:124
  // NOTE: the state equation uses -(k1*k2)/m1 where dHdx uses -(k1+k2)/m1.
  // That inconsistency is the reference's and is reproduced on purpose.
  static constexpr double a20 = -(k1 * k2) / m1, a21 = k2 / m1, a22 = (d1 + d2) / m1, a23 = d2 / m1;
  static constexpr double a30 = k2 / m2, a31 = k2 / m2, a32 = d2 / m2, a33 = d2 / m2;

  template <class X, class U>
  static CGMRES_HD void rhs(double* ret, const X& x, const U& u) {  // model.hpp:36-41 == simulator.hpp:13-18
    ret[0] = x[2];
    ret[1] = x[3];
    ret[2] = a20 * x[0] + a21 * x[1] - a22 * x[2] + a23 * x[3] + u[0] / m1;
    ret[3] = a30 * x[0] - a31 * x[1] + a32 * x[2] - a33 * x[3] + u[1] / m2;
  }
};

struct MassSpringDamperModel : SolverDefaults, MassSpringDamperPlantConstants {
  static constexpr uint16_t dim_x = 4;
  static constexpr uint16_t control_input = 2, constraint = 2, dummy = 2;
  static constexpr uint16_t dim_u = control_input + constraint + dummy;
  static constexpr uint16_t dim_p = 2;
  static constexpr uint16_t dv = 50;
  static constexpr double Tf = 1.0;
  // kernel hint (not part of the reference's contract): dHdu below ignores its x argument (the plant is linear in u
  // with constant gains), so the on-chip kernels may let the costates overwrite the rollout states in place
  static constexpr bool dHdu_reads_x = false;

  // weights, model.hpp:112-115
  static constexpr double sf0 = 10.0, sf1 = 10.0, sf2 = 1.0, sf3 = 1.0;
  static constexpr double q0 = 1.0, q1 = 1.0, q2 = 10.0, q3 = 10.0;
  static constexpr double r0 = 0.1, r1 = 0.1, r2 = 0.01, r3 = 0.01;
  // |u - uc| <= ur, model.hpp:118-121
  static constexpr double umin = -10.0, umax = 10.0;
  static constexpr double uc = (umax + umin) / 2.0, ur = (umax - umin) / 2.0;
  // coefficients of the costate equation, model.hpp:51-54
  static constexpr double b02 = (k1 + k2) / m1, b03 = k2 / m2, b12 = k2 / m1, b13 = k2 / m2;
  static constexpr double b22 = (d1 + d2) / m1, b23 = d2 / m2, b32 = d2 / m1, b33 = d2 / m2;

  static CGMRES_HD void dxdt(double* ret, const double* x, const double* u, const double*) { rhs(ret, x, u); }

  static CGMRES_HD void dPhidx(double* ret, const double* x, const double* p) {  // model.hpp:43-48
    ret[0] = -(p[0] - x[0]) * sf0;
    ret[1] = -(p[1] - x[1]) * sf1;
    ret[2] = x[2] * sf2;
    ret[3] = x[3] * sf3;
  }

  static CGMRES_HD void dHdx(double* ret, const double* x, const double*, const double* p, const double* lmd) {
    ret[0] = -(p[0] - x[0]) * q0 - b02 * lmd[2] + b03 * lmd[3];
    ret[1] = -(p[1] - x[1]) * q1 + b12 * lmd[2] - b13 * lmd[3];
    ret[2] = x[2] * q2 + lmd[0] - b22 * lmd[2] + b23 * lmd[3];
    ret[3] = x[3] * q3 + lmd[1] + b32 * lmd[2] - b33 * lmd[3];
  }

  // u = [f1, f2, dummy1, dummy2, mult1, mult2]; model.hpp:57-64
  static CGMRES_HD void dHdu(double* ret, const double*, const double* u, const double*, const double* lmd) {
    ret[0] = r0 * u[0] + lmd[2] / m1 + 2.0 * u[4] * (u[0] - uc);
    ret[1] = r1 * u[1] + lmd[3] / m2 + 2.0 * u[5] * (u[1] - uc);
    ret[2] = -r2 + 2.0 * u[4] * u[2];
    ret[3] = -r3 + 2.0 * u[5] * u[3];
    ret[4] = (u[0] - uc) * (u[0] - uc) + u[2] * u[2] - ur * ur;
    ret[5] = (u[1] - uc) * (u[1] - uc) + u[3] * u[3] - ur * ur;
  }

  // symmetric 6x6, column major (entry (i,j) at 6*j+i); model.hpp:66-108
  static CGMRES_HD void ddHduu(double* ret, const double*, const double* u, const double*, const double*) {
    for (int i = 0; i < 36; i++) ret[i] = 0;
    ret[6 * 0 + 0] = r0 + 2 * u[4];
    ret[6 * 1 + 1] = r1 + 2 * u[5];
    ret[6 * 2 + 2] = 2 * u[4];
    ret[6 * 3 + 3] = 2 * u[5];
    ret[6 * 0 + 4] = ret[6 * 4 + 0] = 2 * (u[0] - uc);
    ret[6 * 1 + 5] = ret[6 * 5 + 1] = 2 * (u[1] - uc);
    ret[6 * 2 + 4] = ret[6 * 4 + 2] = 2 * u[2];
    ret[6 * 3 + 5] = ret[6 * 5 + 3] = 2 * u[3];
  }
};

struct MassSpringDamperSimulator : MassSpringDamperPlantConstants {  // mass_spring_damper/simulator.hpp:4-22
  static constexpr double t_end = 20;
  static constexpr double dt = 0.001;
  static constexpr uint16_t dim_x = 4, dim_u = 6, dim_p = 2, dv = 50;
  static CGMRES_HD void dxdt(double* ret, const double* x, const double* u) { rhs(ret, x, u); }
};

// ---------------------------------------------------------------------------
// Arm-type (Furuta-like) inverted pendulum, one bounded torque.
// reference: arm_type_inverted_pendulum/model.hpp (and multiple_controller/model2.hpp)
// ---------------------------------------------------------------------------
struct ArmPendulumPlantConstants {
  static constexpr double As = 6.25, Bs = 15.6, A52 = 39.1111, C22 = 0.0407448;  // model.hpp:92-98
  static constexpr double A32a = 5.65635, A32 = 0.905016, A32b = 14.1183;

  // sin/cos of (x0-x1) and of x1 are evaluated once per call; every use in the reference has the same argument,
  // so the values (and hence the results) are the same.  They come from portable_trig.hpp (pure +,-,*; <= 1 ulp
  // from glibc) so that host and device agree bit for bit; see that header for what this means for parity.
  template <class X, class U>
  static CGMRES_HD void rhs(double* ret, const X& x, const U& u) {  // model.hpp:37-42 == simulator.hpp:14-19
    const double d = x[0] - x[1];
    double sd, cd;
    ptrig::psincos(d, &sd, &cd);
    const double s1 = ptrig::psin(x[1]);
    ret[0] = x[2];
    ret[1] = x[3];
    ret[2] = -As * x[2] + Bs * u[0];
    ret[3] = A32 * x[2] * x[2] * sd + A52 * s1 - A32b * cd * u[0] + A32a * cd * x[2] + C22 * (x[2] - x[3]);
  }
};

struct ArmPendulumModel : SolverDefaults, ArmPendulumPlantConstants {
  static constexpr uint16_t dim_x = 4;
  static constexpr uint16_t control_input = 1, constraint = 1, dummy = 1;
  static constexpr uint16_t dim_u = control_input + constraint + dummy;
  static constexpr uint16_t dim_p = 2;
  static constexpr uint16_t dv = 25;
  static constexpr double Tf = 0.5;
  static constexpr bool dHdu_reads_x = true;  // cos(x0 - x1) in dHdu

  static constexpr double sf0 = 3.0, sf1 = 1.0, sf2 = 0.0, sf3 = 0.0;  // model.hpp:81
  static constexpr double q0 = 1.0, q1 = 1.0, q2 = 0.0, q3 = 0.0;      // model.hpp:82
  static constexpr double r0 = 1.0, r1 = 0.1;                          // model.hpp:83
  static constexpr double umin = -3.0, umax = 3.0;
  static constexpr double uc = (umax + umin) / 2.0, ur = (umax - umin) / 2.0;

  static CGMRES_HD void dxdt(double* ret, const double* x, const double* u, const double*) { rhs(ret, x, u); }

  static CGMRES_HD void dPhidx(double* ret, const double* x, const double* p) {  // model.hpp:44-49
    ret[0] = (x[0] - p[0]) * sf0;
    ret[1] = (x[1] - p[1]) * sf1;
    ret[2] = x[2] * sf2;
    ret[3] = x[3] * sf3;
  }

  static CGMRES_HD void dHdx(double* ret, const double* x, const double* u, const double* p, const double* lmd) {
    const double d = x[0] - x[1];
    double sd, cd;
    ptrig::psincos(d, &sd, &cd);  // model.hpp:52-54
    const double c1 = ptrig::pcos(x[1]);
    ret[0] = (x[0] - p[0]) * q0 + lmd[3] * (A32 * x[2] * x[2] * cd + A32b * sd * u[0] - A32a * sd * x[2]);
    ret[1] = (x[1] - p[1]) * q1 + lmd[3] * (-A32 * x[2] * x[2] * cd + A52 * c1 - A32b * sd * u[0] + A32a * sd * x[2]);
    ret[2] = x[2] * q2 + lmd[0] - lmd[2] * As + lmd[3] * (0.2e1 * A32 * x[2] * sd + A32a * cd + C22);
    ret[3] = x[3] * q3 + lmd[1] - lmd[3] * C22;
  }

  // u = [torque, dummy, multiplier]; model.hpp:58-62
  static CGMRES_HD void dHdu(double* ret, const double* x, const double* u, const double*, const double* lmd) {
    const double cd = ptrig::pcos(x[0] - x[1]);
    ret[0] = (r0 * u[0]) + lmd[2] * Bs - lmd[3] * A32b * cd + (u[2] * (2.0 * u[0] - 2.0 * uc));
    ret[1] = -0.5 * r1 + (2.0 * u[2] * u[1]);
    ret[2] = (u[0] - uc) * (u[0] - uc) + u[1] * u[1] - ur * ur;
  }

  // 3x3 column major; model.hpp:64-76
  static CGMRES_HD void ddHduu(double* ret, const double*, const double* u, const double*, const double*) {
    ret[0] = r0 + 2 * u[2];
    ret[1] = 0;
    ret[2] = 2 * u[0] - 2 * uc;
    ret[3] = 0;
    ret[4] = 2 * u[2];
    ret[5] = 2 * u[1];
    ret[6] = 2 * u[0] - 2 * uc;
    ret[7] = 2 * u[1];
    ret[8] = 0;
  }
};

struct ArmPendulumSimulator : ArmPendulumPlantConstants {  // arm_type_inverted_pendulum/simulator.hpp:5-29
  static constexpr double t_end = 10;
  static constexpr double dt = 0.001;
  static constexpr uint16_t dim_x = 4, dim_u = 3, dim_p = 2, dv = 25;
  static CGMRES_HD void dxdt(double* ret, const double* x, const double* u) { rhs(ret, x, u); }
};

// ---------------------------------------------------------------------------
// Semi-active damper: bilinear plant, damping coefficient bounded to [0,1].
// reference: semiactive_damper/model.hpp
// ---------------------------------------------------------------------------
struct SemiactiveDamperPlantConstants {
  static constexpr double a = -1.0, b = -1.0;  // model.hpp:84-85

  template <class X, class U>
  static CGMRES_HD void rhs(double* ret, const X& x, const U& u) {  // model.hpp:36-39 == simulator.hpp:13-16
    ret[0] = x[1];
    ret[1] = a * x[0] + b * u[0] * x[1];
  }
};

struct SemiactiveDamperModel : SolverDefaults, SemiactiveDamperPlantConstants {
  static constexpr uint16_t dim_x = 2;
  static constexpr uint16_t control_input = 1, constraint = 1, dummy = 1;
  static constexpr uint16_t dim_u = control_input + constraint + dummy;
  static constexpr uint16_t dim_p = 0;
  static constexpr uint16_t dv = 50;
  static constexpr double Tf = 1.0;
  static constexpr bool dHdu_reads_x = true;  // b*x[1]*lmd[1] in dHdu

  static constexpr double sf0 = 1.0, sf1 = 10.0;  // model.hpp:74
  static constexpr double q0 = 1.0, q1 = 10.0;    // model.hpp:75
  static constexpr double r0 = 1.0, r1 = 0.01;    // model.hpp:76
  static constexpr double umin = 0.0, umax = 1.0;
  static constexpr double uc = (umax + umin) / 2.0, ur = (umax - umin) / 2.0;

  static CGMRES_HD void dxdt(double* ret, const double* x, const double* u, const double*) { rhs(ret, x, u); }

  static CGMRES_HD void dPhidx(double* ret, const double* x, const double*) {  // model.hpp:41-44
    ret[0] = x[0] * sf0;
    ret[1] = x[1] * sf1;
  }

  static CGMRES_HD void dHdx(double* ret, const double* x, const double* u, const double*, const double* lmd) {
    ret[0] = x[0] * q0 + a * lmd[1];  // model.hpp:47-48
    ret[1] = x[1] * q1 + lmd[0] + b * u[0] * lmd[1];
  }

  // u = [damping, dummy, multiplier]; model.hpp:51-55
  static CGMRES_HD void dHdu(double* ret, const double* x, const double* u, const double*, const double* lmd) {
    ret[0] = r0 * u[0] + b * x[1] * lmd[1] + 2 * u[2] * (u[0] - uc);
    ret[1] = -r1 + 2 * u[1] * u[2];
    ret[2] = (u[0] - uc) * (u[0] - uc) + u[1] * u[1] - ur * ur;
  }

  // 3x3 column major; model.hpp:57-69
  static CGMRES_HD void ddHduu(double* ret, const double*, const double* u, const double*, const double*) {
    ret[0] = r0 + 2 * u[2];
    ret[1] = 0;
    ret[2] = 2 * (u[0] - uc);
    ret[3] = 0;
    ret[4] = 2 * u[2];
    ret[5] = 2 * u[1];
    ret[6] = 2 * (u[0] - uc);
    ret[7] = 2 * u[1];
    ret[8] = 0;
  }
};

struct SemiactiveDamperSimulator : SemiactiveDamperPlantConstants {  // semiactive_damper/simulator.hpp:4-21
  static constexpr double t_end = 20;
  static constexpr double dt = 0.001;
  static constexpr uint16_t dim_x = 2, dim_u = 3, dim_p = 0, dv = 50;
  static CGMRES_HD void dxdt(double* ret, const double* x, const double* u) { rhs(ret, x, u); }
};

}  // namespace cgmres_b200
