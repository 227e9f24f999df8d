/*
 * portable_trig.h -- C restatement (TEST INFRASTRUCTURE, see cgmres_oracle.h) of the +,-,* only sin/cos that the
 * product's arm model uses (include/cgmres_b200/portable_trig.hpp).  Written separately on purpose: the oracle
 * must not include product headers.  Same algorithm (fdlibm: 3 x 33-bit Cody-Waite reduction by pi/2 with exact
 * error recovery, degree-13/14 kernels), therefore the same doubles when built with -ffp-contract=off.
 * Only compiled into the oracle when ORACLE_PORTABLE_TRIG is defined (libcgmres_oracle_ptrig.so); the default
 * oracle keeps glibc's sin/cos like the reference.
 */
#ifndef ORACLE_PORTABLE_TRIG_H
#define ORACLE_PORTABLE_TRIG_H
#include <math.h>

static double opt_ksin(double x, double y) {
  static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                      S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                      S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x, v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static double opt_kcos(double x, double y) {
  static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                      C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                      C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double hz = 0.5 * z, w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * r - x * y));
}

static void opt_sincos(double x, double* s, double* c) {
  static const double invpio2 = 6.36619772367581382433e-01, p1 = 1.57079632673412561417e+00,
                      p2 = 6.07710050630396597660e-11, p3 = 2.02226624871116645580e-21,
                      p3t = 8.47842766036889956997e-32, big = 6755399441055744.0;
  double ax = x < 0 ? -x : x, r0, r1;
  int n = 0;
  if (!(ax < 823549.6)) {
    *s = sin(x);
    *c = cos(x);
    return;
  }
  if (ax <= 0.78539816339744827900) {
    r0 = x;
    r1 = 0;
  } else {
    double fn = (x * invpio2 + big) - big;
    double a1 = fn * p1, q1 = x - a1, e1 = (x - q1) - a1;
    double a2 = fn * p2, q2 = q1 - a2, e2 = (q1 - q2) - a2;
    double a3 = fn * p3, q3 = q2 - a3, e3 = (q2 - q3) - a3;
    double w = ((fn * p3t - e3) - e2) - e1;
    r0 = q3 - w;
    r1 = (q3 - r0) - w;
    n = (int)(((long long)fn) & 3);
  }
  {
    double ks = opt_ksin(r0, r1), kc = opt_kcos(r0, r1);
    double sv = (n & 1) ? kc : ks, cv = (n & 1) ? ks : kc;
    *s = (n & 2) ? -sv : sv;
    *c = ((n + 1) & 2) ? -cv : cv;
  }
}

static double opt_sin(double x) {
  double s, c;
  opt_sincos(x, &s, &c);
  return s;
}
static double opt_cos(double x) {
  double s, c;
  opt_sincos(x, &s, &c);
  return c;
}
#endif
