// portable_trig.hpp -- sin/cos built from +,-,* only, so that host (gcc -ffp-contract=off) and device
// (nvcc -fmad=false) produce the SAME double for the same argument.
//
// Why: arm_type_inverted_pendulum/model.hpp calls libm sin/cos.  CUDA's and glibc's implementations are both
// accurate to < 1 ulp but round differently on a few % of arguments, and the pendulum swing-up amplifies a 1-ulp
// difference past 1e-6 within ~1100 closed-loop steps.  With this header the exact build
// modes are bit-reproducible for the arm model too: the CPU checker of the test suite has its own C restatement
// of the same algorithm and the GPU must match it bit for bit; the distance between this implementation and
// glibc's (<= 1 ulp, measured in tests/test_capi_load.py) is what remains between the GPU and the
// reference-with-glibc, and is held to the north-star tolerances over 1000 steps.
//
// Algorithm: the classic fdlibm scheme (Sun Microsystems, "freely granted" licence): Cody-Waite reduction by
// pi/2 with a 3 x 33-bit split and a two-term remainder (valid for |x| < 2^19 * pi/2; larger arguments fall back to
// libm, outside every operating range of the models), then the degree-13 / degree-14 minimax kernels on
// [-pi/4, pi/4] with the remainder's tail folded in.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CGMRES_TRIG_HD __host__ __device__ __forceinline__
#else
#define CGMRES_TRIG_HD inline
#endif

namespace cgmres_b200 {
namespace ptrig {

// kernel sine on [-pi/4, pi/4]: x + tail y
CGMRES_TRIG_HD double ksin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

// kernel cosine on [-pi/4, pi/4]: x + tail y
CGMRES_TRIG_HD double kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * r - x * y));
}

// x = n*(pi/2) + (r0 + r1), |r0 + r1| <= pi/4 (+ a hair); returns n mod 4.  ok=false: |x| too large / not finite.
CGMRES_TRIG_HD int reduce(double x, double* r0, double* r1, bool* ok) {
  const double invpio2 = 6.36619772367581382433e-01;
  const double p1 = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double p2 = 6.07710050630396597660e-11;  // next 33 bits
  const double p3 = 2.02226624871116645580e-21;  // next 33 bits
  const double p3t = 8.47842766036889956997e-32; // pi/2 - (p1 + p2 + p3)
  const double ax = x < 0 ? -x : x;
  *ok = ax < 823549.6;  // 2^19 * pi/2
  if (!*ok) {
    *r0 = 0;
    *r1 = 0;
    return 0;
  }
  if (ax <= 0.78539816339744827900) {  // already in [-pi/4, pi/4]
    *r0 = x;
    *r1 = 0;
    return 0;
  }
  // fn = rint(x * 2/pi) without rounding-mode intrinsics: add and subtract 1.5 * 2^52
  const double big = 6755399441055744.0;
  const double fn = (x * invpio2 + big) - big;
  // three subtractions of fn times a 33-bit piece of pi/2 (each product is exact); the rounding error of every
  // subtraction is recovered exactly (e_i) and folded into the tail, always all three steps: one fixed
  // instruction sequence on host and device
  const double a1 = fn * p1;
  const double r1s = x - a1;
  const double e1 = (x - r1s) - a1;
  const double a2 = fn * p2;
  const double r2s = r1s - a2;
  const double e2 = (r1s - r2s) - a2;
  const double a3 = fn * p3;
  const double r3s = r2s - a3;
  const double e3 = (r2s - r3s) - a3;
  const double w = ((fn * p3t - e3) - e2) - e1;  // what is still to be subtracted from r3s
  const double y0 = r3s - w;
  *r0 = y0;
  *r1 = (r3s - y0) - w;
  // fn is an integer of magnitude < 2^20: exact conversion
  const long long n = (long long)fn;
  return (int)(n & 3);
}

// sin and cos of x from one reduction; branch-free after the range check (both kernels are evaluated and the
// quadrant only selects and negates), so divergent quadrants inside a warp cost nothing extra
CGMRES_TRIG_HD void psincos(double x, double* s, double* c) {
  double r0, r1;
  bool ok;
  const int n = reduce(x, &r0, &r1, &ok);
  if (!ok) {
    *s = ::sin(x);
    *c = ::cos(x);
    return;
  }
  const double ks = ksin(r0, r1), kc = kcos(r0, r1);
  const double sv = (n & 1) ? kc : ks;  // n = 0: sin, 1: cos, 2: -sin, 3: -cos
  const double cv = (n & 1) ? ks : kc;  // n = 0: cos, 1: -sin, 2: -cos, 3: sin
  *s = (n & 2) ? -sv : sv;
  *c = ((n + 1) & 2) ? -cv : cv;
}

CGMRES_TRIG_HD double psin(double x) {
  double s, c;
  psincos(x, &s, &c);
  return s;
}

CGMRES_TRIG_HD double pcos(double x) {
  double s, c;
  psincos(x, &s, &c);
  return c;
}

}  // namespace ptrig
}  // namespace cgmres_b200
