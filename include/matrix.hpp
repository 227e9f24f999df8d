// matrix.hpp -- host-side stand-in for the reference's include/matrix.hpp (include/matrix.hpp:10-224).
//
// In the reference these free functions ARE the numerical kernels of the controller.  Here the controller's
// arithmetic lives in the sm_100a kernels (fixed-size register / shared-memory tiles fused into the control
// update), so nothing in libcgmres_b200.so calls this header.  It exists for source compatibility only: the
// shipped example mains use `mul` and `add` for the Euler plant step
// (<example>/main.cpp:74-76, e.g. mass_spring_damper/main.cpp:75-76), and user code written against the
// reference may use the other small helpers on its own host-side vectors.
//
// Same names, argument order and int16_t sizes as the reference; same operation order where the order is
// observable (sequential left-to-right sums in norm/dot, reciprocal-then-multiply in div, column-major
// mat*vec accumulated column by column, first-maximum partial pivoting in linsolve), so a host program that
// mixes these helpers with Cgmres<Model> produces the numbers it produced with the reference.
// The unused mat*mat overload (matrix.hpp:95-119) and the DEBUG_MODE alias aborts are not reproduced.
#pragma once
#include <math.h>
#include <stdint.h>

namespace cgmres_b200 {
namespace hostvec {
// ret[i] = op(a[i], b[i]) / op(a[i], c) over `count` entries
template <class Op>
inline void zip(double* ret, const double* a, const double* b, const int count, Op op) {
  for (int i = 0; i < count; i++) ret[i] = op(a[i], b[i]);
}
template <class Op>
inline void map(double* ret, const double* a, const int count, Op op) {
  for (int i = 0; i < count; i++) ret[i] = op(a[i]);
}
}  // namespace hostvec
}  // namespace cgmres_b200

// ret = vec / ret = mat                                                   (matrix.hpp:10-23)
inline void mov(double* ret, const double* vec, const int16_t row) {
  cgmres_b200::hostvec::map(ret, vec, row, [](double v) { return v; });
}
inline void mov(double* ret, const double* mat, const int16_t row, const int16_t col) {
  cgmres_b200::hostvec::map(ret, mat, (int)row * col, [](double v) { return v; });
}

// ret = a + b                                                             (matrix.hpp:26-39)
inline void add(double* ret, const double* vec1, const double* vec2, const int16_t row) {
  cgmres_b200::hostvec::zip(ret, vec1, vec2, row, [](double a, double b) { return a + b; });
}
inline void add(double* ret, const double* mat1, const double* mat2, const int16_t row, const int16_t col) {
  cgmres_b200::hostvec::zip(ret, mat1, mat2, (int)row * col, [](double a, double b) { return a + b; });
}

// ret = a - b                                                             (matrix.hpp:42-55)
inline void sub(double* ret, const double* vec1, const double* vec2, const int16_t row) {
  cgmres_b200::hostvec::zip(ret, vec1, vec2, row, [](double a, double b) { return a - b; });
}
inline void sub(double* ret, const double* mat1, const double* mat2, const int16_t row, const int16_t col) {
  cgmres_b200::hostvec::zip(ret, mat1, mat2, (int)row * col, [](double a, double b) { return a - b; });
}

// ret = a * c                                                             (matrix.hpp:58-71)
inline void mul(double* ret, const double* vec, const double c, const int16_t row) {
  cgmres_b200::hostvec::map(ret, vec, row, [c](double v) { return v * c; });
}
inline void mul(double* ret, const double* mat, const double c, const int16_t row, const int16_t col) {
  cgmres_b200::hostvec::map(ret, mat, (int)row * col, [c](double v) { return v * c; });
}

// ret = mat * vec, mat column major (entry (i,j) at row*j + i); the sum runs over the columns in order,
// starting from 0                                                         (matrix.hpp:74-92)
inline void mul(double* ret, const double* mat, const double* vec, const int16_t row, const int16_t col) {
  for (int i = 0; i < row; i++) ret[i] = 0.0;
  for (int j = 0; j < col; j++) {
    const double* column = mat + (int)row * j;
    const double scale = vec[j];
    for (int i = 0; i < row; i++) ret[i] += column[i] * scale;
  }
}

// ret = a / c, evaluated as a * (1.0 / c) like the reference              (matrix.hpp:122-137)
inline void div(double* ret, const double* vec, const double c, const int16_t row) {
  const double inv_c = 1.0 / c;
  cgmres_b200::hostvec::map(ret, vec, row, [inv_c](double v) { return v * inv_c; });
}
inline void div(double* ret, const double* mat, const double c, const int16_t row, const int16_t col) {
  const double inv_c = 1.0 / c;
  cgmres_b200::hostvec::map(ret, mat, (int)row * col, [inv_c](double v) { return v * inv_c; });
}

// vec1' * vec2 and ||vec||: index-ordered sums from 0                     (matrix.hpp:140-159)
inline double dot(const double* vec1, const double* vec2, const int16_t n) {
  double acc = 0;
  for (int i = 0; i < n; i++) acc += vec1[i] * vec2[i];
  return acc;
}
inline double norm(const double* vec, int16_t n) { return sqrt(dot(vec, vec, n)); }

// sign(0) = +1                                                            (matrix.hpp:162)
inline double sign(const double x) { return (x < 0.0) ? -1.0 : 1.0; }

// vec <- mat \ vec by Gaussian elimination with partial pivoting; mat column major n x n, both overwritten.
// Pivot: the first row of maximal |entry| in the column; multipliers via the rounded reciprocal of the pivot;
// true division in the back substitution                                  (matrix.hpp:166-224)
inline void linsolve(double* vec, double* mat, const int16_t n) {
  auto at = [mat, n](int r, int c) -> double& { return mat[(int)n * c + r]; };
  for (int k = 0; k + 1 < n; k++) {
    int piv = k;
    double best = fabs(at(k, k));
    for (int r = k + 1; r < n; r++) {
      const double cand = fabs(at(r, k));
      if (best < cand) {
        best = cand;
        piv = r;
      }
    }
    if (piv != k) {
      double t = vec[k];
      vec[k] = vec[piv];
      vec[piv] = t;
      for (int c = k; c < n; c++) {
        t = at(k, c);
        at(k, c) = at(piv, c);
        at(piv, c) = t;
      }
    }
    const double inv_pivot = 1.0 / at(k, k);
    for (int r = k + 1; r < n; r++) {
      const double m = at(r, k) * inv_pivot;
      at(r, k) = m;
      for (int c = k + 1; c < n; c++) at(r, c) -= m * at(k, c);
      vec[r] -= m * vec[k];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    for (int c = n - 1; c > r; c--) vec[r] -= at(r, c) * vec[c];
    vec[r] /= at(r, r);
  }
}
