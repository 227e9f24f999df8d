// exact_update.cuh -- "exact" build mode of the batched C/GMRES control update.
//
// One thread advances one controller instance by one control step, with the
// arithmetic of the reference in the reference's order:
// binary64, no FMA contraction (this header is only included from translation
// units compiled with -fmad=false), strictly sequential index-ascending sums.
// Results are bit-identical to the reference for models without libm calls.
//
// What is different from the reference is only where things live and how many
// passes over memory are made:
//   * all per-instance vectors are rows of structure-of-arrays matrices
//     [element][instance] so that the 32 lanes of a warp (32 instances) touch
//     consecutive doubles -- every global access is a fully coalesced 256 B;
//   * the Krylov basis is stored un-normalised, v_k = r_k * s_k with the scale
//     s_k = 1/||r_k|| kept in shared memory; r_k[j]*s_k is re-evaluated where
//     the reference reads v_k[j], which yields the identical double and removes
//     one read+write pass per basis vector;
//   * U + h*v (cgmres.hpp:168-169), (F - F_dxh_h)/h (cgmres.hpp:173-174),
//     b (cgmres.hpp:94-96) and r0 = b - A x0 (gmres.hpp:34) are fused into the
//     horizon sweeps; dHdu is evaluated inside the backward costate sweep
//     (cgmres.hpp:146-161 reads ltau[i+1] only), so ltau is never stored;
//   * every Gram-Schmidt axpy (gmres.hpp:56-57) is fused with the next dot
//     product (gmres.hpp:55) or with the norm (gmres.hpp:60);
//   * x += V y (gmres.hpp:110-111), U += dUdt*dt (cgmres.hpp:102-103) and the
//     Euler plant step (<example>/main.cpp:74-76) share one final pass.
#pragma once
#include <float.h>
#include <stdint.h>

#include <type_traits>

#include "cgmres_b200/models.hpp"
#include "cgmres_b200/plant.hpp"
#include "kernel_args.h"

// Tuning knobs (GPU sweep in profiles/README.md).  The register file is split per SM sub-partition
// (16,384 registers each), so resident warps per SM = 4*floor(16384/(32*regs)): 128 registers -> 16 warps,
// 129..168 -> 12 warps.  168 keeps the kernels free of spills at 12 warps (6 CTAs) per SM.
#ifndef CG_MAXREG
#define CG_MAXREG 168
#endif
#ifndef CG_B2
// (batch sizes that divide every shipped L = 300 / 150 / 75, so no pass ends in a ragged batch: GPU sweep
//  16/12/24/5 -> 15/15/25/5: msd 1.99e7 -> 2.01e7, semiactive 4.33e7 -> 4.43e7, arm 8.96e7 -> 9.31e7)
#define CG_B2 15  // elements per load batch, passes with 2 streams
#define CG_B3 15  // passes with 3+ streams
#define CG_BN 25  // norm pass (1 stream)
#define CG_BF 5   // final pass (7 streams)
#endif

namespace cgmres_b200 {
namespace exact {

template <class M>
struct Sz {
  static constexpr int nx = M::dim_x, nu = M::dim_u, np = M::dim_p, dv = M::dv;
  static constexpr int L = nu * dv, km = M::k_max;
  static constexpr int np1 = np > 0 ? np : 1;
};

// per-thread scalar workspace in shared memory, element e of thread t at sm[e*blockDim.x + t]
template <class M>
struct Ws {
  static constexpr int km = M::k_max;
  static constexpr int R = 0;                        // km*(km+1)/2 packed upper triangle, R(i,j) i<=j
  static constexpr int G = R + km * (km + 1) / 2;    // 3*km reflectors (gmres.hpp:115)
  static constexpr int RHO = G + 3 * km;             // km+1
  static constexpr int VS = RHO + km + 1;            // km+1 basis scales
  static constexpr int HC = VS + km + 1;             // km+2 current Hessenberg column
  static constexpr int COUNT = HC + km + 2;
  static __host__ __device__ constexpr int r(int i, int j) { return R + j * (j + 1) / 2 + i; }
};

enum SweepKind { SWEEP_W = 3 };

// Runs body(j0, integral_constant<int,NB>) over [0,L) in batches of B (tail batch of L%B): inside a batch the
// caller first issues all its loads and then does the (order-preserving) arithmetic, which gives the memory
// system B x streams independent requests per thread instead of one dependent load per element.
template <int L, int B, class Body>
__device__ __forceinline__ void for_batches(Body&& body) {
  constexpr int full = L / B, tail = L % B;
  for (int b = 0; b < full; b++) body(b * B, std::integral_constant<int, B>());
  if (tail > 0) body(full * B, std::integral_constant<int, (tail > 0 ? tail : 1)>());
}

// One evaluation of F (cgmres.hpp:113-162) at U + h*v_k with the Jacobian-vector post-processing fused in:
//   inputs  u = (pert*ps)*h + U   (v_k[j] = r_k[j]*(1/||r_k||), gmres.hpp:44,67; cgmres.hpp:168-169)
//   out[j]  = (F - F1[j])*inv_h   (cgmres.hpp:173-174)
// The next stage's inputs are requested before the current stage's dependent arithmetic starts.
template <class M, bool PFULL>
__device__ __forceinline__ void sweep_w(const ExactArgs& a, const int64_t n, const double* x0, const double dtau,
                                        const double* __restrict__ pert, const double ps, double* __restrict__ out,
                                        const double* pconst) {
  using S = Sz<M>;
  constexpr int nx = S::nx, nu = S::nu, np = S::np, dv = S::dv;
  const int64_t ld = a.ld;
  const double* __restrict__ U = a.U + n;
  const double* __restrict__ F1 = a.F1 + n;
  double* __restrict__ xt = a.xtau + n;
  const double* __restrict__ pt = a.ptau + n;
  constexpr double hh = M::h;
  constexpr double inv_h = 1.0 / M::h;

  auto load_raw = [&](double* uu, double* vv, int i) {
#pragma unroll
    for (int j = 0; j < nu; j++) {
      const int64_t o = (int64_t)(i * nu + j) * ld;
      uu[j] = U[o];
      vv[j] = pert[o];
    }
  };
  auto combine = [&](double* u, const double* uu, const double* vv) {
#pragma unroll
    for (int j = 0; j < nu; j++) {
      double v = vv[j] * ps;
      v = v * hh;
      u[j] = v + uu[j];
    }
  };
  auto load_p = [&](double* p, int i) {
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = PFULL ? pt[(int64_t)(i * np + j) * ld] : pconst[j];
  };

  double xc[nx], u[nu], p[S::np1], uu[nu], vv[nu];
#pragma unroll
  for (int j = 0; j < nx; j++) xc[j] = x0[j];

  // forward Euler rollout (cgmres.hpp:132-140); xtau[1..dv-1] go to scratch, xtau[0]=x0 and xtau[dv] stay in registers
  load_raw(uu, vv, 0);
  for (int i = 0; i < dv; i++) {
    double f[nx];
    combine(u, uu, vv);
    load_p(p, i);
    if (i + 1 < dv) load_raw(uu, vv, i + 1);
    M::dxdt(f, xc, u, p);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = f[j] * dtau;
      xc[j] = m + xc[j];
    }
    if (i + 1 < dv) {
#pragma unroll
      for (int j = 0; j < nx; j++) xt[(int64_t)(i * nx + j) * ld] = xc[j];
    }
  }

  // terminal costate (cgmres.hpp:145) and backward sweep (cgmres.hpp:146-153) with dHdu (cgmres.hpp:156-161) fused in
  double lmd[nx], xi[nx], f1[nu];
  load_p(p, dv);
  M::dPhidx(lmd, xc, p);
  auto load_back = [&](int i) {
    load_raw(uu, vv, i);
#pragma unroll
    for (int j = 0; j < nu; j++) f1[j] = F1[(int64_t)(i * nu + j) * ld];
    if (i > 0) {
#pragma unroll
      for (int j = 0; j < nx; j++) xi[j] = xt[(int64_t)((i - 1) * nx + j) * ld];
    } else {
#pragma unroll
      for (int j = 0; j < nx; j++) xi[j] = x0[j];
    }
  };
  load_back(dv - 1);
  for (int i = dv - 1; i >= 0; i--) {
    double xs[nx], fs[nu], hu[nu], hx[nx];
    combine(u, uu, vv);
#pragma unroll
    for (int j = 0; j < nx; j++) xs[j] = xi[j];
#pragma unroll
    for (int j = 0; j < nu; j++) fs[j] = f1[j];
    load_p(p, i);
    if (i > 0) load_back(i - 1);
    M::dHdu(hu, xs, u, p, lmd);
#pragma unroll
    for (int j = 0; j < nu; j++) {
      const double ax = hu[j] - fs[j];
      out[(int64_t)(i * nu + j) * ld] = ax * inv_h;
    }
    if (i > 0) {  // ltau[0] is never read by the reference (cgmres.hpp:160 uses ltau[i+1] only)
      M::dHdx(hx, xs, u, p, lmd);
#pragma unroll
      for (int j = 0; j < nx; j++) {
        double m = hx[j] * dtau;
        lmd[j] = m + lmd[j];
      }
    }
  }
}

// The three F evaluations that do not depend on the Krylov iteration, in ONE pair of horizon sweeps:
//   A: F(U, x+dx*h, t+h)            -> F1           (cgmres.hpp:88)
//   B: F(U, x, t)                   -> b            (cgmres.hpp:91-96), never stored
//   C: F(U+h*dUdt, x+dx*h, t+h)     -> r0 = b - Ax  (gmres.hpp:33-34, cgmres.hpp:164-175), stored to column 0
// U and dUdt are read once per direction instead of three times and b / F1 are consumed from registers.
template <class M, bool PFULL>
__device__ __forceinline__ void sweep_first3(const ExactArgs& a, const int64_t n, const double* x, const double* xh,
                                             double* __restrict__ r0out, const double* pconst, const double dtau_t,
                                             const double dtau_th) {
  using S = Sz<M>;
  constexpr int nx = S::nx, nu = S::nu, np = S::np, dv = S::dv;
  const int64_t ld = a.ld;
  const double* __restrict__ U = a.U + n;
  const double* __restrict__ dU = a.dUdt + n;
  double* __restrict__ F1 = a.F1 + n;
  const int64_t plane = (int64_t)nx * (dv > 1 ? dv - 1 : 1) * ld;
  double* __restrict__ xtA = a.xtau + n;
  double* __restrict__ xtB = xtA + plane;
  double* __restrict__ xtC = xtB + plane;
  const double* __restrict__ pt = a.ptau + n;
  const double dth = dtau_th, dt0 = dtau_t;
  constexpr double hh = M::h;
  constexpr double inv_h = 1.0 / M::h;
  constexpr double c1 = (1 - M::zeta * M::h);

  auto load_raw = [&](double* uu, double* vv, int i) {
#pragma unroll
    for (int j = 0; j < nu; j++) {
      const int64_t o = (int64_t)(i * nu + j) * ld;
      uu[j] = U[o];
      vv[j] = dU[o];
    }
  };
  auto combine = [&](double* uc, const double* uu, const double* vv) {
#pragma unroll
    for (int j = 0; j < nu; j++) {
      double v = vv[j] * hh;  // mul(U_buf, dUdt, h); add(U_buf, U_buf, U)
      uc[j] = v + uu[j];
    }
  };
  auto load_p = [&](double* p, int i) {
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = PFULL ? pt[(int64_t)(i * np + j) * ld] : pconst[j];
  };
  auto euler = [&](double* xc, const double* u, const double* p, double dtau) {
    double f[nx];
    M::dxdt(f, xc, u, p);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = f[j] * dtau;
      xc[j] = m + xc[j];
    }
  };

  double xa[nx], xb[nx], xc3[nx], ua[nu], uc[nu], uu[nu], vv[nu], p[S::np1];
#pragma unroll
  for (int j = 0; j < nx; j++) {
    xa[j] = xh[j];
    xb[j] = x[j];
    xc3[j] = xh[j];
  }
  load_raw(uu, vv, 0);
  for (int i = 0; i < dv; i++) {
#pragma unroll
    for (int j = 0; j < nu; j++) ua[j] = uu[j];
    combine(uc, uu, vv);
    load_p(p, i);
    if (i + 1 < dv) load_raw(uu, vv, i + 1);
    euler(xa, ua, p, dth);
    euler(xb, ua, p, dt0);
    euler(xc3, uc, p, dth);
    if (i + 1 < dv) {
#pragma unroll
      for (int j = 0; j < nx; j++) {
        const int64_t o = (int64_t)(i * nx + j) * ld;
        xtA[o] = xa[j];
        xtB[o] = xb[j];
        xtC[o] = xc3[j];
      }
    }
  }

  auto costate = [&](double* l, const double* xs, const double* u, const double* p, double dtau) {
    double hx[nx];
    M::dHdx(hx, xs, u, p, l);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = hx[j] * dtau;
      l[j] = m + l[j];
    }
  };
  double la[nx], lb[nx], lc[nx];
  load_p(p, dv);
  M::dPhidx(la, xa, p);
  M::dPhidx(lb, xb, p);
  M::dPhidx(lc, xc3, p);
  load_raw(uu, vv, dv - 1);
  for (int i = dv - 1; i >= 0; i--) {
    double sa[nx], sb[nx], sc[nx], ha[nu], hb[nu], hc[nu];
    if (i > 0) {  // rollout states of this stage: written a few hundred instructions ago, served by L1/L2
#pragma unroll
      for (int j = 0; j < nx; j++) {
        const int64_t o = (int64_t)((i - 1) * nx + j) * ld;
        sa[j] = xtA[o];
        sb[j] = xtB[o];
        sc[j] = xtC[o];
      }
    } else {
#pragma unroll
      for (int j = 0; j < nx; j++) {
        sa[j] = xh[j];
        sb[j] = x[j];
        sc[j] = xh[j];
      }
    }
#pragma unroll
    for (int j = 0; j < nu; j++) ua[j] = uu[j];
    combine(uc, uu, vv);
    load_p(p, i);
    if (i > 0) load_raw(uu, vv, i - 1);
    M::dHdu(ha, sa, ua, p, la);
    M::dHdu(hb, sb, ua, p, lb);
    M::dHdu(hc, sc, uc, p, lc);
#pragma unroll
    for (int j = 0; j < nu; j++) {
      const int64_t o = (int64_t)(i * nu + j) * ld;
      F1[o] = ha[j];
      double b = hb[j] * c1;  // (F*(1-zeta*h) - F1)*(1/h), three roundings (cgmres.hpp:94-96)
      b = b - ha[j];
      b = b * inv_h;
      double ax = hc[j] - ha[j];  // (F(U+h*dUdt) - F1)*(1/h) (cgmres.hpp:173-174)
      ax = ax * inv_h;
      r0out[o] = b - ax;  // gmres.hpp:34
    }
    if (i > 0) {
      costate(la, sa, ua, p, dth);
      costate(lb, sb, ua, p, dt0);
      costate(lc, sc, uc, p, dth);
    }
  }
}

// The whole update for instance n.  Returns the status word (exit path | columns used << 8).
template <class M, class Sim, bool PFULL>
__device__ __forceinline__ int control_update(const ExactArgs& a, const int64_t n, double* sm) {
  using S = Sz<M>;
  using W = Ws<M>;
  constexpr int nx = S::nx, nu = S::nu, np = S::np, L = S::L, km = S::km;
  constexpr int B2 = CG_B2, B3 = CG_B3;  // elements per load batch in passes with 2 / 3+ streams
  const int64_t ld = a.ld;
  const int bs = blockDim.x;
#define WS(e) sm[(e) * bs]

  double x[nx], pc[S::np1], u0[nu];
#pragma unroll
  for (int j = 0; j < nx; j++) x[j] = a.x[(int64_t)j * ld + n];
#pragma unroll
  for (int j = 0; j < np; j++) pc[j] = a.ptau[(int64_t)j * ld + n];  // p(t): first stage of ptau (cgmres.hpp:83)
#pragma unroll
  for (int j = 0; j < nu; j++) u0[j] = a.U[(int64_t)j * ld + n];

  // x + dxdt*h (cgmres.hpp:83-85)
  double xh[nx];
  {
    double f[nx];
    M::dxdt(f, x, u0, pc);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = f[j] * M::h;
      xh[j] = m + x[j];
    }
  }

  double* V = a.V + n;
  auto col = [&](int k) { return V + (int64_t)k * L * ld; };

  // horizon steps: batch-uniform (host libm, bit-identical to the reference) or from this instance's own clock
  double dtau_t = a.dtau_t, dtau_th = a.dtau_th;
  if (a.t_inst) {
    const double ti = a.t_inst[n];
    dtau_t = horizon_dtau<M>(ti);
    dtau_th = horizon_dtau<M>(ti + M::h);
    a.t_inst[n] = ti + M::dt;  // cgmres.hpp:107
  }

  sweep_first3<M, PFULL>(a, n, x, xh, col(0), pc, dtau_t, dtau_th);  // F1, b, r0 = b - A*dUdt in one pair of sweeps

  int code = EXIT_FULL;
  int ncol = 0;
  bool solve = true;

  // rho = ||r0|| (gmres.hpp:37, matrix.hpp:140-148)
  {
    double s = 0;
    const double* r0 = col(0);
    for_batches<L, CG_BN>([&](int j0, auto nb) {
      constexpr int NB = decltype(nb)::value;
      double v[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) v[q] = r0[(int64_t)(j0 + q) * ld];
#pragma unroll
      for (int q = 0; q < NB; q++) s += v[q] * v[q];
    });
    const double rho0 = sqrt(s);
    WS(W::RHO + 0) = rho0;
    if (rho0 < M::tol) {  // gmres.hpp:39-41: silent return, dUdt keeps its old value
      code = EXIT_RHO0;
      solve = false;
    } else {
      WS(W::VS + 0) = 1.0 / rho0;  // div(): multiply by the rounded reciprocal (matrix.hpp:122-128)
    }
  }

  if (solve) {
    int k = 0;
    for (; k < km; k++) {
      double* w = col(k + 1);
      sweep_w<M, PFULL>(a, n, xh, dtau_th, col(k), WS(W::VS + k), w, pc);  // w = A v_k (gmres.hpp:48)

      // modified Gram-Schmidt (gmres.hpp:52-58); h_ik kept in HC[i]
      {
        const double* r0 = col(0);
        const double s0 = WS(W::VS + 0);
        double acc = 0;
        for_batches<L, B2>([&](int j0, auto nb) {
          constexpr int NB = decltype(nb)::value;
          double rv[NB], wv[NB];
#pragma unroll
          for (int q = 0; q < NB; q++) {
            rv[q] = r0[(int64_t)(j0 + q) * ld];
            wv[q] = w[(int64_t)(j0 + q) * ld];
          }
#pragma unroll
          for (int q = 0; q < NB; q++) {
            const double v = rv[q] * s0;
            acc += v * wv[q];
          }
        });
        WS(W::HC + 0) = acc;
      }
      for (int i = 0; i < k; i++) {  // w -= v_i*h_i fused with h_{i+1} = <v_{i+1}, w>
        const double* ri = col(i);
        const double* rn = col(i + 1);
        const double si = WS(W::VS + i), sn = WS(W::VS + i + 1), hi = WS(W::HC + i);
        double acc = 0;
        for_batches<L, B3>([&](int j0, auto nb) {
          constexpr int NB = decltype(nb)::value;
          double av[NB], bv[NB], wv[NB];
#pragma unroll
          for (int q = 0; q < NB; q++) {
            const int64_t o = (int64_t)(j0 + q) * ld;
            av[q] = ri[o];
            bv[q] = rn[o];
            wv[q] = w[o];
          }
#pragma unroll
          for (int q = 0; q < NB; q++) {
            const double vi = av[q] * si;
            const double t = vi * hi;
            const double wj = wv[q] - t;
            w[(int64_t)(j0 + q) * ld] = wj;
            const double vn = bv[q] * sn;
            acc += vn * wj;
          }
        });
        WS(W::HC + i + 1) = acc;
      }
      double hn;
      {  // last axpy fused with the norm (gmres.hpp:59-60)
        const double* rk = col(k);
        const double sk = WS(W::VS + k), hk = WS(W::HC + k);
        double acc = 0;
        for_batches<L, B2>([&](int j0, auto nb) {
          constexpr int NB = decltype(nb)::value;
          double av[NB], wv[NB];
#pragma unroll
          for (int q = 0; q < NB; q++) {
            const int64_t o = (int64_t)(j0 + q) * ld;
            av[q] = rk[o];
            wv[q] = w[o];
          }
#pragma unroll
          for (int q = 0; q < NB; q++) {
            const double vk = av[q] * sk;
            const double t = vk * hk;
            const double wj = wv[q] - t;
            w[(int64_t)(j0 + q) * ld] = wj;
            acc += wj * wj;
          }
        });
        hn = sqrt(acc);
      }
      if (fabs(hn) < DBL_EPSILON) {  // gmres.hpp:63-65 "Breakdown": return without touching dUdt
        code = EXIT_BREAKDOWN;
        ncol = k;
        solve = false;
        break;
      }
      WS(W::VS + k + 1) = 1.0 / hn;  // gmres.hpp:67
      WS(W::HC + k + 1) = hn;

      // apply the stored reflectors to the new column (gmres.hpp:71-77)
      for (int i = 0; i < k; i++) {
        const double g0 = WS(W::G + 3 * i), g1 = WS(W::G + 3 * i + 1), g2 = WS(W::G + 3 * i + 2);
        const double ha = WS(W::HC + i), hb = WS(W::HC + i + 1);
        const double buf = (g0 * ha + g1 * hb) * g2;
        WS(W::HC + i) = ha - buf * g0;
        WS(W::HC + i + 1) = hb - buf * g1;
      }
      {  // new reflector (gmres.hpp:78-85) and residual update (gmres.hpp:88-90)
        const double ha = WS(W::HC + k), hb = WS(W::HC + k + 1);
        const double sg = (ha < 0.0) ? -1.0 : 1.0;                 // matrix.hpp:162
        const double buf = -sg * sqrt((0.0 + ha * ha) + hb * hb);  // norm(.,2), matrix.hpp:140-148
        const double g0 = ha - buf;
        const double g1 = hb;
        const double g2 = 2.0 / ((0.0 + g0 * g0) + g1 * g1);
        WS(W::G + 3 * k) = g0;
        WS(W::G + 3 * k + 1) = g1;
        WS(W::G + 3 * k + 2) = g2;
        WS(W::HC + k) = buf;
        const double rk = WS(W::RHO + k);
        const double rb = g0 * rk * g2;
        WS(W::RHO + k) = rk - rb * g0;
        WS(W::RHO + k + 1) = -rb * g1;
      }
      for (int i = 0; i <= k; i++) WS(W::r(i, k)) = WS(W::HC + i);
      if (fabs(WS(W::RHO + k + 1)) < M::tol) {  // gmres.hpp:93-95: break with k NOT incremented
        code = EXIT_CONVERGED;
        break;
      }
    }
    if (solve) ncol = k;  // == km when the loop ran to completion
  }

  if (solve) {
    // back substitution (gmres.hpp:100-107)
    for (int i = ncol - 1; i >= 0; i--) {
      double ri = WS(W::RHO + i);
      for (int j = ncol - 1; j > i; j--) ri -= WS(W::r(i, j)) * WS(W::RHO + j);
      ri /= WS(W::r(i, i));
      WS(W::RHO + i) = ri;
    }
  }

  // dUdt += V y (gmres.hpp:110-111, accumulation order of matrix.hpp:82-91), then U += dUdt*dt (cgmres.hpp:102-103)
  {
    double y[km], sc[km];
#pragma unroll
    for (int c = 0; c < km; c++) {
      y[c] = (solve && c < ncol) ? WS(W::RHO + c) : 0.0;
      sc[c] = (solve && c < ncol) ? WS(W::VS + c) : 0.0;
    }
    double* __restrict__ Up = a.U + n;
    double* __restrict__ dU = a.dUdt + n;
    constexpr int BF = CG_BF;
    for_batches<L, BF>([&](int j0, auto nb) {
      constexpr int NB = decltype(nb)::value;
      double dv_[NB], uv[NB], cv[NB][km];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int64_t o = (int64_t)(j0 + q) * ld;
        dv_[q] = dU[o];
        uv[q] = Up[o];
#pragma unroll
        for (int c = 0; c < km; c++) cv[q][c] = (solve && c < ncol) ? V[(int64_t)c * L * ld + o] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const int64_t o = (int64_t)(j0 + q) * ld;
        double d = dv_[q];
        if (solve) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < km; c++) {
            if (c < ncol) {
              const double v = cv[q][c] * sc[c];
              s += v * y[c];
            }
          }
          d = d + s;
          dU[o] = d;
        }
        const double inc = d * M::dt;
        Up[o] = uv[q] + inc;
      }
    });
  }

  // u = U[0:dim_u] (cgmres.hpp:109): re-read this thread's own stores
#pragma unroll
  for (int j = 0; j < nu; j++) {
    u0[j] = a.U[(int64_t)j * ld + n];
    a.u_out[(int64_t)j * ld + n] = u0[j];
  }

  if (a.plant) {  // x += Simulator::dxdt(x,u)*dt (<example>/main.cpp:74-76), or the RK4 option (plant.hpp)
    plant_step<Sim>(a.plant, x, u0);
#pragma unroll
    for (int j = 0; j < nx; j++) a.x[(int64_t)j * ld + n] = x[j];
  }
#undef WS
  return code | (ncol << 8);
}

template <class M, class Sim, bool PFULL>
__global__ void __maxnreg__(CG_MAXREG) control_kernel(const ExactArgs a) {
  extern __shared__ double sm_ws[];
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n) return;
  const int st = control_update<M, Sim, PFULL>(a, n, sm_ws + threadIdx.x);
  a.status[n] = st;
}

// ---- batched init_u0_newton (cgmres.hpp:61-76) with linsolve (matrix.hpp:166-224) ----
template <int N>
__device__ __forceinline__ void linsolve(double* vec, double* mat) {
  for (int k = 0; k < N - 1; k++) {
    int piv = k;
    double best = fabs(mat[N * k + k]);
    for (int i = k + 1; i < N; i++) {
      const double c = fabs(mat[N * k + i]);
      if (best < c) {
        best = c;
        piv = i;
      }
    }
    if (piv != k) {
      double s = vec[k];
      vec[k] = vec[piv];
      vec[piv] = s;
      for (int j = k; j < N; j++) {
        s = mat[N * j + k];
        mat[N * j + k] = mat[N * j + piv];
        mat[N * j + piv] = s;
      }
    }
    const double r = 1.0 / mat[N * k + k];
    for (int i = k + 1; i < N; i++) {
      mat[N * k + i] = mat[N * k + i] * r;
      for (int j = k + 1; j < N; j++) mat[N * j + i] -= mat[N * k + i] * mat[N * j + k];
      vec[i] -= mat[N * k + i] * vec[k];
    }
  }
  for (int i = N - 1; i >= 0; i--) {
    for (int j = N - 1; j > i; j--) vec[i] -= mat[N * j + i] * vec[j];
    vec[i] /= mat[N * i + i];
  }
}

// u0[n][nu] (instance-major, in/out), x0[n][nx], p0[n][p_stride] (first dim_p entries used); fills U with the
// result: element e of instance n at U[e*es + n*is]
template <class M>
__global__ void newton_init_kernel(int64_t n_inst, int64_t es, int64_t is, double* __restrict__ u0,
                                   const double* __restrict__ x0, const double* __restrict__ p0, int p_stride,
                                   int n_loop, double* __restrict__ U) {
  using S = Sz<M>;
  constexpr int nx = S::nx, nu = S::nu, np = S::np, dv = S::dv;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_inst) return;
  double x[nx], u[nu], p[S::np1], lmd[nx], vec[nu], mat[nu * nu];
  for (int j = 0; j < nx; j++) x[j] = x0[n * nx + j];
  for (int j = 0; j < nu; j++) u[j] = u0[n * nu + j];
  for (int j = 0; j < np; j++) p[j] = p0[n * p_stride + j];
  M::dPhidx(lmd, x, p);
  for (int it = 0; it < n_loop; it++) {
    M::dHdu(vec, x, u, p, lmd);
    M::ddHduu(mat, x, u, p, lmd);
    linsolve<nu>(vec, mat);
    for (int j = 0; j < nu; j++) u[j] = u[j] - vec[j];
  }
  for (int j = 0; j < nu; j++) u0[n * nu + j] = u[j];
  for (int i = 0; i < dv; i++)
    for (int j = 0; j < nu; j++) U[(int64_t)(i * nu + j) * es + n * is] = u[j];
}

}  // namespace exact
}  // namespace cgmres_b200
