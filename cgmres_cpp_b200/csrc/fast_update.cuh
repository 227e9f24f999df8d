// fast_update.cuh -- "fast" build mode: the whole control update of a group of G instances runs out of
// shared memory and registers of ONE CTA; HBM only sees the algorithmic minimum (read U, dUdt, x, p; write
// U, dUdt, x, u: 9.7 KB per mass_spring_damper update instead of the ~250 KB the streaming exact kernel moves).
//
// Mapping (north star items 2-4):
//   * warp g of the CTA owns instance g for all vector work: lane l holds elements l, l+32, ... of the length-L
//     vectors (the working vector w stays in registers through a whole Gram-Schmidt sweep), dot products and
//     norms are butterfly shuffle reductions, the Hessenberg / Householder / back-substitution scalars
//     (gmres.hpp:71-107) are warp-uniform registers + a few shared-memory words.
//   * the horizon sweeps (cgmres.hpp:113-162) are serial over dv and only dim_x wide, and lanes of a warp
//     cannot run different formulas without divergence, so they are TRANSPOSED: in the sweep phases lane l of
//     the first warp(s) runs the rollout + costate + dHdu of instance l (three trajectories per instance for
//     the fused F(U,x+dx*h,t+h) / F(U,x,t) / F(U+h*dUdt,x+dx*h,t+h) evaluation) straight out of the
//     instances' shared-memory blocks; the per-instance block stride is odd in 8-byte words so those
//     lane-per-instance accesses are bank-conflict free.
//   * Krylov basis (gmres.hpp:11), U, F_dxh_h, the U+h*v buffer and the rollout states never leave the SM.
//
// Arithmetic: FMA contraction on (this header is compiled without -fmad=false), tree-ordered reductions.
// Same algorithm, different rounding: judged to the north-star tolerances (1e-9 per update, 1e-6 closed loop),
// not bit for bit.  EXACT_SUMS=true (debug / verification build of the same kernel inside the -fmad=false
// translation unit) replaces the shuffle reductions by the reference's sequential sums, done lane-per-instance
// by the first warp, and is bit-identical to the reference.
#pragma once
#include <float.h>
#include <stdint.h>

#include "cgmres_b200/models.hpp"
#include "kernel_args.h"

namespace cgmres_b200 {
namespace fast {

constexpr int kSmemBudget = 227 * 1024;

template <class M>
struct Lay {
  static constexpr int nx = M::dim_x, nu = M::dim_u, np = M::dim_p, dv = M::dv, km = M::k_max;
  static constexpr int L = nu * dv;
  static constexpr int np1 = np > 0 ? np : 1;
  static constexpr int XT = nx * (dv > 1 ? dv - 1 : 1);  // stored rollout states xtau[1..dv-1]
  static constexpr int Q = (L + 31) / 32;                // vector elements per lane
  // per-instance shared-memory block, offsets in doubles
  static constexpr int oU = 0;            // U                      (cgmres.hpp:196)
  static constexpr int oF1 = oU + L;      // F(U, x+dx*h, t+h)      (cgmres.hpp:202)
  static constexpr int oX = oF1 + L;      // U + h*v  ->  w         (cgmres.hpp:168-174), products in EXACT_SUMS
  static constexpr int oV = oX + L;       // km basis columns r_0..r_{km-1}, un-normalised (gmres.hpp:11)
  static constexpr int oXT = oV + km * L; // rollout states xtau[1..dv-1] of the Arnoldi sweeps / trajectory A
  // The fused first evaluation runs three trajectories; B and C park their rollout states in basis columns
  // km-1 and 0 (both unused until r0 exists) when a plane fits in a column, else in two extra planes.
  static constexpr bool xt_alias = (XT <= L) && (km >= 5);
  static constexpr int oXTB = xt_alias ? oV + (km - 1) * L : oXT + XT;
  static constexpr int oXTC = xt_alias ? oV : oXT + 2 * XT;
  static constexpr int oS = oXT + (xt_alias ? 1 : 3) * XT; // scalars
  // scalar slots
  static constexpr int sR = 0;                         // packed upper triangle R(i,j), i<=j<km
  static constexpr int sG = sR + km * (km + 1) / 2;    // 3*km reflectors
  static constexpr int sRHO = sG + 3 * km;             // km+1
  static constexpr int sVS = sRHO + km + 1;            // km+1 basis scales 1/||r_k||
  static constexpr int sHC = sVS + km + 1;             // km+2 current Hessenberg column
  static constexpr int sX = sHC + km + 2;              // x
  static constexpr int sXH = sX + nx;                  // x + dxdt*h
  static constexpr int sP = sXH + nx;                  // p(t) (repeat mode) / first stage
  static constexpr int sRED = sP + np1;                // reduction result slot (EXACT_SUMS)
  static constexpr int sFLAG = sRED + 1;               // state: 0 solving, else finished (as double)
  static constexpr int sCount = sFLAG + 1;
  static constexpr int raw = oS + sCount;
  // odd number of 8-byte words per instance => lane-per-instance accesses hit distinct banks
  static constexpr int stride = (raw % 2 == 0) ? raw + 1 : raw;
  static constexpr int G_fit = kSmemBudget / (stride * 8);
  // instances per CTA (one warp each): what fits in shared memory, capped at 12 warps so that every thread
  // can keep ~168 registers (the register file, not shared memory, bounds the small-L models)
  static constexpr int G = G_fit > 12 ? 12 : G_fit;
  static constexpr int threads = 32 * G;
  static constexpr size_t smem_bytes = (size_t)G * stride * 8;
  static __host__ __device__ constexpr int r(int i, int j) { return sR + j * (j + 1) / 2 + i; }
};

// One horizon sweep for one instance, executed by ONE lane (cgmres.hpp:113-162).
//   in[]  : stage inputs u_i (shared memory, L doubles)     out[]: dHdu per stage (may alias in[]: u_i is read
//   xt[]  : rollout scratch plane                                   before out_i is written, same lane)
//   MODE 0: out = F                                   (first three evaluations)
//   MODE 1: out = (F - F1)*inv_h                      (Jacobian-vector product, cgmres.hpp:173-174)
template <class M, int MODE>
__device__ __forceinline__ void lane_sweep(const double* in, double* out, const double* f1, double* xt,
                                           const double* x0, const double dtau, const double* pconst,
                                           const double* pfull /* global [(dv+1)*np] or null */) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  constexpr double inv_h = 1.0 / M::h;
  double xc[nx], u[nu], p[Y::np1];
  auto load_p = [&](int i) {
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull ? pfull[i * np + j] : pconst[j];
  };
#pragma unroll
  for (int j = 0; j < nx; j++) xc[j] = x0[j];
  for (int i = 0; i < dv; i++) {  // forward Euler rollout, cgmres.hpp:132-140
    double f[nx];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * nu + j];
    load_p(i);
    M::dxdt(f, xc, u, p);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = f[j] * dtau;
      xc[j] = m + xc[j];
    }
    if (i + 1 < dv) {
#pragma unroll
      for (int j = 0; j < nx; j++) xt[i * nx + j] = xc[j];
    }
  }
  double lmd[nx];
  load_p(dv);
  M::dPhidx(lmd, xc, p);  // cgmres.hpp:145
  for (int i = dv - 1; i >= 0; i--) {  // costate sweep with dHdu fused, cgmres.hpp:146-161
    double xi[nx], hu[nu], hx[nx];
    if (i > 0) {
#pragma unroll
      for (int j = 0; j < nx; j++) xi[j] = xt[(i - 1) * nx + j];
    } else {
#pragma unroll
      for (int j = 0; j < nx; j++) xi[j] = x0[j];
    }
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * nu + j];
    load_p(i);
    M::dHdu(hu, xi, u, p, lmd);
#pragma unroll
    for (int j = 0; j < nu; j++) {
      if (MODE == 0) {
        out[i * nu + j] = hu[j];
      } else {
        double ax = hu[j] - f1[i * nu + j];
        out[i * nu + j] = ax * inv_h;
      }
    }
    if (i > 0) {
      M::dHdx(hx, xi, u, p, lmd);
#pragma unroll
      for (int j = 0; j < nx; j++) {
        double m = hx[j] * dtau;
        lmd[j] = m + lmd[j];
      }
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The kernel.  grid = ceil(n / G) CTAs of 32*G threads; dynamic shared memory = Lay<M>::smem_bytes.
template <class M, class Sim, bool PFULL, bool EXACT_SUMS>
__global__ void __launch_bounds__(Lay<M>::threads, 1) control_kernel(const FastArgs a) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, L = Y::L, km = Y::km, Q = Y::Q, G = Y::G;
  constexpr double hh = M::h;
  constexpr double inv_h = 1.0 / M::h;
  constexpr double c1 = (1 - M::zeta * M::h);
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * G;
  const int n_here = (int)((a.n - n0) < (int64_t)G ? (a.n - n0) : (int64_t)G);
  const bool has = wid < n_here;            // this warp owns a live instance
  const int64_t n = n0 + wid;               // its global index
  double* const blk = sm + (size_t)wid * Y::stride;
  double* const sc = blk + Y::oS;
  auto col = [&](int k) { return blk + Y::oV + k * L; };
  auto inst_blk = [&](int g) { return sm + (size_t)g * Y::stride; };

  // ---- phase 0: state in.  U -> smem, X = U + h*dUdt (input of the third trajectory), x, p(t) -------------
  double dU[Q];  // this lane's slice of dUdt stays in registers until the final update
  if (has) {
    const double* Ug = a.U + n * (int64_t)L;
    const double* dUg = a.dUdt + n * (int64_t)L;
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      if (j < L) {
        const double uu = Ug[j];
        dU[q] = dUg[j];
        blk[Y::oU + j] = uu;
        double v = dU[q] * hh;  // cgmres.hpp:168-169
        blk[Y::oX + j] = v + uu;
      }
    }
    if (lane < nx) sc[Y::sX + lane] = a.x[n * nx + lane];
    if (lane < np) sc[Y::sP + lane] = a.ptau[n * (int64_t)(PFULL ? (M::dv + 1) * np : np) + lane];
    if (lane == 0) sc[Y::sFLAG] = 0.0;
  }
  __syncthreads();
  if (has && lane == 0) {  // x + dxdt*h, cgmres.hpp:83-85
    double x[nx], u0[nu], p0[Y::np1], f[nx];
#pragma unroll
    for (int j = 0; j < nx; j++) x[j] = sc[Y::sX + j];
#pragma unroll
    for (int j = 0; j < nu; j++) u0[j] = blk[Y::oU + j];
#pragma unroll
    for (int j = 0; j < np; j++) p0[j] = sc[Y::sP + j];
    M::dxdt(f, x, u0, p0);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = f[j] * hh;
      sc[Y::sXH + j] = m + x[j];
    }
  }
  __syncthreads();

  // ---- phase 1: the three Krylov-independent F evaluations, one lane per (instance, trajectory) -----------
  //   A: F(U, x+dx*h, t+h) -> col 1     B: F(U, x, t) -> col 2     C: F(U+h*dUdt, x+dx*h, t+h) -> col 3
  if (threadIdx.x < 3 * n_here) {
    const int g = threadIdx.x / 3, tr = threadIdx.x % 3;
    double* b = inst_blk(g);
    const double* s = b + Y::oS;
    const double* pf = PFULL ? a.ptau + (n0 + g) * (int64_t)((M::dv + 1) * np) : nullptr;
    double* plane = b + (tr == 0 ? Y::oXT : (tr == 1 ? Y::oXTB : Y::oXTC));
    lane_sweep<M, 0>(tr == 2 ? b + Y::oX : b + Y::oU, b + Y::oV + (1 + tr) * L, nullptr, plane,
                     tr == 1 ? s + Y::sX : s + Y::sXH, tr == 1 ? a.dtau_t : a.dtau_th, s + Y::sP, pf);
  }
  __syncthreads();

  // ---- phase 2: F1, b, r0 = b - A*dUdt; rho0 ---------------------------------------------------------------
  double w[Q];  // working vector slice (r0, then each new Krylov vector)
  double ssq = 0.0;
  if (has) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      w[q] = 0.0;
      if (j < L) {
        const double fa = col(1)[j], fb = col(2)[j], fc = col(3)[j];
        blk[Y::oF1 + j] = fa;
        double b = fb * c1;  // cgmres.hpp:94-96
        b = b - fa;
        b = b * inv_h;
        double ax = fc - fa;  // cgmres.hpp:173-174
        ax = ax * inv_h;
        w[q] = b - ax;  // gmres.hpp:34
        col(0)[j] = w[q];
        if (EXACT_SUMS)
          blk[Y::oX + j] = w[q] * w[q];
        else
          ssq += w[q] * w[q];
      }
    }
  }
  // reductions: FAST = butterfly inside the owning warp; EXACT = lane-per-instance sequential sums by warp 0
  auto reduce = [&](double partial) -> double {
    if (!EXACT_SUMS) return warp_sum(partial);
    __syncthreads();
    if (threadIdx.x < n_here) {
      double* b = inst_blk(threadIdx.x);
      double s = 0;
      for (int j = 0; j < L; j++) s += b[Y::oX + j];
      b[Y::oS + Y::sRED] = s;
    }
    __syncthreads();
    return has ? sc[Y::sRED] : 0.0;
  };

  int code = EXIT_FULL, ncol = 0;
  bool solving = has;
  double rho[km + 1], vs[km + 1];
#pragma unroll
  for (int i = 0; i <= km; i++) {
    rho[i] = 0.0;
    vs[i] = 0.0;
  }
  {
    const double rho0 = sqrt(reduce(ssq));  // gmres.hpp:37
    rho[0] = rho0;
    if (solving) {
      if (rho0 < M::tol) {  // gmres.hpp:39-41
        code = EXIT_RHO0;
        solving = false;
      } else {
        vs[0] = 1.0 / rho0;  // gmres.hpp:44
      }
    }
  }
  if (has && lane == 0) sc[Y::sFLAG] = solving ? 0.0 : 1.0;

  // ---- Arnoldi iterations ------------------------------------------------------------------------------------
  double R[km * (km + 1) / 2], gv[3 * km];
#pragma unroll
  for (int i = 0; i < km * (km + 1) / 2; i++) R[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3 * km; i++) gv[i] = 0.0;

#pragma unroll
  for (int k = 0; k < km; k++) {
    // X = U + h*v_k, v_k = r_k*s_k (cgmres.hpp:168-169); w currently holds r_k
    if (solving) {
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) {
          double v = w[q] * vs[k];
          v = v * hh;
          blk[Y::oX + j] = v + blk[Y::oU + j];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < n_here) {  // transposed sweep: lane = instance; w = A v_k lands in X (gmres.hpp:48)
      double* b = inst_blk(threadIdx.x);
      const double* s = b + Y::oS;
      if (s[Y::sFLAG] == 0.0) {
        const double* pf = PFULL ? a.ptau + (n0 + threadIdx.x) * (int64_t)((M::dv + 1) * np) : nullptr;
        lane_sweep<M, 1>(b + Y::oX, b + Y::oX, b + Y::oF1, b + Y::oXT, s + Y::sXH, a.dtau_th, s + Y::sP, pf);
      }
    }
    __syncthreads();
    if (solving) {
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        w[q] = (j < L) ? blk[Y::oX + j] : 0.0;
      }
    }
    // modified Gram-Schmidt (gmres.hpp:52-58) against r_i*s_i, i = 0..k
    double hc[km + 2];
#pragma unroll
    for (int i = 0; i < km + 2; i++) hc[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= k; i++) {
      double part = 0.0;
      if (solving) {
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          if (j < L) {
            const double v = col(i)[j] * vs[i];
            if (EXACT_SUMS)
              blk[Y::oX + j] = v * w[q];
            else
              part += v * w[q];
          }
        }
      }
      const double hik = reduce(part);
      hc[i] = hik;
      if (solving) {
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          if (j < L) {
            const double v = col(i)[j] * vs[i];
            const double t = v * hik;
            w[q] = w[q] - t;
          }
        }
      }
    }
    double part = 0.0;
    if (solving) {
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) {
          if (EXACT_SUMS)
            blk[Y::oX + j] = w[q] * w[q];
          else
            part += w[q] * w[q];
        }
      }
    }
    const double hn = sqrt(reduce(part));  // gmres.hpp:59-60
    if (solving) {
      if (fabs(hn) < DBL_EPSILON) {  // gmres.hpp:63-65
        code = EXIT_BREAKDOWN;
        ncol = k;
        solving = false;
      }
    }
    if (solving) {
      vs[k + 1] = 1.0 / hn;  // gmres.hpp:67
      hc[k + 1] = hn;
      if (k + 1 < km) {  // the last vector only contributes its Hessenberg column
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          if (j < L) col(k + 1)[j] = w[q];
        }
      }
      // stored reflectors on the new column (gmres.hpp:71-77), new reflector (78-85), residual (88-90)
#pragma unroll
      for (int i = 0; i < k; i++) {
        const double g0 = gv[3 * i], g1 = gv[3 * i + 1], g2 = gv[3 * i + 2];
        const double buf = (g0 * hc[i] + g1 * hc[i + 1]) * g2;
        hc[i] = hc[i] - buf * g0;
        hc[i + 1] = hc[i + 1] - buf * g1;
      }
      {
        const double ha = hc[k], hb = hc[k + 1];
        const double sg = (ha < 0.0) ? -1.0 : 1.0;
        const double buf = -sg * sqrt((0.0 + ha * ha) + hb * hb);
        const double g0 = ha - buf, g1 = hb;
        const double g2 = 2.0 / ((0.0 + g0 * g0) + g1 * g1);
        gv[3 * k] = g0;
        gv[3 * k + 1] = g1;
        gv[3 * k + 2] = g2;
        hc[k] = buf;
        const double rb = g0 * rho[k] * g2;
        rho[k] = rho[k] - rb * g0;
        rho[k + 1] = -rb * g1;
      }
#pragma unroll
      for (int i = 0; i <= k; i++) R[Y::r(i, k)] = hc[i];
      ncol = k + 1;
      if (fabs(rho[k + 1]) < M::tol) {  // gmres.hpp:93-95: break with k not incremented
        code = EXIT_CONVERGED;
        ncol = k;
        solving = false;
      }
    }
    if (has && lane == 0) sc[Y::sFLAG] = solving ? 0.0 : 1.0;
  }

  // ---- back substitution (gmres.hpp:100-107), dUdt += V y (110-111), U += dUdt*dt (cgmres.hpp:102-103) ------
  const bool apply = has && (code == EXIT_FULL || code == EXIT_CONVERGED);
  if (apply) {
#pragma unroll
    for (int i = km - 1; i >= 0; i--) {
      if (i < ncol) {
        double ri = rho[i];
#pragma unroll
        for (int j = km - 1; j > i; j--)
          if (j < ncol) ri -= R[Y::r(i, j)] * rho[j];
        ri /= R[Y::r(i, i)];
        rho[i] = ri;
      }
    }
  }
  if (has) {
    double* Ug = a.U + n * (int64_t)L;
    double* dUg = a.dUdt + n * (int64_t)L;
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      if (j < L) {
        double d = dU[q];
        if (apply) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < km; c++) {
            if (c < ncol) {
              const double v = col(c)[j] * vs[c];
              s += v * rho[c];
            }
          }
          d = d + s;
          dUg[j] = d;
        }
        const double inc = d * M::dt;
        const double un = blk[Y::oU + j] + inc;
        Ug[j] = un;
        if (j < nu) blk[Y::oU + j] = un;  // keep u = U[0:dim_u] for the epilogue
      }
    }
    __syncwarp();
    if (lane == 0) {
      double x[nx], u0[nu];
#pragma unroll
      for (int j = 0; j < nu; j++) {
        u0[j] = blk[Y::oU + j];
        a.u_out[n * nu + j] = u0[j];  // cgmres.hpp:109
      }
      if (a.plant) {  // <example>/main.cpp:74-76
        double f[nx];
#pragma unroll
        for (int j = 0; j < nx; j++) x[j] = sc[Y::sX + j];
        Sim::dxdt(f, x, u0);
#pragma unroll
        for (int j = 0; j < nx; j++) {
          double m = f[j] * Sim::dt;
          a.x[n * nx + j] = x[j] + m;
        }
      }
      a.status[n] = code | (ncol << 8);
    }
  }
}

}  // namespace fast
}  // namespace cgmres_b200
