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

constexpr int kSmemBudget = 227 * 1024 - 2048;  // leave the per-CTA reserved shared memory of up to 2 CTAs
#ifndef CG_FAST_GCAP
#define CG_FAST_GCAP 12  // max instances (= warps) per CTA
#endif
#ifndef CG_FAST_WARPS
#define CG_FAST_WARPS 16  // max resident warps (= instances) per SM: 16 warps leave 128 registers per thread
#endif

template <class M>
struct Lay {
  static constexpr int nx = M::dim_x, nu = M::dim_u, np = M::dim_p, dv = M::dv, km = M::k_max;
  static constexpr int L = nu * dv;
  static constexpr int np1 = np > 0 ? np : 1;
  static constexpr int XT = nx * (dv > 1 ? dv - 1 : 1);  // stored rollout states xtau[1..dv-1]
  static constexpr int Q = (L + 31) / 32;                // vector elements per lane
  // per-instance shared-memory block, offsets in doubles
  // (U itself is not kept on chip: it is re-read from global memory -- an L2 hit, the CTA touched it a few
  //  microseconds earlier -- where U + h*v is formed and in the final update.)
  static constexpr int oF1 = 0;           // U during the first evaluation, then F(U, x+dx*h, t+h) (cgmres.hpp:202)
  static constexpr int oX = oF1 + L;      // U + h*v  ->  w         (cgmres.hpp:168-174), products in EXACT_SUMS
  static constexpr int oV = oX + L;       // km basis columns r_0..r_{km-1}, un-normalised (gmres.hpp:11)
  static constexpr int oXT = oV + km * L; // rollout states xtau[1..dv-1] of the Arnoldi sweeps / trajectory A
  // The fused first evaluation runs three trajectories; B and C park their rollout states in basis columns
  // km-1 and 0 (both unused until r0 exists) when a plane fits in a column, else in two extra planes.
  static constexpr bool xt_alias = (XT <= L) && (km >= 5);
  static constexpr int oXTB = xt_alias ? oV + (km - 1) * L : oXT + XT;
  static constexpr int oXTC = xt_alias ? oV : oXT + 2 * XT;
  static constexpr int oLT = oXT + (xt_alias ? 1 : 3) * XT;  // costates ltau[1..dv] of the Arnoldi sweeps
  static constexpr int oS = oLT + nx * dv;                   // scalars
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
  static constexpr int G = G_fit > CG_FAST_GCAP ? CG_FAST_GCAP : G_fit;
  // co-resident CTAs (they run out of phase, which overlaps one CTA's serial sweep with another's vector work):
  // bounded by shared memory (G_fit instances per SM) and by CG_FAST_WARPS warps per SM (register budget)
  static constexpr int inst_per_sm = G_fit > CG_FAST_WARPS ? CG_FAST_WARPS : G_fit;
  static constexpr int ctas_per_sm = (inst_per_sm / G) < 1 ? 1 : (inst_per_sm / G);
  static constexpr int threads = 32 * G;
  static constexpr size_t smem_bytes = (size_t)G * stride * 8;
  static __host__ __device__ constexpr int r(int i, int j) { return sR + j * (j + 1) / 2 + i; }
};

// Forward Euler rollout of one instance by ONE lane (cgmres.hpp:132-140): returns xtau[dv] in xc, stores
// xtau[1..dv-1] to the scratch plane xt.
template <class M>
__device__ __forceinline__ void lane_rollout(const double* in, double* xt, const double* x0, const double dtau,
                                             const double* pconst, const double* pfull, double* xc) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  double u[nu], p[Y::np1];
#pragma unroll
  for (int j = 0; j < nx; j++) xc[j] = x0[j];
#pragma unroll 2
  for (int i = 0; i < dv; i++) {
    double f[nx];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * nu + j];
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull ? pfull[i * np + j] : pconst[j];
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
}

// Full sweep with dHdu inside (used for the three Krylov-independent evaluations, where 3 lanes per instance
// are busy): out[i] = dHdu(x_i, u_i, p_i, lambda_{i+1})   (cgmres.hpp:113-162)
template <class M>
__device__ __forceinline__ void lane_sweep_full(const double* in, double* out, double* xt, const double* x0,
                                                const double dtau, const double* pconst, const double* pfull) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  double xc[nx], lmd[nx], u[nu], p[Y::np1];
  lane_rollout<M>(in, xt, x0, dtau, pconst, pfull, xc);
#pragma unroll
  for (int j = 0; j < np; j++) p[j] = pfull ? pfull[dv * np + j] : pconst[j];
  M::dPhidx(lmd, xc, p);  // cgmres.hpp:145
#pragma unroll 2
  for (int i = dv - 1; i >= 0; i--) {  // cgmres.hpp:146-161
    double xi[nx], hu[nu], hx[nx];
#pragma unroll
    for (int j = 0; j < nx; j++) xi[j] = (i > 0) ? xt[(i - 1) * nx + j] : x0[j];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * nu + j];
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull ? pfull[i * np + j] : pconst[j];
    M::dHdu(hu, xi, u, p, lmd);
#pragma unroll
    for (int j = 0; j < nu; j++) out[i * nu + j] = hu[j];
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

// Arnoldi sweeps: the serial lane only runs the two recursions (rollout and costate) and leaves the costates
// lt[i] = ltau[i+1], i = 0..dv-1; the stage-parallel dHdu (cgmres.hpp:156-161) is evaluated afterwards by the
// owning warp, one stage per lane.  This takes ~40 % of the instructions off the serial critical path.
template <class M>
__device__ __forceinline__ void lane_sweep_costates(const double* in, double* xt, double* lt, const double* x0,
                                                    const double dtau, const double* pconst, const double* pfull) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  double xc[nx], lmd[nx], u[nu], p[Y::np1];
  lane_rollout<M>(in, xt, x0, dtau, pconst, pfull, xc);
#pragma unroll
  for (int j = 0; j < np; j++) p[j] = pfull ? pfull[dv * np + j] : pconst[j];
  M::dPhidx(lmd, xc, p);
#pragma unroll
  for (int j = 0; j < nx; j++) lt[(dv - 1) * nx + j] = lmd[j];
#pragma unroll 2
  for (int i = dv - 1; i > 0; i--) {
    double xi[nx], hx[nx];
#pragma unroll
    for (int j = 0; j < nx; j++) xi[j] = xt[(i - 1) * nx + j];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * nu + j];
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull ? pfull[i * np + j] : pconst[j];
    M::dHdx(hx, xi, u, p, lmd);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = hx[j] * dtau;
      lmd[j] = m + lmd[j];
      lt[(i - 1) * nx + j] = lmd[j];
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
__global__ void __launch_bounds__(Lay<M>::threads, Lay<M>::ctas_per_sm) control_kernel(const FastArgs a) {
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

  // ---- phase 0: state in.  U -> F1 area, X = U + h*dUdt (input of the third trajectory), x, p(t) ----------
  if (has) {
    const double* Ug = a.U + n * (int64_t)L;
    const double* dUg = a.dUdt + n * (int64_t)L;
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      if (j < L) {
        const double uu = Ug[j];
        blk[Y::oF1 + j] = uu;  // the F1 area holds U until F1 exists
        double v = dUg[j] * hh;  // cgmres.hpp:168-169
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
    for (int j = 0; j < nu; j++) u0[j] = blk[Y::oF1 + j];
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
    lane_sweep_full<M>(tr == 2 ? b + Y::oX : b + Y::oF1, b + Y::oV + (1 + tr) * L, plane,
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
  // R (packed triangle) and the reflectors are warp-uniform and touched a few times per iteration only: they
  // live in this instance's shared-memory scalars (written by lane 0, read by all lanes after __syncwarp).

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
          blk[Y::oX + j] = v + a.U[n * (int64_t)L + j];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < n_here) {  // transposed sweep: lane = instance; w = A v_k lands in X (gmres.hpp:48)
      double* b = inst_blk(threadIdx.x);
      const double* s = b + Y::oS;
      if (s[Y::sFLAG] == 0.0) {
        const double* pf = PFULL ? a.ptau + (n0 + threadIdx.x) * (int64_t)((M::dv + 1) * np) : nullptr;
        lane_sweep_costates<M>(b + Y::oX, b + Y::oXT, b + Y::oLT, s + Y::sXH, a.dtau_th, s + Y::sP, pf);
      }
    }
    __syncthreads();
    if (solving) {
      // stage-parallel dHdu (cgmres.hpp:156-161) and (F - F1)*inv_h (cgmres.hpp:173-174), one stage per lane;
      // w_i overwrites u_i in X (same lane reads before it writes)
      const double* pf = PFULL ? a.ptau + n * (int64_t)((M::dv + 1) * np) : nullptr;
      for (int i = lane; i < M::dv; i += 32) {
        double xi[nx], u[nu], p[Y::np1], lm[nx], hu[nu];
#pragma unroll
        for (int j = 0; j < nx; j++) {
          xi[j] = (i > 0) ? blk[Y::oXT + (i - 1) * nx + j] : sc[Y::sXH + j];
          lm[j] = blk[Y::oLT + i * nx + j];
        }
#pragma unroll
        for (int j = 0; j < nu; j++) u[j] = blk[Y::oX + i * nu + j];
#pragma unroll
        for (int j = 0; j < np; j++) p[j] = PFULL ? pf[i * np + j] : sc[Y::sP + j];
        M::dHdu(hu, xi, u, p, lm);
#pragma unroll
        for (int j = 0; j < nu; j++) {
          double ax = hu[j] - blk[Y::oF1 + i * nu + j];
          blk[Y::oX + i * nu + j] = ax * inv_h;
        }
      }
      __syncwarp();
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
        const double g0 = sc[Y::sG + 3 * i], g1 = sc[Y::sG + 3 * i + 1], g2 = sc[Y::sG + 3 * i + 2];
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
        if (lane == 0) {
          sc[Y::sG + 3 * k] = g0;
          sc[Y::sG + 3 * k + 1] = g1;
          sc[Y::sG + 3 * k + 2] = g2;
        }
        hc[k] = buf;
        const double rb = g0 * rho[k] * g2;
        rho[k] = rho[k] - rb * g0;
        rho[k + 1] = -rb * g1;
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i <= k; i++) sc[Y::r(i, k)] = hc[i];
      }
      __syncwarp();
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
          if (j < ncol) ri -= sc[Y::r(i, j)] * rho[j];
        ri /= sc[Y::r(i, i)];
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
        double d = dUg[j];  // second (L2-resident) read instead of 2*Q live registers through the whole solve
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
        const double un = Ug[j] + inc;
        Ug[j] = un;
        if (j < nu) blk[Y::oX + j] = un;  // park u = U[0:dim_u] for the epilogue (X is free now)
      }
    }
    __syncwarp();
    if (lane == 0) {
      double x[nx], u0[nu];
#pragma unroll
      for (int j = 0; j < nu; j++) {
        u0[j] = blk[Y::oX + j];
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
