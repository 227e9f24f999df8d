// fast_update.cuh -- "fast" build mode: the whole control update of a group of G instances runs out of the
// on-chip memories of ONE CTA; HBM only sees the algorithmic minimum (read U, dUdt, x, p; write U, dUdt, x, u:
// ~9 KB per mass_spring_damper update instead of the ~250 KB the streaming exact kernel moves).
//
// Mapping (north star items 2-4):
//   * warp g of the CTA owns instance g for all vector work: lane l holds elements l, l+32, ... of the length-L
//     vectors (the working vector w stays in registers through a whole Gram-Schmidt sweep), dot products and
//     norms are butterfly shuffle reductions, the Hessenberg / Householder / back-substitution scalars
//     (gmres.hpp:71-107) are warp-uniform registers + a few shared-memory words.
//   * the Krylov basis (gmres.hpp:11; 12 KB per mass_spring_damper instance, the largest object by far) lives in
//     TENSOR MEMORY: each lane's slice of a basis vector is a few 32-bit TMEM columns of that lane, written once
//     with tcgen05.st and read back with tcgen05.ld.  TMEM is used purely as 256 KB of extra per-lane private
//     storage (no MMA anywhere); that frees shared memory for 16 instead of 10 resident instances per SM, and
//     resident instances per SM is what bounds this latency-chain-dominated kernel.
//   * the horizon sweeps (cgmres.hpp:113-162) are serial over dv and only dim_x wide, and lanes of a warp
//     cannot run different formulas without divergence, so they are TRANSPOSED: in the sweep phases lane l of
//     the first warp(s) runs the rollout + costate recursion of instance l (three trajectories per instance for
//     the fused F(U,x+dx*h,t+h) / F(U,x,t) / F(U+h*dUdt,x+dx*h,t+h) evaluation) straight out of the instances'
//     shared-memory blocks; the per-instance block stride is odd in 8-byte words so those lane-per-instance
//     accesses are bank-conflict free.  The stage-parallel part of F (dHdu, cgmres.hpp:156-161) is taken off
//     that serial path and evaluated by the owning warp, one stage per lane.
//
// Arithmetic: FMA contraction on (this header is compiled without -fmad=false), tree-ordered reductions.
// Same algorithm, different rounding: judged to the north-star tolerances (1e-9 per update, 1e-6 closed loop),
// not bit for bit.  EXACT_SUMS=true (verification build of the same kernel inside the -fmad=false translation
// unit) replaces the shuffle reductions by the reference's sequential sums, done lane-per-instance by the first
// warp, and is bit-identical to the reference.
#pragma once
#include <float.h>
#include <stdint.h>

#include "cgmres_b200/models.hpp"
#include "cgmres_b200/plant.hpp"
#include "kernel_args.h"

namespace cgmres_b200 {
namespace fast {

constexpr int kSmemBudget = 227 * 1024 - 1024;  // minus the per-CTA reserved kilobyte
// max instances (= warps) per CTA.  Register file: 16 warps leave 128 registers per thread, 24 warps 85; the
// kernels with long vector slices (Q >= 8 per lane) or a 4-state sweep need ~110-125, semiactive 80 (GPU sweep in
// profiles/README.md: msd 16 > 18/20 > 12; semiactive 28 (spills) > 24 > 20 > 16).
#ifndef CG_SWEEP_UNROLL
#define CG_SWEEP_UNROLL 5  // stages per unrolled body of the serial recursions (exposes the next stages' loads)
#endif
#define CG_PRAGMA(x) _Pragma(#x)
#define CG_UNROLL(n) CG_PRAGMA(unroll n)
// stages per unrolled body of a model's serial recursions: the translation unit's CG_SWEEP_UNROLL, halved for models
// whose stage is register-hungry (the arm model's sin/cos: 10 stages deep spilled in the persistent kernels)
template <class M>
struct SweepUnroll {
  static constexpr int value = (M::dim_x > 2 && M::dim_u < 4 && CG_SWEEP_UNROLL > 5) ? 5 : CG_SWEEP_UNROLL;
};
// phase timestamps of one warp of CTA 0 (debug builds only: -DCG_FAST_TIMING), read back by tools/phase_times.py
#ifdef CG_FAST_TIMING
#define CG_MARK(i)                                                                      \
  do {                                                                                  \
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 32 * 5) a.dbg[(i)] = clock64();      \
  } while (0)
#define CG_MARK_SERIAL(i)                                                        \
  do {                                                                           \
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[(i)] = clock64();    \
  } while (0)
#else
#define CG_MARK(i) \
  do {             \
  } while (0)
#define CG_MARK_SERIAL(i) \
  do {                    \
  } while (0)
#endif
// Padding the dim_u rows as well removes the remaining 2-way conflicts of the stage-parallel phase but costs an
// integer division per element in every element-wise pass: measured slower for msd (4.7e7 vs 5.1e7), so off.
#ifndef CG_FAST_PAD_U
#define CG_FAST_PAD_U 0
#endif
#ifndef CG_FAST_MAXCTAS
#define CG_FAST_MAXCTAS 1
#endif
#ifdef CG_FAST_GFORCE  // tuning override: same cap for every model
#define CG_FAST_GCAP(Q, NX) (CG_FAST_GFORCE)
#endif
#ifndef CG_FAST_GCAP
#define CG_FAST_GCAP(Q, NX) (((Q) >= 8 || (NX) > 2) ? 16 : 24)
#endif

template <class M>
struct Lay {
  static constexpr int nx = M::dim_x, nu = M::dim_u, np = M::dim_p, dv = M::dv, km = M::k_max;
  static constexpr int L = nu * dv;
  static constexpr int np1 = np > 0 ? np : 1;
  // Shared-memory vectors are stored stage-major with an ODD row stride (SU doubles per stage of dim_u values, SXT
  // per stage of dim_x values): in the stage-parallel phase lane i works on stage i, and rows 6 or 4 doubles
  // apart would put 2-4 lanes of a half-warp on the same pair of banks (measured: that phase was bound by the
  // replayed shared-memory wavefronts of the 16 warps).  Element e = i*dim_u + j of a length-L vector sits at pos(e).
  static constexpr int SU = CG_FAST_PAD_U ? ((nu % 2 == 0) ? nu + 1 : nu) : nu;
  static constexpr int SXT = (nx % 2 == 0) ? nx + 1 : nx;
  static constexpr int LV = dv * SU;                      // doubles of one stored length-L vector
  static constexpr int XT = SXT * (dv > 1 ? dv - 1 : 1);  // stored rollout states xtau[1..dv-1]
  static constexpr int LTN = SXT * dv;                    // stored costates ltau[1..dv]
  static constexpr int Q = (L + 31) / 32;                // vector elements per lane
  static __host__ __device__ constexpr int pos(int e) { return SU == nu ? e : (e / nu) * SU + (e % nu); }
  // per-instance shared-memory block, offsets in doubles.  (U itself is not kept on chip: it is re-read from
  // global memory -- an L2 hit, the CTA touched it microseconds earlier -- where U + h*v is formed and in the
  // final update.)
  static constexpr int oF1 = 0;        // first evaluation: U -> F(U,x+dx*h,t+h) in place; then F_dxh_h (cgmres.hpp:202)
  static constexpr int oB = oF1 + LV;  // first evaluation: U -> F(U,x,t) in place; afterwards the costate plane
  static constexpr int oX = oB + LV;   // first evaluation: U+h*dUdt -> F(..) in place; then U+h*v -> w (cgmres.hpp:168-174)
  static constexpr int oXT = oX + LV;  // rollout states of trajectory A / of the Arnoldi sweeps (padded rows)
  static constexpr int oXTB = oXT + XT;   // rollout states of trajectory B (first evaluation only)
  static constexpr int oXTC = oXTB + XT;  // rollout states of trajectory C (first evaluation only)
  static constexpr bool lt_alias = LTN <= LV;  // costates of the Arnoldi sweeps reuse the dead F(U,x,t) area
  static constexpr int oLT = lt_alias ? oB : oXTC + XT;
  static constexpr int oS = oXTC + XT + (lt_alias ? 0 : LTN);  // scalars
  // scalar slots
  static constexpr int sR = 0;                         // packed upper triangle R(i,j), i<=j<km
  static constexpr int sG = sR + km * (km + 1) / 2;    // 3*km reflectors
  static constexpr int sX = sG + 3 * km;               // x
  static constexpr int sXH = sX + nx;                  // x + dxdt*h
  static constexpr int sP = sXH + nx;                  // p(t) (repeat mode) / first stage
  static constexpr int sDT = sP + np1;                 // dtau(t), dtau(t+h) of this instance
  static constexpr int sRED = sDT + 2;                 // reduction result slot (EXACT_SUMS)
  static constexpr int sFLAG = sRED + 1;               // state: 0 solving, else finished (as double)
  static constexpr int sCount = sFLAG + 1;
  static constexpr int raw = oS + sCount;
  // odd number of 8-byte words per instance => lane-per-instance accesses hit distinct banks
  static constexpr int stride = (raw % 2 == 0) ? raw + 1 : raw;
  // tensor memory: 2*Q 32-bit columns per basis vector per instance; warps w, w+4, w+8, ... share a lane quarter
  static constexpr int tcols_vec = 2 * Q;
  static constexpr int tcols_inst = km * tcols_vec;
  static constexpr int G_tmem = 4 * (512 / tcols_inst);
  static constexpr int G_smem = (kSmemBudget - 64) / (stride * 8);
  static constexpr int G_fit = G_smem < G_tmem ? G_smem : G_tmem;
  static constexpr int G = G_fit > CG_FAST_GCAP(Q, nx) ? CG_FAST_GCAP(Q, nx) : G_fit;  // instances per CTA, one CTA per SM
  // CTAs that can be co-resident on one SM (shared memory, tensor-memory columns, 16-warp register budget)
  static constexpr int ctas_per_sm_ = (G_fit / G) < 1 ? 1 : (G_fit / G);
  static constexpr int ctas_per_sm = ctas_per_sm_ > CG_FAST_MAXCTAS ? CG_FAST_MAXCTAS : ctas_per_sm_;
  static constexpr int threads = 32 * G;
  static constexpr size_t smem_bytes = (size_t)G * stride * 8 + 64;  // + TMEM base address word
  // TMEM columns this CTA allocates: a power of two >= 32 covering ceil(G/4) column slots
  static constexpr int tcols_need = ((G + 3) / 4) * tcols_inst;
  static constexpr int tcols_alloc = tcols_need <= 32 ? 32 : tcols_need <= 64 ? 64 : tcols_need <= 128 ? 128
                                   : tcols_need <= 256 ? 256 : 512;
  static_assert(tcols_need <= 512, "Krylov basis does not fit in tensor memory");
  static __host__ __device__ constexpr int r(int i, int j) { return sR + j * (j + 1) / 2 + i; }
};

// ---- tensor memory as per-lane private storage -----------------------------------------------------------------
// tcgen05.{st,ld}.32x32b.xN: thread i of the warp moves N consecutive 32-bit columns of TMEM lane (base + i).
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t* r);
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_st<1>(uint32_t t, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(t), "r"(r[0]) : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<2>(uint32_t t, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t), "r"(r[0]), "r"(r[1]) : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<4>(uint32_t t, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<8>(uint32_t t, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t t, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(t),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<1>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(t) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<2>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(t) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(t)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(t)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t t, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(t)
      : "memory");
}

// move NW 32-bit words (any NW <= 31) as power-of-two chunks
template <int NW, int OFF = 0>
__device__ __forceinline__ void tmem_st_words(uint32_t taddr, const uint32_t* r) {
  if constexpr (NW >= 16) {
    tmem_st<16>(taddr + OFF, r + OFF);
    tmem_st_words<NW - 16, OFF + 16>(taddr, r);
  } else if constexpr (NW >= 8) {
    tmem_st<8>(taddr + OFF, r + OFF);
    tmem_st_words<NW - 8, OFF + 8>(taddr, r);
  } else if constexpr (NW >= 4) {
    tmem_st<4>(taddr + OFF, r + OFF);
    tmem_st_words<NW - 4, OFF + 4>(taddr, r);
  } else if constexpr (NW >= 2) {
    tmem_st<2>(taddr + OFF, r + OFF);
    tmem_st_words<NW - 2, OFF + 2>(taddr, r);
  } else if constexpr (NW == 1) {
    tmem_st<1>(taddr + OFF, r + OFF);
  }
}
template <int NW, int OFF = 0>
__device__ __forceinline__ void tmem_ld_words(uint32_t taddr, uint32_t* r) {
  if constexpr (NW >= 16) {
    tmem_ld<16>(taddr + OFF, r + OFF);
    tmem_ld_words<NW - 16, OFF + 16>(taddr, r);
  } else if constexpr (NW >= 8) {
    tmem_ld<8>(taddr + OFF, r + OFF);
    tmem_ld_words<NW - 8, OFF + 8>(taddr, r);
  } else if constexpr (NW >= 4) {
    tmem_ld<4>(taddr + OFF, r + OFF);
    tmem_ld_words<NW - 4, OFF + 4>(taddr, r);
  } else if constexpr (NW >= 2) {
    tmem_ld<2>(taddr + OFF, r + OFF);
    tmem_ld_words<NW - 2, OFF + 2>(taddr, r);
  } else if constexpr (NW == 1) {
    tmem_ld<1>(taddr + OFF, r + OFF);
  }
}
// store / load this lane's Q-double slice of one basis vector (whole warp must call, converged)
template <int Q>
__device__ __forceinline__ void basis_store(uint32_t taddr, const double* v) {
  uint32_t r[2 * Q];
  __syncwarp();  // .sync.aligned: the whole warp must arrive converged
#pragma unroll
  for (int q = 0; q < Q; q++) {
    r[2 * q] = (uint32_t)__double2loint(v[q]);
    r[2 * q + 1] = (uint32_t)__double2hiint(v[q]);
  }
  tmem_st_words<2 * Q>(taddr, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
template <int Q>
__device__ __forceinline__ void basis_load(uint32_t taddr, double* v) {
  uint32_t r[2 * Q];
  __syncwarp();
  tmem_ld_words<2 * Q>(taddr, r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int q = 0; q < Q; q++) v[q] = __hiloint2double((int)r[2 * q + 1], (int)r[2 * q]);
}

// Forward Euler rollout of one instance by ONE lane (cgmres.hpp:132-140): returns xtau[dv] in xc, stores
// xtau[1..dv-1] to the scratch plane xt.
template <class M, bool PFULL, int SX>
__device__ __forceinline__ void lane_rollout(const double* __restrict__ in, double* __restrict__ xt,
                                             const double* __restrict__ x0, const double dtau,
                                             const double* __restrict__ pconst, const double* __restrict__ pfull,
                                             double* __restrict__ xc) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  double u[nu], p[Y::np1];
#pragma unroll
  for (int j = 0; j < np; j++) p[j] = pconst[j];  // constant reference (set_ptau_repeat); PFULL reloads per stage
#pragma unroll
  for (int j = 0; j < nx; j++) xc[j] = x0[j];
CG_UNROLL((SweepUnroll<M>::value))
  for (int i = 0; i < dv; i++) {
    double f[nx];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * Y::SU + j];
    if (PFULL) {
#pragma unroll
      for (int j = 0; j < np; j++) p[j] = pfull[i * np + j];
    }
    M::dxdt(f, xc, u, p);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = f[j] * dtau;
      xc[j] = m + xc[j];
    }
    if (i + 1 < dv) {
#pragma unroll
      for (int j = 0; j < nx; j++) xt[i * SX + j] = xc[j];
    }
  }
}

// Full sweep with dHdu inside (used for the three Krylov-independent evaluations, where 3 lanes per instance
// are busy): out[i] = dHdu(x_i, u_i, p_i, lambda_{i+1})   (cgmres.hpp:113-162)
// share_mask != 0: several lanes of the warp read the SAME `in` while one of them writes `out == in` (the second
// generation's dual-trajectory pass): the lanes of the mask then synchronise between a stage's loads and its stores,
// so the reader never depends on lock-step execution to see u_i before it is overwritten.
template <class M, bool PFULL, int SX>
__device__ __forceinline__ void lane_sweep_full(const double* in, double* out, double* __restrict__ xt,
                                                const double* __restrict__ x0, const double dtau,
                                                const double* __restrict__ pconst,
                                                const double* __restrict__ pfull, const unsigned share_mask = 0u) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  double xc[nx], lmd[nx], u[nu], p[Y::np1];
#pragma unroll
  for (int j = 0; j < np; j++) p[j] = pconst[j];
  lane_rollout<M, PFULL, SX>(in, xt, x0, dtau, pconst, pfull, xc);
  if (PFULL) {
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull[dv * np + j];
  }
  M::dPhidx(lmd, xc, p);  // cgmres.hpp:145
CG_UNROLL((SweepUnroll<M>::value))
  for (int i = dv - 1; i >= 0; i--) {  // cgmres.hpp:146-161
    double xi[nx], hu[nu], hx[nx];
#pragma unroll
    for (int j = 0; j < nx; j++) xi[j] = (i > 0) ? xt[(i - 1) * SX + j] : x0[j];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * Y::SU + j];
    if (PFULL) {
#pragma unroll
      for (int j = 0; j < np; j++) p[j] = pfull[i * np + j];
    }
    M::dHdu(hu, xi, u, p, lmd);
    if (share_mask) __syncwarp(share_mask);  // every sharing lane holds u_i before stage i is overwritten
#pragma unroll
    for (int j = 0; j < nu; j++) out[i * Y::SU + j] = hu[j];
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
template <class M, bool PFULL>
// (in, xt and lt never overlap: __restrict__ lets the compiler hoist the next stages' loads above this stage's
//  stores, which takes the shared-memory latency off the recursion's critical path)
__device__ __forceinline__ void lane_sweep_costates(const double* __restrict__ in, double* __restrict__ xt,
                                                    double* __restrict__ lt, const double* __restrict__ x0,
                                                    const double dtau, const double* __restrict__ pconst,
                                                    const double* __restrict__ pfull) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, dv = Y::dv;
  constexpr int SX = Y::SXT;
  double xc[nx], lmd[nx], u[nu], p[Y::np1];
#pragma unroll
  for (int j = 0; j < np; j++) p[j] = pconst[j];
  lane_rollout<M, PFULL, SX>(in, xt, x0, dtau, pconst, pfull, xc);
  if (PFULL) {
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull[dv * np + j];
  }
  M::dPhidx(lmd, xc, p);
#pragma unroll
  for (int j = 0; j < nx; j++) lt[(dv - 1) * Y::SXT + j] = lmd[j];
CG_UNROLL((SweepUnroll<M>::value))
  for (int i = dv - 1; i > 0; i--) {
    double xi[nx], hx[nx];
#pragma unroll
    for (int j = 0; j < nx; j++) xi[j] = xt[(i - 1) * SX + j];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * Y::SU + j];
    if (PFULL) {
#pragma unroll
      for (int j = 0; j < np; j++) p[j] = pfull[i * np + j];
    }
    M::dHdx(hx, xi, u, p, lmd);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = hx[j] * dtau;
      lmd[j] = m + lmd[j];
      lt[(i - 1) * Y::SXT + j] = lmd[j];
    }
  }
}

// 1.0/x, correctly rounded (identical to the IEEE division the reference performs), without the generic
// division subroutine
__device__ __forceinline__ double reciprocal(double x) { return __drcp_rn(x); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The kernel.  grid = ceil(n / G) CTAs of 32*G threads, one CTA per SM;
// dynamic shared memory = Lay<M>::smem_bytes.
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
  const bool has = wid < n_here;  // this warp owns a live instance (warp-uniform)
  const int64_t n = n0 + wid;     // its global index
  double* const blk = sm + (size_t)wid * Y::stride;
  double* const sc = blk + Y::oS;
  auto inst_blk = [&](int g) { return sm + (size_t)g * Y::stride; };

  // ---- tensor memory: one allocation, released at the end by the same warp --------------------------------------
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)G * Y::stride);
  CG_MARK_SERIAL(60);
  if (wid == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tmem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst),
                 "r"((uint32_t)Y::tcols_alloc)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  CG_MARK_SERIAL(61);

  CG_MARK(1);
  // ---- phase 0: state in.  F1 and B areas <- U, X <- U + h*dUdt (third trajectory), x, p(t) -----------------
  if (has) {
    const double* __restrict__ Ug = a.U + n * (int64_t)L;
    const double* __restrict__ dUg = a.dUdt + n * (int64_t)L;
    // all global loads first (the compiler cannot prove that the generic-pointer loads do not alias the
    // shared-memory stores, and would otherwise serialise Q HBM round trips), then the stores
    double uu[Q], dd[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      uu[q] = (j < L) ? Ug[j] : 0.0;
      dd[q] = (j < L) ? dUg[j] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      if (j < L) {
        const int e = Y::pos(j);
        blk[Y::oF1 + e] = uu[q];
        blk[Y::oB + e] = uu[q];
        double v = dd[q] * hh;  // cgmres.hpp:168-169
        blk[Y::oX + e] = v + uu[q];
      }
    }
    if (lane < nx) sc[Y::sX + lane] = a.x[n * nx + lane];
    if (lane < np) sc[Y::sP + lane] = a.ptau[n * (int64_t)(PFULL ? (M::dv + 1) * np : np) + lane];
    if (lane == 0) {
      sc[Y::sFLAG] = 0.0;
      if (a.t_inst) {  // controllers started at different times: per-instance clock, horizon ramp on the device
        const double ti = a.t_inst[n];
        sc[Y::sDT] = horizon_dtau<M>(ti);
        sc[Y::sDT + 1] = horizon_dtau<M>(ti + hh);
        a.t_inst[n] = ti + M::dt;  // cgmres.hpp:107
      } else {
        sc[Y::sDT] = a.dtau_t;
        sc[Y::sDT + 1] = a.dtau_th;
      }
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this warp's lane quarter (warp id % 4) and column slot (warp id / 4) of the allocation
  const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)((wid >> 2) * Y::tcols_inst);
  auto tcol = [&](int k) { return tbase + (uint32_t)(k * Y::tcols_vec); };

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

  CG_MARK(2);
  // ---- phase 1: the three Krylov-independent F evaluations, one lane per (instance, trajectory), in place ---
  //   A: F(U, x+dx*h, t+h) in the F1 area   B: F(U, x, t) in the B area   C: F(U+h*dUdt, x+dx*h, t+h) in X
  if (threadIdx.x < 3 * n_here) {
    const int g = threadIdx.x / 3, tr = threadIdx.x % 3;
    double* b = inst_blk(g);
    const double* s = b + Y::oS;
    const double* pf = PFULL ? a.ptau + (n0 + g) * (int64_t)((M::dv + 1) * np) : nullptr;
    double* io = b + (tr == 0 ? Y::oF1 : (tr == 1 ? Y::oB : Y::oX));
    const double* x0p = tr == 1 ? s + Y::sX : s + Y::sXH;
    const double dtau = tr == 1 ? s[Y::sDT] : s[Y::sDT + 1];
    // one code path for all three trajectories (lanes of a warp must not diverge here): same padded plane layout
    double* plane = b + (tr == 0 ? Y::oXT : (tr == 1 ? Y::oXTB : Y::oXTC));
    lane_sweep_full<M, PFULL, Y::SXT>(io, io, plane, x0p, dtau, s + Y::sP, pf);
  }
  __syncthreads();

  CG_MARK(3);
  // ---- phase 2: b, r0 = b - A*dUdt (F1 stays where trajectory A left it); rho0 -------------------------------
  double w[Q];  // working vector slice (r0, then each new Krylov vector)
  double ssq = 0.0;
#pragma unroll
  for (int q = 0; q < Q; q++) w[q] = 0.0;
  if (has) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      if (j < L) {
        const int e = Y::pos(j);
        const double fa = blk[Y::oF1 + e], fb = blk[Y::oB + e], fc = blk[Y::oX + e];
        double b = fb * c1;  // cgmres.hpp:94-96
        b = b - fa;
        b = b * inv_h;
        double ax = fc - fa;  // cgmres.hpp:173-174
        ax = ax * inv_h;
        w[q] = b - ax;  // gmres.hpp:34
        if (EXACT_SUMS)
          blk[Y::oX + e] = w[q] * w[q];
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
      // loads of a batch are issued together; the adds keep the reference's index order
      if (Y::SU == nu) {
        constexpr int BS = 10;
        for (int j0 = 0; j0 + BS <= L; j0 += BS) {
          double v[BS];
#pragma unroll
          for (int q = 0; q < BS; q++) v[q] = b[Y::oX + j0 + q];
#pragma unroll
          for (int q = 0; q < BS; q++) s += v[q];
        }
        for (int j = (L / BS) * BS; j < L; j++) s += b[Y::oX + j];
      } else {
        for (int i = 0; i < M::dv; i++) {
          double v[nu];
#pragma unroll
          for (int q = 0; q < nu; q++) v[q] = b[Y::oX + i * Y::SU + q];
#pragma unroll
          for (int q = 0; q < nu; q++) s += v[q];
        }
      }
      b[Y::oS + Y::sRED] = s;
    }
    __syncthreads();
    return has ? sc[Y::sRED] : 0.0;
  };

  int code = EXIT_FULL, ncol = 0;
  bool solving = has;
  double rho[km + 1];
#pragma unroll
  for (int i = 0; i <= km; i++) rho[i] = 0.0;
  {
    const double rho0 = sqrt(reduce(ssq));  // gmres.hpp:37
    rho[0] = rho0;
    if (solving) {
      if (rho0 < M::tol) {  // gmres.hpp:39-41
        code = EXIT_RHO0;
        solving = false;
      } else {
        const double inv = reciprocal(rho0);  // gmres.hpp:44: div() multiplies by the rounded reciprocal
#pragma unroll
        for (int q = 0; q < Q; q++) w[q] = w[q] * inv;
        basis_store<Q>(tcol(0), w);  // v_0
      }
    }
  }
  if (has && lane == 0) sc[Y::sFLAG] = solving ? 0.0 : 1.0;

  // ---- Arnoldi iterations ------------------------------------------------------------------------------------
  CG_MARK(5);
  // R (packed triangle) and the reflectors are warp-uniform and touched a few times per iteration only: they
  // live in this instance's shared-memory scalars (written by lane 0, read by all lanes after __syncwarp).
#pragma unroll
  for (int k = 0; k < km; k++) {
    CG_MARK(10 + 5 * k + 0);
    // X = U + h*v_k (cgmres.hpp:168-169); w currently holds v_k
    if (solving) {
      double uu[Q];  // U from L2: loads first, then the shared-memory stores (see phase 0)
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        uu[q] = (j < L) ? a.U[n * (int64_t)L + j] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) {
          const double v = w[q] * hh;
          blk[Y::oX + Y::pos(j)] = v + uu[q];
        }
      }
    }
    CG_MARK(10 + 5 * k + 1);
    __syncthreads();
    if (threadIdx.x < n_here) {  // transposed recursions: lane = instance (cgmres.hpp:132-153)
      double* b = inst_blk(threadIdx.x);
      const double* s = b + Y::oS;
      if (s[Y::sFLAG] == 0.0) {
        const double* pf = PFULL ? a.ptau + (n0 + threadIdx.x) * (int64_t)((M::dv + 1) * np) : nullptr;
        CG_MARK_SERIAL(50 + 2 * k);
        lane_sweep_costates<M, PFULL>(b + Y::oX, b + Y::oXT, b + Y::oLT, s + Y::sXH, s[Y::sDT + 1], s + Y::sP, pf);
        CG_MARK_SERIAL(51 + 2 * k);
      }
    }
    __syncthreads();
    CG_MARK(10 + 5 * k + 2);
    if (solving) {
      // stage-parallel dHdu (cgmres.hpp:156-161), one stage per lane; F_i overwrites u_i in X (same lane reads
      // before it writes); then w = A v_k = (F - F1)*inv_h (cgmres.hpp:173-174, gmres.hpp:48)
      const double* pf = PFULL ? a.ptau + n * (int64_t)((M::dv + 1) * np) : nullptr;
      for (int i = lane; i < M::dv; i += 32) {
        double xi[nx], u[nu], p[Y::np1], lm[nx], hu[nu];
#pragma unroll
        for (int j = 0; j < nx; j++) {
          xi[j] = (i > 0) ? blk[Y::oXT + (i - 1) * Y::SXT + j] : sc[Y::sXH + j];
          lm[j] = blk[Y::oLT + i * Y::SXT + j];
        }
#pragma unroll
        for (int j = 0; j < nu; j++) u[j] = blk[Y::oX + i * Y::SU + j];
#pragma unroll
        for (int j = 0; j < np; j++) p[j] = PFULL ? pf[i * np + j] : sc[Y::sP + j];
        M::dHdu(hu, xi, u, p, lm);
#pragma unroll
        for (int j = 0; j < nu; j++) blk[Y::oX + i * Y::SU + j] = hu[j];
      }
      __syncwarp();
      // (F - F1)*inv_h in the element-distributed layout: these shared-memory reads are conflict free, whereas
      // reading F1 stage-wise above would replay every access (row stride of dim_u doubles)
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        w[q] = 0.0;
        if (j < L) {
          const double ax = blk[Y::oX + Y::pos(j)] - blk[Y::oF1 + Y::pos(j)];
          w[q] = ax * inv_h;
        }
      }
    }
    CG_MARK(10 + 5 * k + 3);
    // modified Gram-Schmidt (gmres.hpp:52-58) against v_i, i = 0..k; each v_i slice comes from TMEM once
    double hc[km + 2];
#pragma unroll
    for (int i = 0; i < km + 2; i++) hc[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= k; i++) {
      double c[Q];
      double part = 0.0;
      if (solving) {
        basis_load<Q>(tcol(i), c);
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          if (j < L) {
            if (EXACT_SUMS)
              blk[Y::oX + Y::pos(j)] = c[q] * w[q];
            else
              part += c[q] * w[q];
          }
        }
      }
      const double hik = reduce(part);
      hc[i] = hik;
      if (solving) {
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const double t = c[q] * hik;
          w[q] = w[q] - t;
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
            blk[Y::oX + Y::pos(j)] = w[q] * w[q];
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
      hc[k + 1] = hn;
      if (k + 1 < km) {  // the last vector only contributes its Hessenberg column
        const double inv = reciprocal(hn);  // gmres.hpp:67
#pragma unroll
        for (int q = 0; q < Q; q++) w[q] = w[q] * inv;
        basis_store<Q>(tcol(k + 1), w);  // v_{k+1}
      }
      CG_MARK(10 + 5 * k + 4);
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
        const double g2 = 2.0 * reciprocal((0.0 + g0 * g0) + g1 * g1);  // == 2.0/x: scaling by 2 is exact
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

  CG_MARK(39);
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
    double s[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) s[q] = 0.0;
    if (apply) {  // s = sum_c v_c*y_c in column order (matrix.hpp:82-91)
#pragma unroll
      for (int c = 0; c < km; c++) {
        if (c < ncol) {
          double cv[Q];
          basis_load<Q>(tcol(c), cv);
#pragma unroll
          for (int q = 0; q < Q; q++) s[q] += cv[q] * rho[c];
        }
      }
    }
    // second (L2-resident) read of dUdt and U instead of 4*Q live registers through the whole solve; again all
    // loads before the first store
    double dd[Q], uu[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      dd[q] = (j < L) ? dUg[j] : 0.0;
      uu[q] = (j < L) ? Ug[j] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = lane + 32 * q;
      if (j < L) {
        double d = dd[q];
        if (apply) {
          d = d + s[q];
          dUg[j] = d;
        }
        const double inc = d * M::dt;
        const double un = uu[q] + inc;
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
      if (a.plant) {  // <example>/main.cpp:74-76 (Euler) or the RK4 option, include/cgmres_b200/plant.hpp
#pragma unroll
        for (int j = 0; j < nx; j++) x[j] = sc[Y::sX + j];
        plant_step<Sim>(a.plant, x, u0);
#pragma unroll
        for (int j = 0; j < nx; j++) a.x[n * nx + j] = x[j];
      }
      a.status[n] = code | (ncol << 8);
    }
  }

  CG_MARK(40);
  // ---- release tensor memory ------------------------------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot),
                 "r"((uint32_t)Y::tcols_alloc)
                 : "memory");
  }
}

}  // namespace fast
}  // namespace cgmres_b200
