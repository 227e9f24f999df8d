// pipe2_update.cuh -- third-generation on-chip control-update kernel: PERSISTENT, WARP-SPECIALISED, and cut so that
// EVERY serial piece of an instance's update runs lane-per-instance on the group's serial warp while the vector
// warps serve the other group(s):
//
//   * 8 IDENTICAL serial sweeps per update (rollout + costate recursion only, cgmres.hpp:132-153): the three
//     Krylov-independent evaluations F(U,x+dx*h,t+h), F(U,x,t), F(U+h*dUdt,x+dx*h,t+h) run one after the other
//     like the five Arnoldi sweeps, their dHdu (cgmres.hpp:156-161) is stage-parallel vector work.  (The second
//     generation fused the first two as a dual-lane full sweep with dHdu inline: 19.9 k cycles against 2 x 7 k
//     here, and it needed a scratch vector in L2 for F(U,x,t), which is gone.)
//   * EXACT builds (-fmad=false): the 21 sequential sums of an update (15 dots + 6 norms, matrix.hpp:140-159,
//     gmres.hpp:37,52-60) are SERIAL-WARP COMMANDS too.  The owning vector warp parks the element products in the
//     instance's X buffer and hands the group to the serial warp, whose lane l adds instance l's L products in index
//     order; meanwhile the vector warps work on the other group.  In the second generation lane 0 of the owning
//     vector warp did that sum with 31 lanes idle and nothing overlapped (2.7e7 updates/s).
//   * the working vector w of an instance stays in the owning vector warp's registers across those hand-offs (one
//     slice per group); the basis vector of the running Gram-Schmidt step is parked in the (then dead) rollout /
//     costate planes instead of being re-read from L2.
//   * multi-step launches: n_steps closed-loop steps of the SAME resident instances inside one launch (per-step
//     horizon ramps from a device table, optional trajectory log), so small batches pay one launch, not n_steps.
//   * the next round's U is staged global -> shared with cp.async (LDGSTS) while the vector warp finishes the
//     current round's scalars; nothing is staged through registers.
//
// Roles, barriers and placement are the second generation's (pipe_update.cuh): serial warp of group g = warp 4g
// (scheduler 0), vector warps = the warp ids that are not multiples of 4 (+ warp 4*NG as the 16th), hand-off through
// named barriers BX(g) "work for the serial warp of group g" / BL(g) "serial result of group g ready".
//
// Arithmetic: EXACT = false: FMA contraction + butterfly sums (tolerance parity); EXACT = true (instantiated only in
// the -fmad=false translation unit): the reference's operation order and sequential sums, bit-identical results.
#pragma once
#include <type_traits>

#include "pipe_update.cuh"

namespace cgmres_b200 {
namespace pipe2 {

using pipe::bar_sync;
// Producer side of a named-barrier hand-off.  Everything the serial and the vector warps hand to each other lives in
// SHARED memory, and bar.arrive / bar.sync on the same barrier order the producer's prior shared-memory accesses
// before the consumer's later ones (the producer/consumer idiom of the PTX ISA's bar.arrive example and of CUTLASS's
// named barriers): no membar here.  (Generation 2 keeps its fenced pipe::bar_arrive: its first pass hands a vector
// over through GLOBAL memory.)  The CTA-scope fence that used to sit here waited for every outstanding global store
// of 512 vector threads at each of the 29 hand-offs of an update: without it the bit-exact build runs 3.5 % faster,
// the FMA build 1.5 %, results unchanged bit for bit.
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
using pipe::l2_evict_last_policy;
using pipe::ld_keep;
using pipe::st_keep;

// groups in flight per SM, at most.  Measured (msd 65,536 / semiactive 131,072 instances, updates/s): a third msd group
// (compact plane layout below) is SLOWER, 5.4e7 vs 6.6e7; semiactive gains from a third group, bit-exact build 7.1e7 ->
// 8.1e7, FMA build 1.30e8 -> 1.35e8; arm loses (1.3e7 -> 1.1e7).  So: three groups for models with short vector slices
// and a two-state sweep, two otherwise.
#ifndef CG_PIPE2_NGMAX
#define CG_PIPE2_NGMAX(Q, NX) (((Q) <= 5 && (NX) <= 2) ? 3 : 2)
#endif

#ifndef CG_PIPE2_COMPACT
#define CG_PIPE2_COMPACT 0  // measured: a third msd group is SLOWER (5.4e7 vs 6.6e7 updates/s), see profiles/README.md
#endif

template <class M, bool EXACT>
struct Lay {
  using F = fast::Lay<M>;
  static constexpr int nx = F::nx, nu = F::nu, np = F::np, dv = F::dv, km = F::km, L = F::L, np1 = F::np1;
  static constexpr int Q = F::Q;
  // COMPACT plane: models whose dHdu ignores x (Model::dHdu_reads_x == false) keep ONE plane of dv+1 unpadded rows per
  // instance: the rollout writes x_i into row i, the costate recursion overwrites row i with lambda_i once x_i has
  // been consumed, the stage-parallel dHdu reads lambda_{i+1} from row i+1.  That is 204 instead of 495 doubles for
  // mass_spring_damper and lets a third group of instances fit in shared memory.  (The bit-exact build parks a basis
  // vector of L doubles in the planes during Gram-Schmidt and therefore keeps the two-plane layout.)
  static constexpr bool compact = CG_PIPE2_COMPACT && !M::dHdu_reads_x && !EXACT;
  static constexpr int SXT = compact ? nx : F::SXT;
  static constexpr int XT = compact ? (dv + 1) * nx : F::XT;
  static constexpr int LTN = compact ? 0 : F::LTN;
  static_assert(F::SU == nu, "unpadded dim_u rows");
  // instances per group = vector warps per CTA: 16 (18 warps, 96 registers), but 12 for a model whose sweep stage is
  // register-hungry (the arm model's sin/cos): 14 warps = at most 4 per scheduler = 128 registers, no spills
#ifndef CG_PIPE2_GI
#define CG_PIPE2_GI(NX, NU) (((NX) > 2 && (NU) < 4) ? 12 : 16)
#endif
  static constexpr int GI = CG_PIPE2_GI(nx, nu);
  // per-instance shared-memory block (doubles)
  static constexpr int oX = 0;          // sweep input U (+ h*v) -> F in place; EXACT: element products of a sum
  static constexpr int oXT = oX + L;    // rollout states xtau[1..dv-1] (padded rows)
  static constexpr int oLT = oXT + XT;  // costates ltau[1..dv]
  static constexpr int oS = oLT + LTN;  // scalars
  static_assert(!EXACT || XT + LTN >= L, "the Gram-Schmidt basis vector is parked in the rollout/costate planes");
  static constexpr int sR = 0;                        // packed upper triangle R(i,j), i<=j<km
  static constexpr int sG = sR + km * (km + 1) / 2;   // 3*km reflectors
  static constexpr int sX = sG + 3 * km;              // x
  static constexpr int sXH = sX + nx;                 // x + dxdt*h
  static constexpr int sP = sXH + nx;                 // p(t) (repeat mode) / first stage
  static constexpr int sDT = sP + np1;                // dtau(t), dtau(t+h)
  static constexpr int sRHO = sDT + 2;                // rho_e_vec (gmres.hpp:13), km + 1 entries
  static constexpr int sHC = sRHO + km + 1;           // EXACT: Hessenberg column being assembled, km + 2 entries
  static constexpr int sRED = sHC + km + 2;           // EXACT: result of the serial warp's last sum
  static constexpr int sFLAG = sRED + 1;              // 0: solving, else finished
  static constexpr int sCODE = sFLAG + 1;             // exit code | columns << 8 (as a double)
  static constexpr int sXO = sCODE + 1;               // x of the step whose final update is still pending
  static constexpr int sCount = sXO + nx;
  static constexpr int raw = oS + sCount;
  static constexpr int stride = (raw % 2 == 0) ? raw + 1 : raw;  // odd: lane-per-instance accesses hit distinct banks
  static constexpr int NG_fit = (fast::kSmemBudget - 64) / (GI * stride * 8);
  static constexpr int NG = NG_fit < 2 ? 2 : (NG_fit > CG_PIPE2_NGMAX(Q, nx) ? CG_PIPE2_NGMAX(Q, nx) : NG_fit);
  static constexpr int NI = GI * NG;
  static constexpr int NVEC = km + 1;  // stored vectors per instance: id 0 = F1, id 1+i = v_i (v_0's slot first holds b)
  static constexpr int tcols_vec = 2 * Q;
  // warp roles.  CG_PIPE2_SPREAD = 1: warps 0..NG-1 are the serial warps (warp g on scheduler / TMEM lane quarter g),
  // warps NG..NG+GI-1 the vector warps, spread evenly over the four schedulers: the serial warps do NOT share an FP64
  // pipe with each other (a sweep alone keeps ~50 % of one sub-partition's FP64 issue slots busy; two serial warps on
  // one scheduler slowed each other 1.5x, tools/micro/fp64_interference.cu).  CG_PIPE2_SPREAD = 0: the second
  // generation's placement (serial warps 0, 4, .. all on scheduler 0; vector warps on schedulers 1..3, the 16th on 0).
#ifndef CG_PIPE2_SPREAD
#define CG_PIPE2_SPREAD 1
#endif
  static constexpr bool spread = CG_PIPE2_SPREAD != 0;
  static constexpr int GV3 = GI < 15 ? GI : 15;
  static constexpr int last_vec_wid = spread ? NG + GI - 1 : 1 + (GV3 - 1) + (GV3 - 1) / 3;
  static constexpr int last_s0_wid = spread ? NG - 1 : 4 * (NG - 1 + GI - GV3);
  static constexpr int NW = (last_vec_wid > last_s0_wid ? last_vec_wid : last_s0_wid) + 1;
  // column groups of the TMEM allocation: vector warps that share a lane quarter (warp id % 4) need distinct columns
  static constexpr int wq = spread ? (GI + 3) / 4 : (last_vec_wid >> 2) + 1;
  static_assert(spread || last_s0_wid <= last_vec_wid, "warps 4, 8, .. reuse existing column groups");
  static constexpr int NVT_fit = 512 / (wq * NG * tcols_vec);
  static constexpr int NVT = NVT_fit < NVEC ? NVT_fit : NVEC;  // vectors of an instance kept in TMEM
  static_assert(NVT >= 2, "F1 and b / v_0 must fit in tensor memory");
  static constexpr int tcols_slot = NVT * tcols_vec;
  static constexpr int tcols_warp = NG * tcols_slot;
  static constexpr int tcols_need = wq * tcols_warp;
  static constexpr int tcols_alloc = tcols_need <= 32 ? 32 : tcols_need <= 64 ? 64 : tcols_need <= 128 ? 128
                                   : tcols_need <= 256 ? 256 : 512;
  static constexpr int NSCR = NVEC - NVT;  // basis vectors per instance in the per-CTA (L2-resident) global scratch
  static constexpr int threads = 32 * NW;
  static constexpr int bar_threads = 32 * (GI + 1);
  static constexpr size_t smem_bytes = (size_t)NI * stride * 8 + 64;
  static_assert(smem_bytes <= (size_t)fast::kSmemBudget, "groups do not fit in shared memory");
  static constexpr size_t scratch_doubles_per_cta = (size_t)NI * (NSCR > 0 ? NSCR : 1) * L;
  static __host__ __device__ constexpr int r(int i, int j) { return sR + j * (j + 1) / 2 + i; }
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// sum_{j=0..L-1} p[j] in index order starting from 0 (matrix.hpp:140-159), one lane per instance: the loads of the
// next batch are in flight while the current batch is added (the adds are one dependent chain by definition)
#ifndef CG_SUM_DEPTH
#define CG_SUM_DEPTH 3  // batches of 10 shared-memory loads in flight ahead of the dependent adds
#endif
template <int L>
__device__ __forceinline__ double lane_seq_sum(const double* __restrict__ p) {
  constexpr int BS = 10, D = CG_SUM_DEPTH;
  constexpr int NB = L / BS;
  double acc = 0.0;
  double buf[D][BS];
#pragma unroll
  for (int d = 0; d < D; d++) {
    if (d < NB) {
#pragma unroll
      for (int q = 0; q < BS; q++) buf[d][q] = p[d * BS + q];
    }
  }
#pragma unroll
  for (int b = 0; b < NB; b++) {  // fully unrolled: every buffer index is static, no register moves
#pragma unroll
    for (int q = 0; q < BS; q++) acc += buf[b % D][q];
    if (b + D < NB) {
#pragma unroll
      for (int q = 0; q < BS; q++) buf[b % D][q] = p[(b + D) * BS + q];
    }
  }
#pragma unroll
  for (int j = NB * BS; j < L; j++) acc += p[j];
  return acc;
}

// COMPACT-plane sweep (see Lay::compact): rollout + costate recursion of one instance by one lane, one plane P of
// dv+1 rows of dim_x doubles.  On return row i+1 holds lambda_{i+1} = ltau[i+1] for i = 0..dv-1 (cgmres.hpp:132-153).
template <class M, bool PFULL>
__device__ __forceinline__ void lane_sweep_compact(const double* __restrict__ in, double* __restrict__ P,
                                                   const double* __restrict__ x0, const double dtau,
                                                   const double* __restrict__ pconst,
                                                   const double* __restrict__ pfull) {
  constexpr int nx = M::dim_x, nu = M::dim_u, np = M::dim_p, dv = M::dv;
  constexpr int np1 = np > 0 ? np : 1;
  double xc[nx], lmd[nx], u[nu], p[np1];
#pragma unroll
  for (int j = 0; j < np; j++) p[j] = pconst[j];
  // rows 1..dv-1 <- xtau[1..dv-1]; xtau[dv] stays in xc
  fast::lane_rollout<M, PFULL, nx>(in, P + nx, x0, dtau, pconst, pfull, xc);
  if (PFULL) {
#pragma unroll
    for (int j = 0; j < np; j++) p[j] = pfull[dv * np + j];
  }
  M::dPhidx(lmd, xc, p);  // cgmres.hpp:145
#pragma unroll
  for (int j = 0; j < nx; j++) P[dv * nx + j] = lmd[j];
CG_UNROLL((fast::SweepUnroll<M>::value))
  for (int i = dv - 1; i > 0; i--) {
    double xi[nx], hx[nx];
#pragma unroll
    for (int j = 0; j < nx; j++) xi[j] = P[i * nx + j];
#pragma unroll
    for (int j = 0; j < nu; j++) u[j] = in[i * nu + j];
    if (PFULL) {
#pragma unroll
      for (int j = 0; j < np; j++) p[j] = pfull[i * np + j];
    }
    M::dHdx(hx, xi, u, p, lmd);
#pragma unroll
    for (int j = 0; j < nx; j++) {
      double m = hx[j] * dtau;
      lmd[j] = m + lmd[j];
      P[i * nx + j] = lmd[j];  // lambda_i over x_i
    }
  }
}

template <class M, class Sim, bool PFULL, bool EXACT>
__global__ void __launch_bounds__(Lay<M, EXACT>::threads, 1) control_kernel(const FastArgs a) {
  using Y = Lay<M, EXACT>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, L = Y::L, km = Y::km, Q = Y::Q, GI = Y::GI, NG = Y::NG, NI = Y::NI;
  constexpr int T = Y::bar_threads;
  constexpr double hh = M::h;
  constexpr double inv_h = 1.0 / M::h;
  constexpr double c1 = (1 - M::zeta * M::h);
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sg = Y::spread ? (wid < NG ? wid : -1) : (((wid & 3) == 0 && (wid >> 2) < NG) ? (wid >> 2) : -1);
  const int vw_ = Y::spread ? wid - NG
                            : ((wid & 3) ? wid - 1 - (wid >> 2) : ((wid >> 2) >= NG ? Y::GV3 + (wid >> 2) - NG : -1));
  const int vw = (vw_ >= 0 && vw_ < GI) ? vw_ : -1;
  const int64_t nrounds = (a.n + NI - 1) / NI;
  const int64_t prow = (int64_t)(PFULL ? (M::dv + 1) * np : np);
  const int n_steps = a.n_steps > 0 ? a.n_steps : 1;

  // ---- tensor memory: one allocation per CTA for the whole (persistent) kernel -----------------------------------
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)NI * Y::stride);
  if (wid == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tmem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst),
                 "r"((uint32_t)Y::tcols_alloc)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  auto BX = [](int g) { return 1 + g; };
  auto BL = [](int g) { return 1 + NG + g; };
#ifdef CG_PIPE_TIMING
  long long t_wait = 0, t_sw = 0, t_sum = 0, t_dh = 0, t_fin = 0, t_si = 0, t_lap = 0;
  long long t_p1 = 0, t_p2 = 0, t_p3 = 0, t_it = 0, t_mg = 0, t_fi = 0;
  (void)t_lap;
  const long long t_begin = clock64();
#endif

  if (sg >= 0) {
    // =============================== serial warp of group g: lane = instance =======================================
    const int g = sg;
    for (int64_t r = blockIdx.x; r < nrounds; r += gridDim.x) {
      const int64_t n0 = r * NI + (int64_t)g * GI;
      const int n_here = (int)((a.n - n0) < (int64_t)GI ? ((a.n - n0) > 0 ? (a.n - n0) : 0) : (int64_t)GI);
      double* b = sm + (size_t)(g * GI + (lane < GI ? lane : 0)) * Y::stride;
      const double* s = b + Y::oS;
      const double* pf = PFULL ? a.ptau + (n0 + lane) * prow : nullptr;
      // one serial command: wait for the vector warps, run it for the live lanes, hand the group back
      auto sweep = [&](bool at_x, bool check_flag) {  // at_x: F(.., x, t); else F(.., x + dx*h, t + h)
        { CG_PIPE_WAIT_BEGIN; bar_sync(BX(g), T); CG_PIPE_WAIT_END(t_wait); }
        if (lane < n_here && (!check_flag || s[Y::sFLAG] == 0.0)) {
          CG_PIPE_WORK_BEGIN;
          if (Y::compact)
            lane_sweep_compact<M, PFULL>(b + Y::oX, b + Y::oXT, at_x ? s + Y::sX : s + Y::sXH,
                                         at_x ? s[Y::sDT] : s[Y::sDT + 1], s + Y::sP, pf);
          else
            fast::lane_sweep_costates<M, PFULL>(b + Y::oX, b + Y::oXT, b + Y::oLT, at_x ? s + Y::sX : s + Y::sXH,
                                                at_x ? s[Y::sDT] : s[Y::sDT + 1], s + Y::sP, pf);
          CG_PIPE_WORK_END(t_sw);
        }
        bar_arrive(BL(g), T);
      };
      auto seq_sum = [&]() {
        { CG_PIPE_WAIT_BEGIN; bar_sync(BX(g), T); CG_PIPE_WAIT_END(t_wait); }
        if (lane < n_here && s[Y::sFLAG] == 0.0) {
          CG_PIPE_WORK_BEGIN;
          b[Y::oS + Y::sRED] = lane_seq_sum<L>(b + Y::oX);
          CG_PIPE_WORK_END(t_sum);
        }
        bar_arrive(BL(g), T);
      };
      for (int step = 0; step < n_steps; step++) {
        sweep(false, false);  // F(U, x+dx*h, t+h)          cgmres.hpp:88
        sweep(true, false);   // F(U, x, t)                 cgmres.hpp:91
        sweep(false, false);  // F(U + h*dUdt, x+dx*h, t+h) gmres.hpp:33 -> cgmres.hpp:164-175
        if (EXACT) seq_sum();  // ||r0||^2                  gmres.hpp:37
        for (int k = 0; k < km; k++) {
          sweep(false, true);  // F(U + h*v_k, x+dx*h, t+h) gmres.hpp:48
          if (EXACT)
            for (int i = 0; i <= k + 1; i++) seq_sum();  // k+1 dots (gmres.hpp:53) and the norm (gmres.hpp:59)
        }
      }
    }
  } else if (vw >= 0) {
    // =============================== vector warps: warp = instance (of each group) ================================
    const uint64_t keep = l2_evict_last_policy();
    const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * (wid & 3)) << 16) +
                           (uint32_t)((Y::spread ? (vw >> 2) : (wid >> 2)) * Y::tcols_warp);
    double W[EXACT ? NG : 1][Q];  // EXACT: working vector slice of each group across the sum hand-offs

    auto blk_of = [&](int g) { return sm + (size_t)(g * GI + vw) * Y::stride; };
    auto tslot_of = [&](int g) { return tbase + (uint32_t)(g * Y::tcols_slot); };
    // (recomputed from the kernel parameter on every use: a live 64-bit register less through the whole kernel)
    auto scr_of = [&](int g) {
      return a.scratch + (size_t)blockIdx.x * Y::scratch_doubles_per_cta +
             (size_t)(g * GI + vw) * (Y::NSCR > 0 ? Y::NSCR : 1) * L;
    };

    auto vec_store = [&](int id, int g, const double* v) {
      if (id < Y::NVT) {
        fast::basis_store<Q>(tslot_of(g) + (uint32_t)(id * Y::tcols_vec), v);
      } else {
        double* scr = scr_of(g);
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          if (j < L) st_keep(scr + (size_t)(id - Y::NVT) * L + j, v[q], keep);
        }
      }
    };
    auto vec_load = [&](int id, int g, double* v) {
      if (id < Y::NVT) {
        fast::basis_load<Q>(tslot_of(g) + (uint32_t)(id * Y::tcols_vec), v);
      } else {
        const double* scr = scr_of(g);
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          v[q] = (j < L) ? ld_keep(scr + (size_t)(id - Y::NVT) * L + j, keep) : 0.0;
        }
      }
    };
    // X = U + h*v (cgmres.hpp:168-169); U comes from L2 (all loads before the shared-memory stores)
    auto form_x = [&](double* blk, const double* __restrict__ Ug, const double* v) {
      double uu[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        uu[q] = (j < L) ? Ug[j] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) {
          const double t = v[q] * hh;
          blk[Y::oX + j] = t + uu[q];
        }
      }
    };
    // element-distributed read / write of a length-L shared-memory vector
    auto sm_get = [&](const double* base, double* v) {
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        v[q] = (j < L) ? base[j] : 0.0;
      }
    };
    auto sm_put = [&](double* base, const double* v) {
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) base[j] = v[q];
      }
    };
    // stage-parallel dHdu (cgmres.hpp:156-161), one stage per lane; F_i overwrites u_i in X
    auto stage_dhdu = [&](double* blk, const double* sc, const double* x0, int64_t n) {
      CG_PIPE_WORK_BEGIN;
      const double* pf = PFULL ? a.ptau + n * prow : nullptr;
      const bool x16 = ((uint32_t)__cvta_generic_to_shared(blk + Y::oX) & 15u) == 0;  // warp-uniform
      for (int i = lane; i < M::dv; i += 32) {
        double xi[nx], u[nu], p[Y::np1], lm[nx], hu[nu];
#pragma unroll
        for (int j = 0; j < nx; j++) {
          if (Y::compact) {  // dHdu ignores x; lambda_{i+1} sits in row i+1 of the single plane
            xi[j] = 0.0;
            lm[j] = blk[Y::oXT + (i + 1) * nx + j];
          } else {
            xi[j] = (i > 0) ? blk[Y::oXT + (i - 1) * Y::SXT + j] : x0[j];
            lm[j] = blk[Y::oLT + i * Y::SXT + j];
          }
        }
        // u_i / F_i rows: with an even dim_u the row stride (dim_u doubles) puts lanes i and i + 8 on the same banks, a
        // 4-way conflict for 8-byte accesses but none for 16-byte ones (8 consecutive rows cover 8 distinct 16-byte
        // segments of a 128-byte line).  The block stride is odd, so an instance block is 16-byte aligned or 8 bytes off
        // (warp-uniform): pairs (0,1),(2,3).. or a scalar head, pairs (1,2),(3,4).., a scalar tail.
        double* row = blk + Y::oX + i * nu;
        if (nu % 2 == 0 && x16) {
#pragma unroll
          for (int j = 0; j < nu; j += 2) {
            const double2 t = *reinterpret_cast<const double2*>(row + j);
            u[j] = t.x;
            u[j + 1] = t.y;
          }
        } else if (nu % 2 == 0) {
          u[0] = row[0];
#pragma unroll
          for (int j = 1; j + 1 < nu; j += 2) {
            const double2 t = *reinterpret_cast<const double2*>(row + j);
            u[j] = t.x;
            u[j + 1] = t.y;
          }
          u[nu - 1] = row[nu - 1];
        } else {
#pragma unroll
          for (int j = 0; j < nu; j++) u[j] = row[j];
        }
#pragma unroll
        for (int j = 0; j < np; j++) p[j] = PFULL ? pf[i * np + j] : sc[Y::sP + j];
        M::dHdu(hu, xi, u, p, lm);
        if (nu % 2 == 0 && x16) {
#pragma unroll
          for (int j = 0; j < nu; j += 2) *reinterpret_cast<double2*>(row + j) = make_double2(hu[j], hu[j + 1]);
        } else if (nu % 2 == 0) {
          row[0] = hu[0];
#pragma unroll
          for (int j = 1; j + 1 < nu; j += 2) *reinterpret_cast<double2*>(row + j) = make_double2(hu[j], hu[j + 1]);
          row[nu - 1] = hu[nu - 1];
        } else {
#pragma unroll
          for (int j = 0; j < nu; j++) row[j] = hu[j];
        }
      }
      __syncwarp();
      CG_PIPE_WORK_END(t_dh);
    };

    // ---- final update of (round r, group g, global step index gs): back substitution (gmres.hpp:100-107),
    //      dUdt += V y (110-111), U += dUdt*dt (cgmres.hpp:102-103), u, plant step, status, trajectory log ---------
    auto final_update = [&](int64_t r, int g, int step) {
      const int64_t n = r * NI + (int64_t)g * GI + vw;
      if (n >= a.n) return;
      double* blk = blk_of(g);
      const double* sc = blk + Y::oS;
      const int cw = (int)sc[Y::sCODE];
      const int code = cw & 0xFF, ncol = cw >> 8;
      const bool apply = (code == EXIT_FULL || code == EXIT_CONVERGED);
      double rho[km];
#pragma unroll
      for (int i = 0; i < km; i++) rho[i] = sc[Y::sRHO + i];
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
            vec_load(1 + c, g, cv);
#pragma unroll
            for (int q = 0; q < Q; q++) s[q] += cv[q] * rho[c];
          }
        }
      }
      double dd[Q], uu[Q];
      const bool last = step + 1 >= n_steps;  // last touch of U / dUdt in this launch: evict-first accesses
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        dd[q] = (j < L) ? (last ? __ldcs(dUg + j) : dUg[j]) : 0.0;
        uu[q] = (j < L) ? (last ? __ldcs(Ug + j) : Ug[j]) : 0.0;
      }
      double un0 = 0.0;  // element `lane` of the new U: lanes 0..dim_u-1 hold u = U[0:dim_u]
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) {
          double d = dd[q];
          if (apply) {
            d = d + s[q];
            if (last) __stcs(dUg + j, d); else dUg[j] = d;
          }
          const double inc = d * M::dt;
          const double un = uu[q] + inc;
          if (last) __stcs(Ug + j, un); else Ug[j] = un;
          if (q == 0) un0 = un;
        }
      }
      double u0[nu];
#pragma unroll
      for (int j = 0; j < nu; j++) u0[j] = __shfl_sync(0xffffffffu, un0, j);
      if (lane == 0) {
        double x[nx];
#pragma unroll
        for (int j = 0; j < nu; j++) a.u_out[n * nu + j] = u0[j];  // cgmres.hpp:109
        if (a.u_log) {
#pragma unroll
          for (int j = 0; j < nu; j++) a.u_log[((int64_t)step * a.n + n) * nu + j] = u0[j];
        }
        if (a.plant) {  // <example>/main.cpp:74-76 (Euler) or the RK4 option, include/cgmres_b200/plant.hpp
#pragma unroll
          for (int j = 0; j < nx; j++) x[j] = sc[Y::sXO + j];
          plant_step<Sim>(a.plant, x, u0);
#pragma unroll
          for (int j = 0; j < nx; j++) a.x[n * nx + j] = x[j];
          if (a.x_log) {
#pragma unroll
            for (int j = 0; j < nx; j++) a.x_log[((int64_t)step * a.n + n) * nx + j] = x[j];
          }
        }
        a.status[n] = code | (ncol << 8);
      }
      __syncwarp();
    };

    // ---- state in (round r, group g, step): X <- U (cp.async), x, p(t), dtau, x + dxdt*h (cgmres.hpp:83-85) ------
    auto state_in = [&](int64_t r, int g, int step) {
      const int64_t n = r * NI + (int64_t)g * GI + vw;
      if (n >= a.n) return;
      double* blk = blk_of(g);
      double* sc = blk + Y::oS;
      const double* __restrict__ Ug = a.U + n * (int64_t)L;
      if (step == 0) {
#pragma unroll
        for (int q = 0; q < Q; q++) {  // global -> shared without a register round trip; waited for below
          const int j = lane + 32 * q;
          if (j < L) cp_async8(blk + Y::oX + j, Ug + j);
        }
      } else {  // same instances, next step: this lane wrote these elements a moment ago (plain loads see them)
        double uu[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          uu[q] = (j < L) ? Ug[j] : 0.0;
        }
        sm_put(blk + Y::oX, uu);
      }
      {  // dUdt is first needed two sweeps from now: pull its lines into L2 meanwhile (no registers held)
        const char* dline = reinterpret_cast<const char*>(a.dUdt + n * (int64_t)L) + 128 * lane;
        if (128 * lane < L * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(dline));
      }
      double xv = 0.0, pv = 0.0, uv = 0.0;
      if (lane < nx) xv = a.x[n * nx + lane];
      if (lane < np) pv = a.ptau[n * prow + lane];
      if (lane < nu) uv = Ug[lane];
      if (lane < nx) sc[Y::sX + lane] = xv;
      if (lane < np) sc[Y::sP + lane] = pv;
      double dt0, dt1;
      if (a.t_inst) {  // controllers started at different times: per-instance clock, horizon ramp on the device
        double ti = 0.0;
        if (lane == 0) {
          ti = a.t_inst[n];
          a.t_inst[n] = ti + M::dt;  // cgmres.hpp:107
        }
        ti = __shfl_sync(0xffffffffu, ti, 0);
        dt0 = horizon_dtau<M>(ti);
        dt1 = horizon_dtau<M>(ti + hh);
      } else if (a.dtau_tab) {
        dt0 = a.dtau_tab[2 * step];
        dt1 = a.dtau_tab[2 * step + 1];
      } else {
        dt0 = a.dtau_t;
        dt1 = a.dtau_th;
      }
      double x[nx], u0[nu], p0[Y::np1], f[nx];
#pragma unroll
      for (int j = 0; j < nx; j++) x[j] = __shfl_sync(0xffffffffu, xv, j);
#pragma unroll
      for (int j = 0; j < nu; j++) u0[j] = __shfl_sync(0xffffffffu, uv, j);
#pragma unroll
      for (int j = 0; j < Y::np1; j++) p0[j] = __shfl_sync(0xffffffffu, pv, j);
      if (lane == 0) {
        sc[Y::sFLAG] = 0.0;
        sc[Y::sDT] = dt0;
        sc[Y::sDT + 1] = dt1;
        M::dxdt(f, x, u0, p0);
#pragma unroll
        for (int j = 0; j < nx; j++) {
          double m = f[j] * hh;
          sc[Y::sXH + j] = m + x[j];
        }
      }
      cp_async_wait_all();
      __syncwarp();
    };

    // ---- after ||r0||^2: rho0, exit test, v_0, X <- U + h*v_0 (gmres.hpp:37-44) ------------------------------------
    auto after_norm0 = [&](double* blk, int g, int64_t n, double ssq, double* w) {
      double* sc = blk + Y::oS;
      const double rho0 = sqrt(ssq);
      int code = EXIT_FULL;
      bool solving = true;
      if (rho0 < M::tol) {  // gmres.hpp:39-41
        code = EXIT_RHO0;
        solving = false;
      } else {
        const double inv = fast::reciprocal(rho0);  // gmres.hpp:44
#pragma unroll
        for (int q = 0; q < Q; q++) w[q] = w[q] * inv;
        vec_store(1, g, w);  // v_0 (overwrites b)
        form_x(blk, a.U + n * (int64_t)L, w);
      }
      if (lane == 0) {
        sc[Y::sRHO] = rho0;
#pragma unroll
        for (int i = 1; i <= km; i++) sc[Y::sRHO + i] = 0.0;
        sc[Y::sFLAG] = solving ? 0.0 : 1.0;
        sc[Y::sCODE] = (double)code;
      }
      __syncwarp();
    };

    // ---- end of Arnoldi iteration k: breakdown test, v_{k+1}, reflectors, residual, exit test, next sweep input ----
    auto finish_iter = [&](auto kc, double* blk, int g, int64_t n, double hn, double* w, double* hc) {
      constexpr int k = decltype(kc)::value;
      double* sc = blk + Y::oS;
      int code = EXIT_FULL, ncol = k;
      bool solving = true;
      double un[Q];
      if (!EXACT && k + 1 < km) {  // U for the next sweep's input: requested now, consumed after the scalar chain below
        const double* __restrict__ Ug = a.U + n * (int64_t)L;
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          un[q] = (j < L) ? Ug[j] : 0.0;
        }
      }
      if (fabs(hn) < DBL_EPSILON) {  // gmres.hpp:63-65
        code = EXIT_BREAKDOWN;
        solving = false;
      } else {
        hc[k + 1] = hn;
        if (k + 1 < km) {  // the last vector only contributes its Hessenberg column
          const double inv = fast::reciprocal(hn);  // gmres.hpp:67
#pragma unroll
          for (int q = 0; q < Q; q++) w[q] = w[q] * inv;
          vec_store(2 + k, g, w);  // v_{k+1}
        }
        // stored reflectors on the new column (gmres.hpp:71-77), new reflector (78-85), residual (88-90)
#pragma unroll
        for (int i = 0; i < k; i++) {
          const double g0 = sc[Y::sG + 3 * i], g1 = sc[Y::sG + 3 * i + 1], g2 = sc[Y::sG + 3 * i + 2];
          const double buf = (g0 * hc[i] + g1 * hc[i + 1]) * g2;
          hc[i] = hc[i] - buf * g0;
          hc[i + 1] = hc[i + 1] - buf * g1;
        }
        const double ha = hc[k], hb = hc[k + 1];
        const double sgn = (ha < 0.0) ? -1.0 : 1.0;
        const double buf = -sgn * sqrt((0.0 + ha * ha) + hb * hb);
        const double g0 = ha - buf, g1 = hb;
        const double g2 = 2.0 * fast::reciprocal((0.0 + g0 * g0) + g1 * g1);
        hc[k] = buf;
        const double rk = sc[Y::sRHO + k];
        const double rb = g0 * rk * g2;
        const double rk_new = rk - rb * g0;
        const double rk1 = -rb * g1;
        __syncwarp();  // every lane has read rho[k] and the old reflectors
        if (lane == 0) {
          sc[Y::sG + 3 * k] = g0;
          sc[Y::sG + 3 * k + 1] = g1;
          sc[Y::sG + 3 * k + 2] = g2;
          sc[Y::sRHO + k] = rk_new;
          sc[Y::sRHO + k + 1] = rk1;
#pragma unroll
          for (int i = 0; i <= k; i++) sc[Y::r(i, k)] = hc[i];
        }
        ncol = k + 1;
        if (fabs(rk1) < M::tol) {  // gmres.hpp:93-95: break with k not incremented
          code = EXIT_CONVERGED;
          ncol = k;
          solving = false;
        }
      }
      if (lane == 0) {
        sc[Y::sFLAG] = solving ? 0.0 : 1.0;
        sc[Y::sCODE] = (double)(code | (ncol << 8));
      }
      if (solving && k + 1 < km) {  // input of the next sweep (cgmres.hpp:168-169)
        if (EXACT) {  // (the bit-exact build keeps a second working slice live: no registers for the early request)
          form_x(blk, a.U + n * (int64_t)L, w);
        } else {
#pragma unroll
          for (int q = 0; q < Q; q++) {
            const int j = lane + 32 * q;
            if (j < L) {
              const double t = w[q] * hh;
              blk[Y::oX + j] = t + un[q];
            }
          }
        }
      }
      __syncwarp();
    };

    // ---- end of a step of (round r, group g): what the group does next ---------------------------------------------
    // (when another round follows, the final update of (r, g) is deferred into that round's first phase)
    auto end_of_step = [&](int64_t r, int g, int step, bool more_rounds) {
      const int64_t n = r * NI + (int64_t)g * GI + vw;
      double* blk = blk_of(g);
      double* sc = blk + Y::oS;
      if (n < a.n && lane < nx) sc[Y::sXO + lane] = sc[Y::sX + lane];
      __syncwarp();
      if (step + 1 < n_steps) {  // same instances, next closed-loop step: the new U / x go out and come straight back
        { CG_PIPE_WORK_BEGIN; final_update(r, g, step); CG_PIPE_WORK_END(t_fin); }
        { CG_PIPE_WORK_BEGIN; state_in(r, g, step + 1); CG_PIPE_WORK_END(t_si); }
        bar_arrive(BX(g), T);
        return;
      }
      if (more_rounds) {  // next round's state first, so that the serial warp never waits for a round to drain
        { CG_PIPE_WORK_BEGIN; state_in(r + gridDim.x, g, 0); CG_PIPE_WORK_END(t_si); }
        bar_arrive(BX(g), T);
        return;
      }
      { CG_PIPE_WORK_BEGIN; final_update(r, g, step); CG_PIPE_WORK_END(t_fin); }
    };

    if ((int64_t)blockIdx.x < nrounds) {
#pragma unroll
      for (int g = 0; g < NG; g++) {
        state_in(blockIdx.x, g, 0);
        bar_arrive(BX(g), T);
      }
    }

    for (int64_t r = blockIdx.x; r < nrounds; r += gridDim.x) {
      const bool more = r + gridDim.x < nrounds;
      for (int step = 0; step < n_steps; step++) {
        // ---- after sweep 1: F1 = F(U, x+dx*h, t+h) -> TMEM; X <- U for F(U, x, t) ----------------------------------
#pragma unroll
        for (int g = 0; g < NG; g++) {
          const int64_t n = r * NI + (int64_t)g * GI + vw;
          const bool has = n < a.n;
          double* blk = blk_of(g);
          // the previous round of this CTA deferred its final updates to here, where they overlap the serial warp's sweep
          if (r != (int64_t)blockIdx.x && step == 0) {
            CG_PIPE_WORK_BEGIN;
            final_update(r - gridDim.x, g, n_steps - 1);
            CG_PIPE_WORK_END(t_fin);
          }
          { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
          if (has) {
            CG_PIPE_WORK_BEGIN;
            const double* __restrict__ Ug = a.U + n * (int64_t)L;
            double uu[Q], f1[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) {
              const int j = lane + 32 * q;
              uu[q] = (j < L) ? Ug[j] : 0.0;
            }
            stage_dhdu(blk, blk + Y::oS, blk + Y::oS + Y::sXH, n);
            sm_get(blk + Y::oX, f1);
            vec_store(0, g, f1);
            sm_put(blk + Y::oX, uu);
            __syncwarp();
            CG_PIPE_WORK_END(t_p1);
          }
          bar_arrive(BX(g), T);
        }
        // ---- after sweep 2: b = (F(U,x,t)*(1 - zeta*h) - F1)/h (cgmres.hpp:94-96) -> v_0's slot; X <- U + h*dUdt ----
#pragma unroll
        for (int g = 0; g < NG; g++) {
          const int64_t n = r * NI + (int64_t)g * GI + vw;
          const bool has = n < a.n;
          double* blk = blk_of(g);
          { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
          if (has) {
            CG_PIPE_WORK_BEGIN;
            stage_dhdu(blk, blk + Y::oS, blk + Y::oS + Y::sX, n);
            double dd[Q];  // dUdt: an L2 hit (state_in prefetched the lines), in flight during the TMEM round trip
            {
              const double* __restrict__ dUg = a.dUdt + n * (int64_t)L;
#pragma unroll
              for (int q = 0; q < Q; q++) {
                const int j = lane + 32 * q;
                dd[q] = (j < L) ? dUg[j] : 0.0;
              }
            }
            double fa[Q], fb[Q], bb[Q];
            vec_load(0, g, fa);
            sm_get(blk + Y::oX, fb);
#pragma unroll
            for (int q = 0; q < Q; q++) {
              double t = fb[q] * c1;
              t = t - fa[q];
              bb[q] = t * inv_h;
            }
            vec_store(1, g, bb);
            form_x(blk, a.U + n * (int64_t)L, dd);
            __syncwarp();
            CG_PIPE_WORK_END(t_p2);
          }
          bar_arrive(BX(g), T);
        }
        // ---- after sweep 3: r0 = b - A*dUdt (gmres.hpp:33-34); ||r0|| ------------------------------------------------
#pragma unroll
        for (int g = 0; g < NG; g++) {
          const int64_t n = r * NI + (int64_t)g * GI + vw;
          const bool has = n < a.n;
          double* blk = blk_of(g);
          { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
          double w[Q];
#pragma unroll
          for (int q = 0; q < Q; q++) w[q] = 0.0;
          if (has) {
            CG_PIPE_WORK_BEGIN;
            stage_dhdu(blk, blk + Y::oS, blk + Y::oS + Y::sXH, n);
            double fa[Q], bb[Q], fc[Q];
            vec_load(0, g, fa);
            vec_load(1, g, bb);
            sm_get(blk + Y::oX, fc);
#pragma unroll
            for (int q = 0; q < Q; q++) {
              double ax = fc[q] - fa[q];  // cgmres.hpp:173-174
              ax = ax * inv_h;
              w[q] = bb[q] - ax;  // gmres.hpp:34   (slots beyond L: 0 - 0)
            }
            if (EXACT) {
#pragma unroll
              for (int q = 0; q < Q; q++) {
                const int j = lane + 32 * q;
                if (j < L) blk[Y::oX + j] = w[q] * w[q];
              }
              __syncwarp();
            } else {
              double ssq = 0.0;
#pragma unroll
              for (int q = 0; q < Q; q++) ssq += w[q] * w[q];
              after_norm0(blk, g, n, fast::warp_sum(ssq), w);
            }
            CG_PIPE_WORK_END(t_p3);
          }
          if (EXACT) {  // (unconditional: the previous step's slice is dead from here, the registers are free before)
#pragma unroll
            for (int q = 0; q < Q; q++) W[EXACT ? g : 0][q] = w[q];
          }
          bar_arrive(BX(g), T);
        }
        if (EXACT) {
#pragma unroll
          for (int g = 0; g < NG; g++) {
            const int64_t n = r * NI + (int64_t)g * GI + vw;
            const bool has = n < a.n;
            double* blk = blk_of(g);
            { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
            if (has) after_norm0(blk, g, n, blk[Y::oS + Y::sRED], W[EXACT ? g : 0]);
            bar_arrive(BX(g), T);
          }
        }

        // ---- Arnoldi iterations ---------------------------------------------------------------------------------
        auto iteration = [&](auto kc) {
          constexpr int k = decltype(kc)::value;
          // after the sweep: w = A v_k = (F - F1)/h (cgmres.hpp:173-174, gmres.hpp:48), then Gram-Schmidt
#pragma unroll
          for (int g = 0; g < NG; g++) {
            const int64_t n = r * NI + (int64_t)g * GI + vw;
            const bool has = n < a.n;
            double* blk = blk_of(g);
            double* sc = blk + Y::oS;
            { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
            const bool live = has && sc[Y::sFLAG] == 0.0;
            if (live) {
              CG_PIPE_WORK_BEGIN;
              stage_dhdu(blk, sc, sc + Y::sXH, n);
              double w[Q], f1[Q], fx[Q];
              vec_load(0, g, f1);
              sm_get(blk + Y::oX, fx);
#pragma unroll
              for (int q = 0; q < Q; q++) {
                const double ax = fx[q] - f1[q];
                w[q] = ax * inv_h;
              }
              if (EXACT) {  // first dot of the modified Gram-Schmidt sweep (gmres.hpp:52-58): products v_0 . w
                double c[Q];
                vec_load(1, g, c);
#pragma unroll
                for (int q = 0; q < Q; q++) {
                  const int j = lane + 32 * q;
                  W[EXACT ? g : 0][q] = w[q];
                  if (j < L) {
                    blk[Y::oXT + j] = c[q];  // parked in the dead rollout / costate planes for the axpy
                    blk[Y::oX + j] = c[q] * w[q];
                  }
                }
                __syncwarp();
              } else {
                double hc[km + 2];
#pragma unroll
                for (int i = 0; i < km + 2; i++) hc[i] = 0.0;
#pragma unroll
                for (int i = 0; i <= k; i++) {
                  double c[Q];
                  vec_load(1 + i, g, c);
                  double part = 0.0;
#pragma unroll
                  for (int q = 0; q < Q; q++) part += c[q] * w[q];  // slots beyond L hold zeros in both
                  const double hik = fast::warp_sum(part);
                  hc[i] = hik;
#pragma unroll
                  for (int q = 0; q < Q; q++) {
                    const double t = c[q] * hik;
                    w[q] = w[q] - t;
                  }
                }
                double part = 0.0;
#pragma unroll
                for (int q = 0; q < Q; q++) part += w[q] * w[q];
                const double hn = sqrt(fast::warp_sum(part));  // gmres.hpp:59-60
                { CG_PIPE_WORK_BEGIN; finish_iter(kc, blk, g, n, hn, w, hc); CG_PIPE_WORK_END(t_fi); }
              }
              CG_PIPE_WORK_END(t_it);
            }
            if (!EXACT) {
              if (k + 1 < km)
                bar_arrive(BX(g), T);
              else
                end_of_step(r, g, step, more);
            } else {
              bar_arrive(BX(g), T);
            }
          }
          if (EXACT) {
            // after dot i-1 (i = 1..k+1): h_{i-1,k}, w -= h*v_{i-1}; then the next dot's products, or ||w||^2's
#pragma unroll
            for (int i = 1; i <= k + 1; i++) {
#pragma unroll
              for (int g = 0; g < NG; g++) {
                const int64_t n = r * NI + (int64_t)g * GI + vw;
                const bool has = n < a.n;
                double* blk = blk_of(g);
                double* sc = blk + Y::oS;
                double cn[Q];
                if (i <= k) vec_load(1 + i, g, cn);  // next basis vector: requested before the wait
                { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
                const bool live = has && sc[Y::sFLAG] == 0.0;
                if (live) {
                  CG_PIPE_WORK_BEGIN;
                  const double hik = sc[Y::sRED];
                  double* w = W[EXACT ? g : 0];
                  __syncwarp();
                  if (lane == 0) sc[Y::sHC + i - 1] = hik;
#pragma unroll
                  for (int q = 0; q < Q; q++) {
                    const int j = lane + 32 * q;
                    if (j < L) {
                      const double t = blk[Y::oXT + j] * hik;  // the parked v_{i-1}
                      w[q] = w[q] - t;
                      if (i <= k) {
                        blk[Y::oXT + j] = cn[q];
                        blk[Y::oX + j] = cn[q] * w[q];
                      } else {
                        blk[Y::oX + j] = w[q] * w[q];
                      }
                    }
                  }
                  __syncwarp();
                  CG_PIPE_WORK_END(t_mg);
                }
                bar_arrive(BX(g), T);
              }
            }
            // after the norm's sum: h_{k+1,k} = sqrt(.), the rest of the iteration
#pragma unroll
            for (int g = 0; g < NG; g++) {
              const int64_t n = r * NI + (int64_t)g * GI + vw;
              const bool has = n < a.n;
              double* blk = blk_of(g);
              double* sc = blk + Y::oS;
              { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
              const bool live = has && sc[Y::sFLAG] == 0.0;
              if (live) {
                double hc[km + 2];
#pragma unroll
                for (int i = 0; i < km + 2; i++) hc[i] = (i <= k) ? sc[Y::sHC + i] : 0.0;
                const double hn = sqrt(sc[Y::sRED]);  // gmres.hpp:59-60
                __syncwarp();
                { CG_PIPE_WORK_BEGIN; finish_iter(kc, blk, g, n, hn, W[EXACT ? g : 0], hc); CG_PIPE_WORK_END(t_fi); }
              }
              if (k + 1 < km)
                bar_arrive(BX(g), T);
              else
                end_of_step(r, g, step, more);
            }
          }
        };
        iteration(std::integral_constant<int, 0>{});
        if (km > 1) iteration(std::integral_constant<int, (km > 1 ? 1 : 0)>{});
        if (km > 2) iteration(std::integral_constant<int, (km > 2 ? 2 : 0)>{});
        if (km > 3) iteration(std::integral_constant<int, (km > 3 ? 3 : 0)>{});
        if (km > 4) iteration(std::integral_constant<int, (km > 4 ? 4 : 0)>{});
        static_assert(km <= 5, "unrolled for k_max <= 5 (every shipped model uses 5)");
      }
    }
  }

#ifdef CG_PIPE_TIMING
  if (a.dbg && blockIdx.x == 0 && lane == 0 && wid < 24) {  // [2*wid] = cycles blocked, [2*wid+1] = total
    a.dbg[2 * wid] = t_wait;
    a.dbg[2 * wid + 1] = clock64() - t_begin;
    if (wid == 0) {
      a.dbg[48] = t_sw;
      a.dbg[49] = t_sum;
    }
    if (wid == NG) {  // the first vector warp
      a.dbg[51] = t_dh;
      a.dbg[52] = t_fin;
      a.dbg[53] = t_si;
      a.dbg[54] = t_p1;
      a.dbg[55] = t_p2;
      a.dbg[56] = t_p3;
      a.dbg[57] = t_it;
      a.dbg[58] = t_fi;
      a.dbg[59] = t_mg;
    }
  }
#endif
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

}  // namespace pipe2
}  // namespace cgmres_b200
