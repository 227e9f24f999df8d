// pipe_update.cuh -- the "fast" build mode, second generation: a PERSISTENT, WARP-SPECIALISED version of the
// on-chip kernel of fast_update.cuh in which the serial horizon recursions and the vector work of two groups of
// instances overlap.
//
// Why: the on-chip kernel is bound by  resident instances per SM / latency of one instance's dependent chain
// (DESIGN.md section 5).  About half of that chain is the serial rollout + costate recursion (cgmres.hpp:132-153),
// during which the 16 vector warps of the CTA idle at a barrier, and during the vector phases the serial warp
// idles.  Here
//   * ONE CTA per SM holds NG = 2 groups of GI = 16 instances (32 resident instances per SM instead of 16);
//   * VECTOR warp v owns instance v of group 0 AND instance v of group 1 and alternates between them (dHdu per
//     stage, (F - F1)/h, Gram-Schmidt with shuffle reductions, Householder scalars, U + h*v for the next sweep,
//     final update);
//   * each group has its own SERIAL warp: lane l runs the transposed recursion of instance l of that group (two
//     lanes per instance for the fused F(U,x+dx*h,t+h) / F(U,x,t) pair of the first evaluation);
//   * the roles hand a group to each other through named barriers (bar.arrive by the producer, bar.sync by the
//     consumer): while group A is being swept by its serial warp the vector warps work on group B and vice versa,
//     in steady state across rounds as well (the kernel is persistent: CTA b processes rounds b, b + gridDim.x, ..
//     and a round's final update is deferred into the next round's first step, where the vector warps would
//     otherwise wait for the first sweep);
//   * warp placement matters: the serial warps are warps 0 and 4, i.e. they have scheduler 0 (warp id % 4) almost
//     to themselves, because their dependent chains must not queue for issue slots behind the bursty vector
//     warps (measured: 5.3e7 -> 6.0e7 updates/s from the placement alone); the vector warps are the warps whose id
//     is not a multiple of 4 (schedulers and TMEM lane quarters 1..3, five warps each = 96 registers per thread),
//     the 16th is warp 8.
// Doubling the residency needs the per-instance shared-memory block to shrink from 13.4 KB to 6.8 KB (msd):
// F(U,x+dx*h,t+h) ("F1", cgmres.hpp:202) moves into tensor memory next to the first Krylov vector; what does not
// fit in the TMEM quarters the vector warps can reach (v_1..v_4 for msd; nothing for the smaller models) and the
// transient F(U,x,t) go to a per-CTA global scratch that is reused every round and therefore stays in L2
// (148 x 384 KB << 126 MB; the L2 still writes about half of those stores back to HBM).
//
// Arithmetic is the same as fast_update.cuh's FAST instantiation (FMA contraction, butterfly sums): results are
// identical to that kernel, tolerance parity against the reference.
#pragma once
#include "fast_update.cuh"

namespace cgmres_b200 {
namespace pipe {

// groups in flight per SM.  The small models would fit three or four (each with its own serial warp), measured
// slower: semiactive 1.29e8 with 3 vs 1.45e8 with 2, arm 1.4e7 vs 1.95e7 (three or four serial warps share one
// scheduler, and fewer basis vectors fit in tensor memory).
#ifndef CG_PIPE_NGMAX
#define CG_PIPE_NGMAX 2
#endif
#ifndef CG_PIPE_GI
// instances per group = vector warps per CTA: 16, but 12 for a model whose sweep stage is register-hungry (the arm
// model's sin/cos): 16 instead of 20 warps = 128 instead of 96 registers per thread, no spills
#define CG_PIPE_GI(NX, NU) (((NX) > 2 && (NU) < 4) ? 12 : 16)
#endif

template <class M>
struct Lay {
  using F = fast::Lay<M>;
  static constexpr int nx = F::nx, nu = F::nu, np = F::np, dv = F::dv, km = F::km, L = F::L, np1 = F::np1;
  static constexpr int SXT = F::SXT, XT = F::XT, LTN = F::LTN, Q = F::Q;
  static_assert(F::SU == nu, "pipelined kernel assumes unpadded dim_u rows");
  static constexpr int GI = CG_PIPE_GI(nx, nu);
  // groups in flight per SM: as many as fit in shared memory (each brings its own serial warp), at most CG_PIPE_NGMAX
  static constexpr int raw_ = L + XT + LTN + (km * (km + 1) / 2 + 3 * km + 3 * nx + np1 + 2 + km + 1 + 2);
  static constexpr int stride_ = (raw_ % 2 == 0) ? raw_ + 1 : raw_;
  static constexpr int NG_fit = (fast::kSmemBudget - 64) / (GI * stride_ * 8);
  static constexpr int NG = NG_fit < 2 ? 2 : (NG_fit > CG_PIPE_NGMAX ? CG_PIPE_NGMAX : NG_fit);
  static constexpr int NI = GI * NG;
  static constexpr int NVEC = km + 1;  // stored vectors per instance: id 0 = F1, id 1+i = v_i
  // tensor memory: 2*Q columns per vector; warps w, w+4, ... share a lane quarter
  static constexpr int tcols_vec = 2 * Q;
  // warp roles (see the header comment): serial warps 0, 4; vector warp v < 15 is warp 1 + v + v/3; further vector
  // warps are warps 8, 12, ..; other multiples of 4 below NW idle.  (A 16th vector warp on schedulers 1..3 would be
  // the sixth on one of them and cap the kernel at 80 registers.)
  static constexpr int GV3 = GI < 15 ? GI : 15;  // vector warps on schedulers 1..3
  static constexpr int last_vec_wid = 1 + (GV3 - 1) + (GV3 - 1) / 3;
  static constexpr int last_s0_wid = 4 * (NG - 1 + GI - GV3);  // serial warps 0, 4; vector warps GV3.. are 8, ...
  static constexpr int NW = (last_vec_wid > last_s0_wid ? last_vec_wid : last_s0_wid) + 1;
  static constexpr int wq = (last_vec_wid >> 2) + 1;  // column groups of the TMEM allocation (warp id / 4)
  static_assert(last_s0_wid <= last_vec_wid, "warps 4, 8, .. reuse existing column groups");
  static constexpr int NVT_fit = 512 / (wq * NG * tcols_vec);
  static constexpr int NVT = NVT_fit < NVEC ? NVT_fit : NVEC;  // vectors of an instance kept in TMEM
  static_assert(NVT >= 1, "F1 must fit in tensor memory");
  static constexpr int tcols_slot = NVT * tcols_vec;  // one (warp, group) slot
  static constexpr int tcols_warp = NG * tcols_slot;
  static constexpr int tcols_need = wq * tcols_warp;
  static constexpr int tcols_alloc = tcols_need <= 32 ? 32 : tcols_need <= 64 ? 64 : tcols_need <= 128 ? 128
                                   : tcols_need <= 256 ? 256 : 512;
  // global (L2-resident) scratch vectors per instance: the basis vectors that did not fit + F(U,x,t)
  static constexpr int NSCR = (NVEC - NVT) + 1;
  // per-instance shared-memory block (doubles)
  static constexpr int oX = 0;          // U -> F1 (first pass), U+h*dUdt -> F (second pass), U+h*v_k -> F (sweeps)
  static constexpr int oXT = oX + L;    // rollout states xtau[1..dv-1] (padded rows)
  static constexpr int oLT = oXT + XT;  // costates ltau[1..dv]; first pass: rollout states of the F(U,x,t) lane
  static constexpr int oS = oLT + LTN;  // scalars
  static constexpr int sR = 0;
  static constexpr int sG = sR + km * (km + 1) / 2;
  static constexpr int sX = sG + 3 * km;
  static constexpr int sXH = sX + nx;
  static constexpr int sP = sXH + nx;
  static constexpr int sDT = sP + np1;
  static constexpr int sRHO = sDT + 2;      // rho_e_vec (gmres.hpp:13), km + 1 entries
  static constexpr int sFLAG = sRHO + km + 1;  // 0: solving, else finished
  static constexpr int sCODE = sFLAG + 1;      // exit code | columns << 8 (as a double)
  static constexpr int sXO = sCODE + 1;        // x of the round whose final update is still pending
  static constexpr int sCount = sXO + nx;
  static constexpr int raw = oS + sCount;
  static constexpr int stride = (raw % 2 == 0) ? raw + 1 : raw;
  static_assert(stride == stride_, "keep raw_ in step with the scalar slots");
  static constexpr int threads = 32 * NW;
  static constexpr int bar_threads = 32 * (GI + 1);  // participants of the named barriers: vector warps + serial warp
  static constexpr size_t smem_bytes = (size_t)NI * stride * 8 + 64;
  static_assert(smem_bytes <= (size_t)fast::kSmemBudget, "two groups do not fit in shared memory");
  static_assert(2 * GI <= 32, "first pass uses two lanes of the serial warp per instance");
  static constexpr size_t scratch_doubles_per_cta = (size_t)NI * NSCR * L;
  static __host__ __device__ constexpr int r(int i, int j) { return sR + j * (j + 1) / 2 + i; }
};

// wait-time accounting of debug builds (-DCG_PIPE_TIMING): cycles a warp spends blocked in bar_sync
#ifdef CG_PIPE_TIMING
#define CG_PIPE_WAIT_BEGIN const long long t_wait0_ = clock64()
// (BAR.SYNC returns before the warp is actually released; a dependent shared-memory read makes the wait visible)
#define CG_PIPE_WAIT_END(acc)                                   \
  do {                                                          \
    if (*(volatile double*)sm == 1.2345e300) (acc) += 1;        \
    t_lap = clock64();                                          \
    (acc) += t_lap - t_wait0_;                                  \
  } while (0)
#define CG_PIPE_LAP(acc)                  \
  do {                                    \
    const long long t_now_ = clock64();   \
    (acc) += t_now_ - t_lap;              \
    t_lap = t_now_;                       \
  } while (0)
#define CG_PIPE_WORK_BEGIN const long long t_work0_ = clock64()
#define CG_PIPE_WORK_END(acc) (acc) += clock64() - t_work0_
#else
#define CG_PIPE_LAP(acc) \
  do {                   \
  } while (0)
#define CG_PIPE_WORK_BEGIN \
  do {                     \
  } while (0)
#define CG_PIPE_WORK_END(acc) \
  do {                        \
  } while (0)
#define CG_PIPE_WAIT_BEGIN \
  do {                     \
  } while (0)
#define CG_PIPE_WAIT_END(acc) \
  do {                        \
  } while (0)
#endif
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  __threadfence_block();  // the producer's shared/global stores are visible before the consumer is released
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// scratch accesses carry an L2 evict-last policy: the per-CTA scratch is rewritten every round and must not be
// pushed out (and written back to HBM) by the U / dUdt lines streaming through the L2
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep(double* p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld_keep(const double* p, uint64_t pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}

// EXACT = true (instantiated only in the -fmad=false translation unit; the ABI's PIPELINED_EXACT verification mode):
// every dot product and norm is the reference's sequential sum (matrix.hpp:140-159) -- the owning warp parks the
// element products in the instance's (free) X buffer and lane 0 adds them in index order -- and the kernel is
// bit-identical to the reference.  It proves that the fast build differs from a bit-exact kernel in nothing but FMA
// contraction and the order of those sums.
template <class M, class Sim, bool PFULL, bool EXACT = false>
__global__ void __launch_bounds__(Lay<M>::threads, 1) control_kernel(const FastArgs a) {
  using Y = Lay<M>;
  constexpr int nx = Y::nx, nu = Y::nu, np = Y::np, L = Y::L, km = Y::km, Q = Y::Q, GI = Y::GI, NG = Y::NG, NI = Y::NI;
  constexpr int T = Y::bar_threads;
  constexpr double hh = M::h;
  constexpr double inv_h = 1.0 / M::h;
  constexpr double c1 = (1 - M::zeta * M::h);
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // roles: warp 4*g (g < NG) is the serial warp of group g; vector warps are the warps with id % 4 != 0 in order,
  // then warps 4*NG, 4*NG+4, ... (only GI > 15); everything else idles
  const int sg = ((wid & 3) == 0 && (wid >> 2) < NG) ? (wid >> 2) : -1;
  const int vw_ = (wid & 3) ? wid - 1 - (wid >> 2) : ((wid >> 2) >= NG ? Y::GV3 + (wid >> 2) - NG : -1);
  const int vw = (vw_ >= 0 && vw_ < GI) ? vw_ : -1;
  const int64_t nrounds = (a.n + NI - 1) / NI;
  const int64_t prow = (int64_t)(PFULL ? (M::dv + 1) * np : np);

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

  // barrier ids: BX(g) "inputs of group g ready for the serial warp", BL(g) "serial results of group g ready"
  auto BX = [](int g) { return 1 + g; };
  auto BL = [](int g) { return 1 + NG + g; };
  double* const scr_cta = a.scratch + (size_t)blockIdx.x * Y::scratch_doubles_per_cta;
#ifdef CG_PIPE_TIMING
  long long t_wait = 0, t_p1 = 0, t_p2 = 0, t_sw = 0;
  long long t_lap = 0, t_fin = 0, t_v1 = 0, t_v2 = 0, t_dh = 0, t_w = 0, t_mgs = 0, t_hh = 0, t_fx = 0, t_si = 0;
  const long long t_begin = clock64();
#endif

  if (sg >= 0) {
    // =============================== serial warps: one per group, lane = instance ================================
    const int g = sg;
    for (int64_t r = blockIdx.x; r < nrounds; r += gridDim.x) {
      // pass 1: A = F(U, x+dx*h, t+h) in place in X (even lanes), B = F(U, x, t) into the scratch (odd lanes)
      {
        const int64_t n0 = r * NI + (int64_t)g * GI;
        const int n_here = (int)((a.n - n0) < (int64_t)GI ? ((a.n - n0) > 0 ? (a.n - n0) : 0) : (int64_t)GI);
        { CG_PIPE_WAIT_BEGIN; bar_sync(BX(g), T); CG_PIPE_WAIT_END(t_wait); }
        const unsigned pair_mask = __ballot_sync(0xffffffffu, lane < 2 * n_here);  // the lanes that run pass 1
        if (lane < 2 * n_here) {
          const int inst = lane >> 1, tr = lane & 1;
          double* b = sm + (size_t)(g * GI + inst) * Y::stride;
          const double* s = b + Y::oS;
          const double* pf = PFULL ? a.ptau + (n0 + inst) * prow : nullptr;
          double* out = tr == 0 ? b + Y::oX : scr_cta + ((size_t)(g * GI + inst) * Y::NSCR + (Y::NSCR - 1)) * L;
          const double* x0p = tr == 0 ? s + Y::sXH : s + Y::sX;
          const double dtau = tr == 0 ? s[Y::sDT + 1] : s[Y::sDT];
          double* plane = b + (tr == 0 ? Y::oXT : Y::oLT);
          CG_PIPE_WORK_BEGIN;
          // (the even lane overwrites X in place while the odd lane of the pair still reads it: share_mask)
          fast::lane_sweep_full<M, PFULL, Y::SXT>(b + Y::oX, out, plane, x0p, dtau, s + Y::sP, pf, pair_mask);
          CG_PIPE_WORK_END(t_p1);
        }
        bar_arrive(BL(g), T);
      }
      // pass 2: rollout + costates of C = F(U + h*dUdt, x+dx*h, t+h); its dHdu is stage-parallel vector work
      {
        const int64_t n0 = r * NI + (int64_t)g * GI;
        const int n_here = (int)((a.n - n0) < (int64_t)GI ? ((a.n - n0) > 0 ? (a.n - n0) : 0) : (int64_t)GI);
        { CG_PIPE_WAIT_BEGIN; bar_sync(BX(g), T); CG_PIPE_WAIT_END(t_wait); }
        if (lane < n_here) {
          double* b = sm + (size_t)(g * GI + lane) * Y::stride;
          const double* s = b + Y::oS;
          const double* pf = PFULL ? a.ptau + (n0 + lane) * prow : nullptr;
          CG_PIPE_WORK_BEGIN;
          fast::lane_sweep_costates<M, PFULL>(b + Y::oX, b + Y::oXT, b + Y::oLT, s + Y::sXH, s[Y::sDT + 1],
                                              s + Y::sP, pf);
          CG_PIPE_WORK_END(t_p2);
        }
        bar_arrive(BL(g), T);
      }
      // Arnoldi sweeps: rollout + costate recursion only (cgmres.hpp:132-153)
      for (int k = 0; k < km; k++) {
        {
          const int64_t n0 = r * NI + (int64_t)g * GI;
          const int n_here = (int)((a.n - n0) < (int64_t)GI ? ((a.n - n0) > 0 ? (a.n - n0) : 0) : (int64_t)GI);
          { CG_PIPE_WAIT_BEGIN; bar_sync(BX(g), T); CG_PIPE_WAIT_END(t_wait); }
          if (lane < n_here) {
            double* b = sm + (size_t)(g * GI + lane) * Y::stride;
            const double* s = b + Y::oS;
            if (s[Y::sFLAG] == 0.0) {
              const double* pf = PFULL ? a.ptau + (n0 + lane) * prow : nullptr;
              CG_PIPE_WORK_BEGIN;
              fast::lane_sweep_costates<M, PFULL>(b + Y::oX, b + Y::oXT, b + Y::oLT, s + Y::sXH, s[Y::sDT + 1],
                                                  s + Y::sP, pf);
              CG_PIPE_WORK_END(t_sw);
            }
          }
          bar_arrive(BL(g), T);
        }
      }
    }
  } else if (vw >= 0) {
    // =============================== vector warps: warp = instance (of each group) ==============================
    const uint64_t keep = l2_evict_last_policy();
    const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)((wid >> 2) * Y::tcols_warp);

    // stored vectors: id 0 = F1, 1+i = v_i; the first NVT live in this warp's TMEM slot of the group, the rest in
    // the CTA's global scratch (element-distributed, coalesced)
    auto vec_store = [&](int id, uint32_t tslot, double* scr, const double* v) {
      if (id < Y::NVT) {
        fast::basis_store<Q>(tslot + (uint32_t)(id * Y::tcols_vec), v);
      } else {
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int j = lane + 32 * q;
          if (j < L) st_keep(scr + (size_t)(id - Y::NVT) * L + j, v[q], keep);
        }
      }
    };
    auto vec_load = [&](int id, uint32_t tslot, const double* scr, double* v) {
      if (id < Y::NVT) {
        fast::basis_load<Q>(tslot + (uint32_t)(id * Y::tcols_vec), v);
      } else {
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

    // EXACT: sum_{j=0..L-1} prod_j in index order, starting from 0 (the reference's loop); X is free whenever
    // this is called.  Loads of a batch are issued together, the adds keep their order.
    auto seq_sum = [&](double* blk, const double* prod) -> double {
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) blk[Y::oX + j] = prod[q];
      }
      __syncwarp();
      double acc = 0.0;
      if (lane == 0) {
        constexpr int BS = 10;
        for (int j0 = 0; j0 + BS <= L; j0 += BS) {
          double v[BS];
#pragma unroll
          for (int q = 0; q < BS; q++) v[q] = blk[Y::oX + j0 + q];
#pragma unroll
          for (int q = 0; q < BS; q++) acc += v[q];
        }
        for (int j = (L / BS) * BS; j < L; j++) acc += blk[Y::oX + j];
      }
      acc = __shfl_sync(0xffffffffu, acc, 0);
      __syncwarp();
      return acc;
    };

    // stage-parallel dHdu (cgmres.hpp:156-161), one stage per lane; F_i overwrites u_i in X
    auto stage_dhdu = [&](double* blk, const double* sc, int64_t n) {
      const double* pf = PFULL ? a.ptau + n * prow : nullptr;
      for (int i = lane; i < M::dv; i += 32) {
        double xi[nx], u[nu], p[Y::np1], lm[nx], hu[nu];
#pragma unroll
        for (int j = 0; j < nx; j++) {
          xi[j] = (i > 0) ? blk[Y::oXT + (i - 1) * Y::SXT + j] : sc[Y::sXH + j];
          lm[j] = blk[Y::oLT + i * Y::SXT + j];
        }
#pragma unroll
        for (int j = 0; j < nu; j++) u[j] = blk[Y::oX + i * nu + j];
#pragma unroll
        for (int j = 0; j < np; j++) p[j] = PFULL ? pf[i * np + j] : sc[Y::sP + j];
        M::dHdu(hu, xi, u, p, lm);
#pragma unroll
        for (int j = 0; j < nu; j++) blk[Y::oX + i * nu + j] = hu[j];
      }
      __syncwarp();
    };

    // ---- final update of (round r, group g): back substitution (gmres.hpp:100-107), dUdt += V y (110-111),
    //      U += dUdt*dt (cgmres.hpp:102-103), u, plant step, status.  Runs after the group's state_in of the NEXT
    //      round: it only needs the solve scalars, x (kept in sXO) and the stored basis, none of which state_in or
    //      the first pass touch.
    auto final_update = [&](int64_t r, int g) {
      const int64_t n = r * NI + (int64_t)g * GI + vw;
      if (n >= a.n) return;
      double* blk = sm + (size_t)(g * GI + vw) * Y::stride;
      const double* sc = blk + Y::oS;
      const uint32_t tslot = tbase + (uint32_t)(g * Y::tcols_slot);
      const double* scr = scr_cta + (size_t)(g * GI + vw) * Y::NSCR * L;
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
            vec_load(1 + c, tslot, scr, cv);
#pragma unroll
            for (int q = 0; q < Q; q++) s[q] += cv[q] * rho[c];
          }
        }
      }
      double dd[Q], uu[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        // last touch of this instance's U / dUdt in this launch: evict-first loads and stores, so that the L2
        // keeps the CTAs' scratch vectors (re-used every round) instead of these dead lines
        dd[q] = (j < L) ? __ldcs(dUg + j) : 0.0;
        uu[q] = (j < L) ? __ldcs(Ug + j) : 0.0;
      }
      double un0 = 0.0;  // element `lane` of the new U: lanes 0..dim_u-1 hold u = U[0:dim_u]
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) {
          double d = dd[q];
          if (apply) {
            d = d + s[q];
            __stcs(dUg + j, d);
          }
          const double inc = d * M::dt;
          const double un = uu[q] + inc;
          __stcs(Ug + j, un);
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
        if (a.plant) {  // <example>/main.cpp:74-76 (Euler) or the RK4 option, include/cgmres_b200/plant.hpp
#pragma unroll
          for (int j = 0; j < nx; j++) x[j] = sc[Y::sXO + j];
          plant_step<Sim>(a.plant, x, u0);
#pragma unroll
          for (int j = 0; j < nx; j++) a.x[n * nx + j] = x[j];
        }
        a.status[n] = code | (ncol << 8);
      }
      __syncwarp();
    };

    // ---- state in (round r, group g): X <- U, x, p(t), dtau, x + dxdt*h (cgmres.hpp:83-85) --------------------
    auto state_in = [&](int64_t r, int g) {
      const int64_t n = r * NI + (int64_t)g * GI + vw;
      if (n >= a.n) return;
      double* blk = sm + (size_t)(g * GI + vw) * Y::stride;
      double* sc = blk + Y::oS;
      const double* __restrict__ Ug = a.U + n * (int64_t)L;
      double uu[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        uu[q] = (j < L) ? Ug[j] : 0.0;
      }
      double xv = 0.0, pv = 0.0;
      if (lane < nx) xv = a.x[n * nx + lane];
      if (lane < np) pv = a.ptau[n * prow + lane];
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = lane + 32 * q;
        if (j < L) blk[Y::oX + j] = uu[q];
      }
      if (lane < nx) sc[Y::sX + lane] = xv;
      if (lane < np) sc[Y::sP + lane] = pv;
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
      __syncwarp();
      if (lane == 0) {
        double x[nx], u0[nu], p0[Y::np1], f[nx];
#pragma unroll
        for (int j = 0; j < nx; j++) x[j] = sc[Y::sX + j];
#pragma unroll
        for (int j = 0; j < nu; j++) u0[j] = blk[Y::oX + j];
#pragma unroll
        for (int j = 0; j < np; j++) p0[j] = sc[Y::sP + j];
        M::dxdt(f, x, u0, p0);
#pragma unroll
        for (int j = 0; j < nx; j++) {
          double m = f[j] * hh;
          sc[Y::sXH + j] = m + x[j];
        }
      }
    };

    if ((int64_t)blockIdx.x < nrounds) {
      for (int g = 0; g < NG; g++) {
        state_in(blockIdx.x, g);
        bar_arrive(BX(g), T);
      }
    }

    for (int64_t r = blockIdx.x; r < nrounds; r += gridDim.x) {
      const int64_t r_next = r + gridDim.x;
      const bool more = r_next < nrounds;

      // ---- after pass 1: park F1 in TMEM, X <- U + h*dUdt for pass 2 -------------------------------------------
      for (int g = 0; g < NG; g++) {
        const int64_t n = r * NI + (int64_t)g * GI + vw;
        const bool has = n < a.n;
        double* blk = sm + (size_t)(g * GI + vw) * Y::stride;
        const uint32_t tslot = tbase + (uint32_t)(g * Y::tcols_slot);
        double* scr = scr_cta + (size_t)(g * GI + vw) * Y::NSCR * L;
        CG_PIPE_LAP(t_wait);
        if (r != (int64_t)blockIdx.x) final_update(r - gridDim.x, g);  // overlaps the serial warp's first pass
        CG_PIPE_LAP(t_fin);
        { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
        if (has) {
          double f1[Q], dd[Q];
          const double* __restrict__ dUg = a.dUdt + n * (int64_t)L;
#pragma unroll
          for (int q = 0; q < Q; q++) {
            const int j = lane + 32 * q;
            dd[q] = (j < L) ? dUg[j] : 0.0;
            f1[q] = (j < L) ? blk[Y::oX + j] : 0.0;
          }
          vec_store(0, tslot, scr, f1);
          form_x(blk, a.U + n * (int64_t)L, dd);
        }
        CG_PIPE_LAP(t_v1);
        bar_arrive(BX(g), T);
      }

      // ---- after pass 2: b, r0 = b - A*dUdt, rho0, v_0 (cgmres.hpp:94-96, gmres.hpp:33-44); X <- U + h*v_0 ------
      for (int g = 0; g < NG; g++) {
        const int64_t n = r * NI + (int64_t)g * GI + vw;
        const bool has = n < a.n;
        double* blk = sm + (size_t)(g * GI + vw) * Y::stride;
        double* sc = blk + Y::oS;
        const uint32_t tslot = tbase + (uint32_t)(g * Y::tcols_slot);
        double* scr = scr_cta + (size_t)(g * GI + vw) * Y::NSCR * L;
        { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
        if (has) {
          stage_dhdu(blk, sc, n);
          double w[Q], fa[Q];
          const double* __restrict__ fbg = scr + (size_t)(Y::NSCR - 1) * L;
          double fb[Q];
#pragma unroll
          for (int q = 0; q < Q; q++) {
            const int j = lane + 32 * q;
            fb[q] = (j < L) ? fbg[j] : 0.0;
          }
          vec_load(0, tslot, scr, fa);
          double ssq = 0.0;
#pragma unroll
          for (int q = 0; q < Q; q++) {
            const int j = lane + 32 * q;
            w[q] = 0.0;
            if (j < L) {
              const double fc = blk[Y::oX + j];
              double b = fb[q] * c1;  // cgmres.hpp:94-96
              b = b - fa[q];
              b = b * inv_h;
              double ax = fc - fa[q];  // cgmres.hpp:173-174
              ax = ax * inv_h;
              w[q] = b - ax;  // gmres.hpp:34
              if (!EXACT) ssq += w[q] * w[q];
            }
          }
          if (EXACT) {
            double pr[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) pr[q] = w[q] * w[q];
            ssq = seq_sum(blk, pr);
          } else {
            ssq = fast::warp_sum(ssq);
          }
          const double rho0 = sqrt(ssq);  // gmres.hpp:37
          int code = EXIT_FULL;
          bool solving = true;
          if (rho0 < M::tol) {  // gmres.hpp:39-41
            code = EXIT_RHO0;
            solving = false;
          } else {
            const double inv = fast::reciprocal(rho0);  // gmres.hpp:44
#pragma unroll
            for (int q = 0; q < Q; q++) w[q] = w[q] * inv;
            vec_store(1, tslot, scr, w);  // v_0
            form_x(blk, a.U + n * (int64_t)L, w);
          }
          if (lane == 0) {
            sc[Y::sRHO] = rho0;
#pragma unroll
            for (int i = 1; i <= km; i++) sc[Y::sRHO + i] = 0.0;
            sc[Y::sFLAG] = solving ? 0.0 : 1.0;
            sc[Y::sCODE] = (double)code;
          }
        }
        CG_PIPE_LAP(t_v2);
        bar_arrive(BX(g), T);
      }

      // ---- Arnoldi iterations -----------------------------------------------------------------------------------
#pragma unroll
      for (int k = 0; k < km; k++) {
        for (int g = 0; g < NG; g++) {
          const int64_t n = r * NI + (int64_t)g * GI + vw;
          const bool has = n < a.n;
          double* blk = sm + (size_t)(g * GI + vw) * Y::stride;
          double* sc = blk + Y::oS;
          const uint32_t tslot = tbase + (uint32_t)(g * Y::tcols_slot);
          double* scr = scr_cta + (size_t)(g * GI + vw) * Y::NSCR * L;
          { CG_PIPE_WAIT_BEGIN; bar_sync(BL(g), T); CG_PIPE_WAIT_END(t_wait); }
          if (has && sc[Y::sFLAG] == 0.0) {
            stage_dhdu(blk, sc, n);
            CG_PIPE_LAP(t_dh);
            double w[Q];
            {
              double f1[Q];
              vec_load(0, tslot, scr, f1);
#pragma unroll
              for (int q = 0; q < Q; q++) {
                const int j = lane + 32 * q;
                w[q] = 0.0;
                if (j < L) {
                  const double ax = blk[Y::oX + j] - f1[q];  // cgmres.hpp:173-174, gmres.hpp:48
                  w[q] = ax * inv_h;
                }
              }
            }
            CG_PIPE_LAP(t_w);
            // modified Gram-Schmidt (gmres.hpp:52-58)
            double hc[km + 2];
#pragma unroll
            for (int i = 0; i < km + 2; i++) hc[i] = 0.0;
#pragma unroll
            for (int i = 0; i <= k; i++) {
              double c[Q];
              vec_load(1 + i, tslot, scr, c);
              double hik;
              if (EXACT) {
                double pr[Q];
#pragma unroll
                for (int q = 0; q < Q; q++) pr[q] = c[q] * w[q];
                hik = seq_sum(blk, pr);
              } else {
                double part = 0.0;
#pragma unroll
                for (int q = 0; q < Q; q++) part += c[q] * w[q];  // slots beyond L hold zeros in both
                hik = fast::warp_sum(part);
              }
              hc[i] = hik;
#pragma unroll
              for (int q = 0; q < Q; q++) {
                const double t = c[q] * hik;
                w[q] = w[q] - t;
              }
            }
            double part = 0.0;
            if (EXACT) {
              double pr[Q];
#pragma unroll
              for (int q = 0; q < Q; q++) pr[q] = w[q] * w[q];
              part = seq_sum(blk, pr);
            } else {
#pragma unroll
              for (int q = 0; q < Q; q++) part += w[q] * w[q];
              part = fast::warp_sum(part);
            }
            const double hn = sqrt(part);  // gmres.hpp:59-60
            CG_PIPE_LAP(t_mgs);
            int code = EXIT_FULL, ncol = k;
            bool solving = true;
            // U for the next sweep's input U + h*v_{k+1}: requested now, consumed after the reflector / residual
            // scalars below, whose dependent sqrt / reciprocal chain hides the L2 round trip
            double un[Q];
            if (k + 1 < km) {
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
                vec_store(2 + k, tslot, scr, w);  // v_{k+1}
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
              const double sg = (ha < 0.0) ? -1.0 : 1.0;
              const double buf = -sg * sqrt((0.0 + ha * ha) + hb * hb);
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
            CG_PIPE_LAP(t_hh);
            if (solving && k + 1 < km) {  // input of the next sweep (cgmres.hpp:168-169)
#pragma unroll
              for (int q = 0; q < Q; q++) {
                const int j = lane + 32 * q;
                if (j < L) {
                  const double t = w[q] * hh;
                  blk[Y::oX + j] = t + un[q];
                }
              }
            }
            CG_PIPE_LAP(t_fx);
            __syncwarp();
          }
          if (k + 1 < km) {
            bar_arrive(BX(g), T);
          } else {
            // the final update of this round (back substitution, U, dUdt, outputs) is deferred to the start of
            // the next step of this group, where it overlaps the serial warp's first pass; keep x for its plant step
            if (has && lane < nx) sc[Y::sXO + lane] = sc[Y::sX + lane];
            __syncwarp();
            if (more) {  // next round's state in, so that the serial warp never waits for a whole round to drain
              CG_PIPE_LAP(t_hh);
              state_in(r_next, g);
              CG_PIPE_LAP(t_si);
              bar_arrive(BX(g), T);
            }
          }
        }
      }
      if (!more) {  // last round of this CTA: nothing left to overlap the final updates with
        for (int g = 0; g < NG; g++) final_update(r, g);
      }
    }
  }

#ifdef CG_PIPE_TIMING
  if (a.dbg && blockIdx.x == 0 && lane == 0 && wid < 24) {  // [2*wid] = cycles blocked, [2*wid+1] = total
    a.dbg[2 * wid] = t_wait;
    a.dbg[2 * wid + 1] = clock64() - t_begin;
    if (wid == 0) {  // serial warp of group 0: cycles inside the first pass, the second pass, the Arnoldi sweeps
      a.dbg[48] = t_p1;
      a.dbg[49] = t_p2;
      a.dbg[50] = t_sw;
    }
    if (wid == 1) {  // vector warp 0: cycles per phase of its steps (both groups, all rounds)
      a.dbg[51] = t_fin;
      a.dbg[52] = t_v1;
      a.dbg[53] = t_v2;
      a.dbg[54] = t_dh;
      a.dbg[55] = t_w;
      a.dbg[56] = t_mgs;
      a.dbg[57] = t_hh;
      a.dbg[58] = t_fx;
      a.dbg[59] = t_si;
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

}  // namespace pipe
}  // namespace cgmres_b200
