// capi.cu -- the C ABI (include/cgmres_b200.h): handle bookkeeping, host<->device
// staging and stream-ordered launches.  No arithmetic of the control law lives here
// except get_dtau (cgmres.hpp:32-34), which is batch-uniform and evaluated once per
// step with the host libm so that it is bit-identical to the reference's.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <omp.h>

#include <atomic>
#include <new>
#include <string>
#include <vector>

#include "cgmres_b200.h"
#include "cgmres_b200/models.hpp"
#include "cgmres_b200/plant.hpp"
#include "kernel_args.h"

namespace cgmres_b200 {

template <class M, class Sim>
static ModelInfo make_info(const char* name) {
  ModelInfo i;
  i.dim_x = M::dim_x;
  i.dim_u = M::dim_u;
  i.dim_p = M::dim_p;
  i.dv = M::dv;
  i.k_max = M::k_max;
  i.n_ctrl = M::control_input;
  i.dt = M::dt;
  i.h = M::h;
  i.zeta = M::zeta;
  i.Tf = M::Tf;
  i.alpha = M::alpha;
  i.tol = M::tol;
  i.plant_dt = Sim::dt;
  i.name = name;
  return i;
}

const ModelInfo* model_info(int model) {
  static const ModelInfo infos[MODEL_COUNT] = {
      make_info<MassSpringDamperModel, MassSpringDamperSimulator>("mass_spring_damper"),
      make_info<ArmPendulumModel, ArmPendulumSimulator>("arm_type_inverted_pendulum"),
      make_info<SemiactiveDamperModel, SemiactiveDamperSimulator>("semiactive_damper"),
  };
  return (model >= 0 && model < MODEL_COUNT) ? &infos[model] : nullptr;
}

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(CGMRES_B200_ECUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define CU(call)                                       \
  do {                                                 \
    cudaError_t e_ = (call);                           \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

}  // namespace cgmres_b200

using namespace cgmres_b200;

static int sm_count_of(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
  return v;
}

struct cgmres_b200_controller {
  int model = 0, mode = 0, device = 0;
  int64_t n = 0, ld = 0;
  // set when a multi-stream control() failed half way: part of the batch may be one step ahead of the rest, so
  // every later call on the handle is refused (destroy still works)
  bool failed = false;
  const ModelInfo* mi = nullptr;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // pipelined host-buffer control(): slices of the batch on side streams so that one slice's PCIe copies overlap
  // another slice's kernel (instance-major modes only; instances are independent, so slicing changes nothing)
  static constexpr int kSlices = 5;
  cudaStream_t side[kSlices] = {};
  cudaEvent_t ev_begin = nullptr, ev_done[kSlices] = {};
  double t = 0.0;          // cgmres.hpp:195 -- all instances of a handle step in lock step
  bool ptau_full = false;  // false: ptau holds one p per instance (set_ptau_repeat)
  int integrator = PLANT_EULER;  // plant step of step_closed_loop: the reference's Euler, or RK4
  double *x = nullptr, *U = nullptr, *dUdt = nullptr, *ptau = nullptr, *F1 = nullptr, *V = nullptr, *xtau = nullptr,
         *u_out = nullptr;
  int32_t* status = nullptr;
  double* t_inst = nullptr;  // per-instance controller clocks (set_t); null = lock step with `t`
  long long* dbg = nullptr;  // 64 phase timestamps of one warp (debug builds of the on-chip kernel)
  double* scratch = nullptr;       // fast mode: spill regions of the pipelined kernel, one per concurrent launch
  size_t scratch_region = 0;       // doubles per region
  double* stage = nullptr;  // instance-major staging, grown on demand
  size_t stage_doubles = 0;
  // third-generation persistent kernel (pipe2_update.cuh) behind MODE_FAST / MODE_PIPELINED_EXACT; generation 2
  // (pipe_update.cuh) stays reachable for A/B runs with CGMRES_B200_PIPE_GEN=2 in the environment at create time
  int pipe_gen = 3;
  static constexpr int kMaxFusedSteps = 256;  // closed-loop steps per multi-step launch (horizon-ramp table rows)
  double* dtau_tab = nullptr;                 // [2][kMaxFusedSteps][2] device table, double-buffered across launches
  int dtau_tab_next = 0;
  double *x_log = nullptr, *u_log = nullptr;  // device trajectory log of step_closed_loop_log, grown on demand
  size_t log_steps = 0;

  bool soa() const { return mode == CGMRES_B200_MODE_EXACT; }
  // the modes whose kernel advances several steps of the same instances per launch
  bool fused_steps() const {
    return pipe_gen == 3 && (mode == CGMRES_B200_MODE_FAST || mode == CGMRES_B200_MODE_PIPELINED_EXACT);
  }  // exact: [element][instance]; on-chip kernels: [instance][element]
  int L() const { return mi->L(); }
  int ptau_rows_full() const { return (mi->dv + 1) * mi->dim_p; }

  int ensure_stage(size_t doubles) {
    if (doubles <= stage_doubles) return 0;
    if (stage) {
      CU(cudaStreamSynchronize(stream));
      CU(cudaFree(stage));
      stage = nullptr;
      stage_doubles = 0;
    }
    CU(cudaMalloc(&stage, sizeof(double) * doubles));
    stage_doubles = doubles;
    return 0;
  }

  template <class T>
  int dalloc(T** p, size_t count) {
    cudaError_t e = cudaMalloc(p, sizeof(T) * (count ? count : 1));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    e = cudaMemsetAsync(*p, 0, sizeof(T) * (count ? count : 1), stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    return 0;
  }

  int allocate() {
    const size_t l = (size_t)ld, Lz = (size_t)L();
    int rc;
    if (!soa()) {  // on-chip kernels: only the persistent state lives in HBM, instance-major like the ABI
      if ((rc = dalloc(&x, l * mi->dim_x))) return rc;
      if ((rc = dalloc(&U, l * Lz))) return rc;
      if ((rc = dalloc(&dUdt, l * Lz))) return rc;
      if ((rc = dalloc(&ptau, l * (size_t)ptau_rows_full()))) return rc;
      if ((rc = dalloc(&u_out, l * mi->dim_u))) return rc;
      if ((rc = dalloc(&status, l))) return rc;
      if ((rc = dalloc(&dbg, 64))) return rc;
      // spill regions of the pipelined kernel: region 0 for full-batch launches, 1..kSlices for the slices
      scratch_region = (mode == CGMRES_B200_MODE_FAST)              ? fast_scratch_doubles(model, device, n)
                       : (mode == CGMRES_B200_MODE_PIPELINED_EXACT) ? pipelined_exact_scratch_doubles(model, device, n)
                                                                    : 0;
      if (fused_steps()) {
        const size_t s3 = pipe2_scratch_doubles(model, device, n, mode == CGMRES_B200_MODE_PIPELINED_EXACT);
        scratch_region = s3 > scratch_region ? s3 : scratch_region;
        if ((rc = dalloc(&dtau_tab, (size_t)2 * kMaxFusedSteps * 2))) return rc;
      }
      if (scratch_region && (rc = dalloc(&scratch, scratch_region * (size_t)(kSlices + 1)))) return rc;
      if ((rc = ensure_stage((size_t)n * (size_t)(mi->dim_x + mi->dim_u + mi->dim_p + 1)))) return rc;
      return 0;
    }
    if ((rc = dalloc(&x, l * mi->dim_x))) return rc;
    if ((rc = dalloc(&U, l * Lz))) return rc;
    if ((rc = dalloc(&dUdt, l * Lz))) return rc;  // zero: the de-facto contract of cgmres.hpp:14
    if ((rc = dalloc(&ptau, l * (size_t)ptau_rows_full()))) return rc;
    if ((rc = dalloc(&F1, l * Lz))) return rc;
    if ((rc = dalloc(&V, l * Lz * (size_t)(mi->k_max + 1)))) return rc;
    if ((rc = dalloc(&xtau, 3 * l * (size_t)mi->dim_x * (size_t)(mi->dv > 1 ? mi->dv - 1 : 1)))) return rc;  // 3 planes
    if ((rc = dalloc(&u_out, l * mi->dim_u))) return rc;
    if ((rc = dalloc(&status, l))) return rc;
    if ((rc = ensure_stage((size_t)n * (size_t)(mi->dim_x + mi->dim_u + mi->dim_p + 1)))) return rc;
    return 0;
  }

  void release() {
    cudaFree(x);
    cudaFree(U);
    cudaFree(dUdt);
    cudaFree(ptau);
    cudaFree(F1);
    cudaFree(V);
    cudaFree(xtau);
    cudaFree(u_out);
    cudaFree(status);
    cudaFree(t_inst);
    cudaFree(dbg);
    cudaFree(scratch);
    cudaFree(stage);
    cudaFree(dtau_tab);
    cudaFree(x_log);
    cudaFree(u_log);
    for (int i = 0; i < kSlices; i++) {
      if (side[i]) cudaStreamDestroy(side[i]);
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
    }
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (own_stream) cudaStreamDestroy(own_stream);
  }

  int ensure_side_streams() {
    if (side[0]) return 0;
    CU(cudaEventCreateWithFlags(&ev_begin, cudaEventDisableTiming));
    for (int i = 0; i < kSlices; i++) {
      CU(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    }
    return 0;
  }

  // on-chip modes: one update of instances [lo, lo+cnt) on stream s (pointer offsets into the instance-major state)
  int launch_slice(int64_t lo, int64_t cnt, int plant, double dt_t, double dt_th, cudaStream_t s, int region = 0,
                   int n_steps = 1, const double* tab = nullptr, double* xl = nullptr, double* ul = nullptr) {
    FastArgs f;
    f.n_steps = n_steps;
    f.dtau_tab = tab;
    f.x_log = xl;
    f.u_log = ul;
    const int64_t prow = ptau_full ? (int64_t)ptau_rows_full() : (int64_t)mi->dim_p;
    f.n = cnt;
    f.x = x + lo * mi->dim_x;
    f.U = U + lo * L();
    f.dUdt = dUdt + lo * L();
    f.ptau = ptau + lo * prow;
    f.u_out = u_out + lo * mi->dim_u;
    f.status = status + lo;
    f.dtau_t = dt_t;
    f.dtau_th = dt_th;
    f.t_inst = t_inst ? t_inst + lo : nullptr;
    f.plant = plant;
    f.dbg = dbg;
    f.scratch = scratch ? scratch + scratch_region * (size_t)region : nullptr;
    if (mode == CGMRES_B200_MODE_FAST) {
      // small single-step batches: the first-generation kernel's dependent chain per update is the shortest
      const bool big = cnt > (int64_t)onchip_exact_instances_per_cta(model) * sm_count_of(device);
      if (pipe_gen == 3 && (big || n_steps > 1 || xl || ul))
        CU(pipe2_fast_launch_control(model, ptau_full, f, s));
      else
        CU(fast_launch_control(model, ptau_full, f, s));
    } else if (mode == CGMRES_B200_MODE_PIPELINED_EXACT) {
      if (pipe_gen == 3)
        CU(pipe2_exact_launch_control(model, ptau_full, f, s));
      else
        CU(pipelined_exact_launch_control(model, ptau_full, f, s));
    } else {
      CU(onchip_exact_launch_control(model, ptau_full, f, s));
    }
    g_launches++;
    return 0;
  }

  // n_steps closed-loop steps of the whole batch in ONE launch of the third-generation kernel (fused_steps() modes):
  // the horizon ramps of every step are evaluated here with the host libm, exactly like the per-step path
  int launch_fused(int n_steps, int plant, double* xl, double* ul) {
    if (n_steps <= 0 || n == 0) return 0;
    std::vector<double> tab((size_t)2 * n_steps);
    double tt = t;
    for (int s = 0; s < n_steps; s++) {
      tab[2 * s] = dtau(tt);
      tab[2 * s + 1] = dtau(tt + mi->h);
      tt = tt + mi->dt;  // cgmres.hpp:107 (accumulated, not i*dt)
    }
    double* dtab = dtau_tab + (size_t)dtau_tab_next * kMaxFusedSteps * 2;
    dtau_tab_next ^= 1;
    CU(cudaMemcpyAsync(dtab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, stream));
    int rc = launch_slice(0, n, plant, tab[0], tab[1], stream, 0, n_steps, dtab, xl, ul);
    if (rc) return rc;
    t = tt;
    return 0;
  }

  int ensure_log(size_t steps) {
    if (steps <= log_steps) return 0;
    if (x_log) {
      CU(cudaStreamSynchronize(stream));
      CU(cudaFree(x_log));
      CU(cudaFree(u_log));
      x_log = u_log = nullptr;
      log_steps = 0;
    }
    CU(cudaMalloc(&x_log, sizeof(double) * steps * (size_t)(n ? n : 1) * mi->dim_x));
    CU(cudaMalloc(&u_log, sizeof(double) * steps * (size_t)(n ? n : 1) * mi->dim_u));
    log_steps = steps;
    return 0;
  }

  double dtau(double tt) const { return mi->Tf * (1 - exp(-mi->alpha * tt)) / (double)mi->dv; }

  // one update for every instance; plant=1 also advances x (closed loop on device)
  int launch_update(int plant) {
    if (!soa()) {
      int rc = launch_slice(0, n, plant, dtau(t), dtau(t + mi->h), stream);
      if (rc) return rc;
      t = t + mi->dt;
      return 0;
    }
    ExactArgs a;
    a.n = n;
    a.ld = ld;
    a.x = x;
    a.U = U;
    a.dUdt = dUdt;
    a.ptau = ptau;
    a.F1 = F1;
    a.V = V;
    a.xtau = xtau;
    a.u_out = u_out;
    a.status = status;
    a.dtau_t = dtau(t);
    a.dtau_th = dtau(t + mi->h);
    a.t_inst = t_inst;
    a.plant = plant;
    CU(exact_launch_control(model, ptau_full, a, stream));
    g_launches++;
    t = t + mi->dt;  // cgmres.hpp:107 (accumulated, not i*dt)
    return 0;
  }

  // host instance-major -> device SoA rows
  int upload_rows(const double* host, double* dst, int rows) {
    if (rows == 0 || n == 0) return 0;
    if (!soa()) {
      CU(cudaMemcpyAsync(dst, host, sizeof(double) * (size_t)n * rows, cudaMemcpyHostToDevice, stream));
      return 0;
    }
    double* soa = dst;
    int rc = ensure_stage((size_t)n * rows);
    if (rc) return rc;
    CU(cudaMemcpyAsync(stage, host, sizeof(double) * (size_t)n * rows, cudaMemcpyHostToDevice, stream));
    CU(launch_aos_to_soa(stage, soa, n, rows, ld, stream));
    g_launches++;
    return 0;
  }
  int download_rows(double* host, const double* src, int rows) {
    if (rows == 0 || n == 0) return 0;
    if (!soa()) {
      CU(cudaMemcpyAsync(host, src, sizeof(double) * (size_t)n * rows, cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      return 0;
    }
    const double* soa = src;
    int rc = ensure_stage((size_t)n * rows);
    if (rc) return rc;
    CU(launch_soa_to_aos(soa, stage, n, rows, ld, stream));
    g_launches++;
    CU(cudaMemcpyAsync(host, stage, sizeof(double) * (size_t)n * rows, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return 0;
  }
};

#define CHECK_H(h)                                               \
  if (!(h)) return fail(CGMRES_B200_EINVAL, "null handle");      \
  if ((h)->failed)                                               \
  return fail(CGMRES_B200_ECUDA, "handle is unusable: an earlier sliced control() failed part way (destroy it)")
#define ON_DEVICE(h) CU(cudaSetDevice((h)->device))

template <class Sim>
static void plant_host(int64_t n, double* x, const double* u) {  // Euler: mul(dxdt,dxdt,dt); add(x,x,dxdt)
  constexpr int nx = Sim::dim_x, nu = Sim::dim_u;
  // instances are independent: big batches are split over a few host threads (OMP_NUM_THREADS caps it; torchrun
  // sets that to 1 per rank)
  const int cap = omp_get_max_threads() < 8 ? omp_get_max_threads() : 8;
#pragma omp parallel for schedule(static) num_threads(cap) if (n >= 16384)
  for (int64_t i = 0; i < n; i++) plant_euler<Sim>(x + i * nx, u + i * nu);
}

extern "C" {

const char* cgmres_b200_last_error(void) { return g_err.c_str(); }

int cgmres_b200_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return c;
}

int cgmres_b200_model_dims(int model, int* dims) {
  const ModelInfo* mi = model_info(model);
  if (!mi || !dims) return fail(CGMRES_B200_EINVAL, "unknown model");
  dims[0] = mi->dim_x;
  dims[1] = mi->dim_u;
  dims[2] = mi->dim_p;
  dims[3] = mi->dv;
  dims[4] = mi->k_max;
  dims[5] = mi->n_ctrl;
  return 0;
}

int cgmres_b200_model_params(int model, double* par) {
  const ModelInfo* mi = model_info(model);
  if (!mi || !par) return fail(CGMRES_B200_EINVAL, "unknown model");
  par[0] = mi->dt;
  par[1] = mi->h;
  par[2] = mi->zeta;
  par[3] = mi->Tf;
  par[4] = mi->alpha;
  par[5] = mi->tol;
  return 0;
}

const char* cgmres_b200_model_name(int model) {
  const ModelInfo* mi = model_info(model);
  return mi ? mi->name : nullptr;
}

int cgmres_b200_create(int model, int64_t n, int device, int mode, cgmres_b200_handle* out) {
  if (!out) return fail(CGMRES_B200_EINVAL, "out is null");
  *out = nullptr;
  const ModelInfo* mi = model_info(model);
  if (!mi) return fail(CGMRES_B200_EINVAL, "unknown model id");
  if (n < 0) return fail(CGMRES_B200_EINVAL, "negative instance count");
  if (mode != CGMRES_B200_MODE_EXACT && mode != CGMRES_B200_MODE_FAST && mode != CGMRES_B200_MODE_ONCHIP_EXACT &&
      mode != CGMRES_B200_MODE_PIPELINED_EXACT)
    return fail(CGMRES_B200_EINVAL, "unknown mode");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(CGMRES_B200_ECUDA, "no usable CUDA device (this library has no CPU path)");
  }
  if (device < 0 || device >= count) return fail(CGMRES_B200_EINVAL, "device index out of range");
  CU(cudaSetDevice(device));
  cgmres_b200_controller* h = new (std::nothrow) cgmres_b200_controller();
  if (!h) return fail(CGMRES_B200_ENOMEM, "host allocation failed");
  h->model = model;
  h->mode = mode;
  h->device = device;
  h->n = n;
  h->ld = (n + 31) & ~(int64_t)31;
  if (h->ld == 0) h->ld = 32;
  h->mi = mi;
  if (const char* gen = getenv("CGMRES_B200_PIPE_GEN")) h->pipe_gen = (atoi(gen) == 2) ? 2 : 3;
  e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return cuda_fail(e, "cudaStreamCreate");
  }
  h->stream = h->own_stream;
  int rc = h->allocate();
  if (rc == 0) {
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (rc) {
    h->release();
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int cgmres_b200_destroy(cgmres_b200_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->release();
  delete h;
  return 0;
}

int64_t cgmres_b200_size(cgmres_b200_handle h) { return h ? h->n : -1; }
int cgmres_b200_model(cgmres_b200_handle h) { return h ? h->model : -1; }
int cgmres_b200_mode(cgmres_b200_handle h) { return h ? h->mode : -1; }

int cgmres_b200_set_stream(cgmres_b200_handle h, void* stream) {
  CHECK_H(h);
  ON_DEVICE(h);
  CU(cudaStreamSynchronize(h->stream));
  h->stream = stream ? (cudaStream_t)stream : h->own_stream;
  return 0;
}
void* cgmres_b200_get_stream(cgmres_b200_handle h) { return h ? (void*)h->stream : nullptr; }

int cgmres_b200_synchronize(cgmres_b200_handle h) {
  CHECK_H(h);
  ON_DEVICE(h);
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

double cgmres_b200_get_dtau(cgmres_b200_handle h, double t) { return h ? h->dtau(t) : 0.0; }

int cgmres_b200_set_ptau(cgmres_b200_handle h, const double* ptau) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (h->mi->dim_p == 0) return 0;
  if (!ptau) return fail(CGMRES_B200_EINVAL, "ptau is null");
  int rc = h->upload_rows(ptau, h->ptau, h->ptau_rows_full());
  if (rc) return rc;
  h->ptau_full = true;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_set_ptau_repeat(cgmres_b200_handle h, const double* p) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (h->mi->dim_p == 0) return 0;
  if (!p) return fail(CGMRES_B200_EINVAL, "p is null");
  int rc = h->upload_rows(p, h->ptau, h->mi->dim_p);
  if (rc) return rc;
  h->ptau_full = false;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_init_u0(cgmres_b200_handle h, const double* u0) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!u0) return fail(CGMRES_B200_EINVAL, "u0 is null");
  const int nu = h->mi->dim_u;
  int rc = h->ensure_stage((size_t)h->n * nu);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->stage, u0, sizeof(double) * (size_t)h->n * nu, cudaMemcpyHostToDevice, h->stream));
  if (h->soa())
    CU(launch_broadcast_rows(h->stage, h->U, h->n, nu, h->mi->dv, h->ld, h->stream));
  else
    CU(launch_broadcast_inst(h->stage, h->U, h->n, nu, h->mi->dv, h->stream));
  g_launches++;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_init_u0_newton(cgmres_b200_handle h, double* u0, const double* x0, const double* p0, int n_loop) {
  CHECK_H(h);
  ON_DEVICE(h);
  const int nu = h->mi->dim_u, nx = h->mi->dim_x, np = h->mi->dim_p;
  if (!u0 || !x0 || (np > 0 && !p0) || n_loop < 0) return fail(CGMRES_B200_EINVAL, "bad init_u0_newton arguments");
  const size_t n = (size_t)h->n;
  int rc = h->ensure_stage(n * (size_t)(nu + nx + np + 1));
  if (rc) return rc;
  double* d_u = h->stage;
  double* d_x = d_u + n * nu;
  double* d_p = d_x + n * nx;
  CU(cudaMemcpyAsync(d_u, u0, sizeof(double) * n * nu, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_x, x0, sizeof(double) * n * nx, cudaMemcpyHostToDevice, h->stream));
  if (np > 0) CU(cudaMemcpyAsync(d_p, p0, sizeof(double) * n * np, cudaMemcpyHostToDevice, h->stream));
  CU(exact_launch_newton(h->model, h->n, h->soa() ? h->ld : 1, h->soa() ? 1 : (int64_t)h->L(), d_u, d_x, d_p, np,
                         n_loop, h->U, h->stream));
  g_launches++;
  CU(cudaMemcpyAsync(u0, d_u, sizeof(double) * n * nu, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_control_dev(cgmres_b200_handle h, double* u_dev, const double* x_dev) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!u_dev || !x_dev) return fail(CGMRES_B200_EINVAL, "null device pointer");
  if (!h->soa()) {  // instance-major state == ABI layout: plain device copies
    const size_t n = (size_t)h->n;
    if (n == 0) return 0;
    CU(cudaMemcpyAsync(h->x, x_dev, sizeof(double) * n * h->mi->dim_x, cudaMemcpyDeviceToDevice, h->stream));
    int rc = h->launch_update(0);
    if (rc) return rc;
    CU(cudaMemcpyAsync(u_dev, h->u_out, sizeof(double) * n * h->mi->dim_u, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
  }
  CU(launch_aos_to_soa(x_dev, h->x, h->n, h->mi->dim_x, h->ld, h->stream));
  g_launches++;
  int rc = h->launch_update(0);
  if (rc) return rc;
  CU(launch_soa_to_aos(h->u_out, u_dev, h->n, h->mi->dim_u, h->ld, h->stream));
  g_launches++;
  return 0;
}

int cgmres_b200_control(cgmres_b200_handle h, double* u, const double* x) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!u || !x) return fail(CGMRES_B200_EINVAL, "null pointer");
  const size_t n = (size_t)h->n;
  const int nx = h->mi->dim_x, nu = h->mi->dim_u;
  if (!h->soa()) {  // instance-major state == ABI layout: x straight into place, u straight out
    if (n == 0) return 0;
    // slices are whole "waves" (the instances all SMs hold at once) so that no slice ends in a partly filled wave
    const int per_cta = h->fused_steps() ? pipe2_instances_per_cta(h->model, h->mode == CGMRES_B200_MODE_PIPELINED_EXACT)
                                          : onchip_instances_per_cta(h->model, h->mode);
    const int64_t wave = (int64_t)per_cta * sm_count_of(h->device);
    const int64_t waves = (h->n + wave - 1) / wave;
    const int S = (h->n >= 8192 && waves >= 2)
                      ? (int)(waves < cgmres_b200_controller::kSlices ? waves : cgmres_b200_controller::kSlices)
                      : 1;
    if (S == 1) {
      CU(cudaMemcpyAsync(h->x, x, sizeof(double) * n * nx, cudaMemcpyHostToDevice, h->stream));
      int rcu = h->launch_update(0);
      if (rcu) return rcu;
      CU(cudaMemcpyAsync(u, h->u_out, sizeof(double) * n * nu, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      return 0;
    }
    // large batches: S slices, each {H2D x, update, D2H u} on its own stream, all ordered after the handle's stream
    int rcs = h->ensure_side_streams();
    if (rcs) return rcs;
    const double dt_t = h->dtau(h->t), dt_th = h->dtau(h->t + h->mi->h);
    CU(cudaEventRecord(h->ev_begin, h->stream));
    // slice boundaries in waves: with enough waves the first and the last slice are ONE wave, so that the copies
    // nothing can overlap (x of the first slice in, u of the last slice out) are as small as possible
    int64_t wb[cgmres_b200_controller::kSlices + 1];
    if (S >= 3 && waves >= 2 + (S - 2)) {
      wb[0] = 0;
      wb[1] = 1;
      for (int i = 2; i < S; i++) wb[i] = 1 + (waves - 2) * (i - 1) / (S - 2);
      wb[S] = waves;
    } else {
      for (int i = 0; i <= S; i++) wb[i] = waves * i / S;
    }
    // anything that fails inside the loop leaves earlier slices already running: drain them and poison the handle
    auto abandon = [&](int enqueued, int rc_) {
      for (int k = 0; k < enqueued; k++) cudaStreamSynchronize(h->side[k]);
      h->failed = true;
      return rc_;
    };
#define CU_SLICE(call)                                                     \
  do {                                                                     \
    cudaError_t e_ = (call);                                               \
    if (e_ != cudaSuccess) return abandon(i + 1, cuda_fail(e_, #call));    \
  } while (0)
    for (int i = 0; i < S; i++) {
      const int64_t lo = wb[i] * wave;
      const int64_t hi_ = wb[i + 1] * wave;
      const int64_t hi = hi_ < h->n ? hi_ : h->n;
      const int64_t cnt = hi - lo;
      cudaStream_t st = h->side[i];
      CU_SLICE(cudaStreamWaitEvent(st, h->ev_begin, 0));
      CU_SLICE(cudaMemcpyAsync(h->x + lo * nx, x + lo * nx, sizeof(double) * (size_t)cnt * nx, cudaMemcpyHostToDevice,
                               st));
      int rcu = h->launch_slice(lo, cnt, 0, dt_t, dt_th, st, 1 + i);
      if (rcu) return abandon(i + 1, rcu);
      CU_SLICE(cudaMemcpyAsync(u + lo * nu, h->u_out + lo * nu, sizeof(double) * (size_t)cnt * nu,
                               cudaMemcpyDeviceToHost, st));
      CU_SLICE(cudaEventRecord(h->ev_done[i], st));
      CU_SLICE(cudaStreamWaitEvent(h->stream, h->ev_done[i], 0));  // later work on the handle's stream sees the slices
    }
#undef CU_SLICE
    h->t = h->t + h->mi->dt;
    for (int i = 0; i < S; i++) CU(cudaEventSynchronize(h->ev_done[i]));
    return 0;
  }
  int rc = h->ensure_stage(n * (size_t)(nx + nu));
  if (rc) return rc;
  double* d_x = h->stage;
  double* d_u = h->stage + n * nx;
  CU(cudaMemcpyAsync(d_x, x, sizeof(double) * n * nx, cudaMemcpyHostToDevice, h->stream));
  rc = cgmres_b200_control_dev(h, d_u, d_x);
  if (rc) return rc;
  CU(cudaMemcpyAsync(u, d_u, sizeof(double) * n * nu, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_set_x(cgmres_b200_handle h, const double* x) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!x) return fail(CGMRES_B200_EINVAL, "x is null");
  int rc = h->upload_rows(x, h->x, h->mi->dim_x);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}
int cgmres_b200_get_x(cgmres_b200_handle h, double* x) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!x) return fail(CGMRES_B200_EINVAL, "x is null");
  return h->download_rows(x, h->x, h->mi->dim_x);
}
int cgmres_b200_get_u(cgmres_b200_handle h, double* u) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!u) return fail(CGMRES_B200_EINVAL, "u is null");
  return h->download_rows(u, h->u_out, h->mi->dim_u);
}

int cgmres_b200_step_closed_loop(cgmres_b200_handle h, int n_steps) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (n_steps < 0) return fail(CGMRES_B200_EINVAL, "negative step count");
  if (h->fused_steps() && n_steps > 1) {  // the same resident instances advance several steps per launch
    for (int done = 0; done < n_steps;) {
      const int chunk = (n_steps - done) < cgmres_b200_controller::kMaxFusedSteps
                            ? (n_steps - done)
                            : cgmres_b200_controller::kMaxFusedSteps;
      int rc = h->launch_fused(chunk, h->integrator, nullptr, nullptr);
      if (rc) return rc;
      done += chunk;
    }
    return 0;
  }
  for (int s = 0; s < n_steps; s++) {
    int rc = h->launch_update(h->integrator);
    if (rc) return rc;
  }
  return 0;
}

int cgmres_b200_step_closed_loop_log(cgmres_b200_handle h, int n_steps, double* x_log, double* u_log) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (n_steps < 0 || (n_steps > 0 && (!x_log || !u_log))) return fail(CGMRES_B200_EINVAL, "bad log arguments");
  const size_t n = (size_t)h->n;
  if (n == 0 || n_steps == 0) return 0;
  const int nx = h->mi->dim_x, nu = h->mi->dim_u;
  // chunks of the trajectory stay on the device until the chunk is complete (at most ~64 MB of log per chunk)
  size_t per_step = n * (size_t)(nx + nu) * sizeof(double);
  int chunk_max = (int)((size_t)(64u << 20) / (per_step ? per_step : 1));
  if (chunk_max < 1) chunk_max = 1;
  if (chunk_max > cgmres_b200_controller::kMaxFusedSteps) chunk_max = cgmres_b200_controller::kMaxFusedSteps;
  for (int done = 0; done < n_steps;) {
    const int chunk = (n_steps - done) < chunk_max ? (n_steps - done) : chunk_max;
    int rc = h->ensure_log((size_t)chunk);
    if (rc) return rc;
    if (h->fused_steps()) {
      rc = h->launch_fused(chunk, h->integrator, h->x_log, h->u_log);
      if (rc) return rc;
    } else {  // one launch per step; the step's x and u are appended to the device log by stream-ordered copies
      for (int s = 0; s < chunk; s++) {
        rc = h->launch_update(h->integrator);
        if (rc) return rc;
        if (h->soa()) {
          CU(launch_soa_to_aos(h->x, h->x_log + (size_t)s * n * nx, h->n, nx, h->ld, h->stream));
          CU(launch_soa_to_aos(h->u_out, h->u_log + (size_t)s * n * nu, h->n, nu, h->ld, h->stream));
          g_launches += 2;
        } else {
          CU(cudaMemcpyAsync(h->x_log + (size_t)s * n * nx, h->x, sizeof(double) * n * nx, cudaMemcpyDeviceToDevice,
                             h->stream));
          CU(cudaMemcpyAsync(h->u_log + (size_t)s * n * nu, h->u_out, sizeof(double) * n * nu,
                             cudaMemcpyDeviceToDevice, h->stream));
        }
      }
    }
    CU(cudaMemcpyAsync(x_log + (size_t)done * n * nx, h->x_log, sizeof(double) * (size_t)chunk * n * nx,
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(u_log + (size_t)done * n * nu, h->u_log, sizeof(double) * (size_t)chunk * n * nu,
                       cudaMemcpyDeviceToHost, h->stream));
    done += chunk;
  }
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_set_t(cgmres_b200_handle h, const double* t) {
  CHECK_H(h);
  ON_DEVICE(h);
  CU(cudaStreamSynchronize(h->stream));
  if (!t) {  // back to lock step
    if (h->t_inst) CU(cudaFree(h->t_inst));
    h->t_inst = nullptr;
    return 0;
  }
  if (!h->t_inst) CU(cudaMalloc(&h->t_inst, sizeof(double) * (size_t)(h->ld ? h->ld : 1)));
  CU(cudaMemcpyAsync(h->t_inst, t, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_get_t(cgmres_b200_handle h, double* t) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!t) return fail(CGMRES_B200_EINVAL, "t is null");
  if (!h->t_inst) {
    for (int64_t i = 0; i < h->n; i++) t[i] = h->t;
    return 0;
  }
  CU(cudaMemcpyAsync(t, h->t_inst, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_set_plant_integrator(cgmres_b200_handle h, int integrator) {
  CHECK_H(h);
  if (integrator != CGMRES_B200_PLANT_EULER && integrator != CGMRES_B200_PLANT_RK4)
    return fail(CGMRES_B200_EINVAL, "unknown plant integrator");
  h->integrator = integrator;
  return 0;
}

int cgmres_b200_get_state(cgmres_b200_handle h, double* t, double* U, double* dUdt) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (t) *t = h->t;
  int rc = 0;
  if (U && (rc = h->download_rows(U, h->U, h->L()))) return rc;
  if (dUdt && (rc = h->download_rows(dUdt, h->dUdt, h->L()))) return rc;
  return 0;
}

int cgmres_b200_set_state(cgmres_b200_handle h, const double* t, const double* U, const double* dUdt) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (t) h->t = *t;
  int rc = 0;
  if (U && (rc = h->upload_rows(U, h->U, h->L()))) return rc;
  if (dUdt && (rc = h->upload_rows(dUdt, h->dUdt, h->L()))) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_get_status(cgmres_b200_handle h, int32_t* status) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!status) return fail(CGMRES_B200_EINVAL, "status is null");
  CU(cudaMemcpyAsync(status, h->status, sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cgmres_b200_plant_step_host(int model, int64_t n, double* x, const double* u) {
  if (n < 0 || (n > 0 && (!x || !u))) return fail(CGMRES_B200_EINVAL, "bad plant_step_host arguments");
  switch (model) {
    case MODEL_MSD: plant_host<MassSpringDamperSimulator>(n, x, u); return 0;
    case MODEL_ARM: plant_host<ArmPendulumSimulator>(n, x, u); return 0;
    case MODEL_SEMIACTIVE: plant_host<SemiactiveDamperSimulator>(n, x, u); return 0;
  }
  return fail(CGMRES_B200_EINVAL, "unknown model");
}

void cgmres_b200_portable_sincos(double x, double* s, double* c) { ptrig::psincos(x, s, c); }

int cgmres_b200_debug_phase_times(cgmres_b200_handle h, int64_t* out64) {
  CHECK_H(h);
  ON_DEVICE(h);
  if (!out64 || !h->dbg) return fail(CGMRES_B200_EINVAL, "no phase-timing buffer (exact mode or null pointer)");
  CU(cudaMemcpyAsync(out64, h->dbg, sizeof(long long) * 64, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t cgmres_b200_launch_count(void) { return g_launches.load(); }

}  // extern "C"
