/*
 * cgmres_oracle.c -- plain-C restatement of the reference C/GMRES controller
 * (blockahead/CGMRES_cpp), used ONLY as the parity checker and the "port" CPU
 * baseline.  TEST INFRASTRUCTURE: see cgmres_oracle.h for who may call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit for
 * bit against oracle/_ref/libcgmres_ref.so (the unmodified reference headers
 * compiled by oracle/ref_harness.cpp) and against the committed fixtures under
 * tests/golden/ that were generated from that library (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off (no -march=native, no -ffast-math): every
 * expression below is binary64 round-to-nearest, evaluated in the order the
 * reference's C++ parses it.
 *
 * All file:line citations are relative to the reference repository root.
 */
#define _POSIX_C_SOURCE 200809L
#include "cgmres_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* Problem definitions                                                       */
/* ------------------------------------------------------------------------- */

typedef void (*fn_xup)(double* out, const double* x, const double* u, const double* p);
typedef void (*fn_xp)(double* out, const double* x, const double* p);
typedef void (*fn_xupl)(double* out, const double* x, const double* u, const double* p, const double* lmd);
typedef void (*fn_plant)(double* out, const double* x, const double* u);

typedef struct {
  int dim_x, dim_u, dim_p, dv, k_max, n_ctrl;
  double dt, h, zeta, Tf, alpha, tol;
  double plant_dt;
  fn_xup f;       /* state equation          Model::dxdt   */
  fn_xp phix;     /* terminal cost gradient  Model::dPhidx */
  fn_xupl hx;     /* dH/dx                   Model::dHdx   */
  fn_xupl hu;     /* dH/du                   Model::dHdu   */
  fn_xupl huu;    /* d2H/du2 (column major)  Model::ddHduu */
  fn_plant plant; /* Simulator::dxdt                        */
} problem_t;

/* ---- mass_spring_damper: mass_spring_damper/model.hpp:4-125 (== multiple_controller/model1.hpp) ---- */
#define MSD_M1 1.0
#define MSD_M2 1.0
#define MSD_D1 1.0
#define MSD_D2 1.0
#define MSD_K1 1.0
#define MSD_K2 1.0
#define MSD_UC ((10.0 + -10.0) / 2.0) /* model.hpp:118-120 */
#define MSD_UR ((10.0 - -10.0) / 2.0) /* model.hpp:121 */

/* model.hpp:36-41.  NOTE the stiffness term is -(k1*k2)/m1 here but -(k1+k2)/m1 in dHdx. */
static void msd_f(double* o, const double* x, const double* u, const double* p) {
  (void)p;
  o[0] = x[2];
  o[1] = x[3];
  o[2] = -(MSD_K1 * MSD_K2) / MSD_M1 * x[0] + MSD_K2 / MSD_M1 * x[1] - (MSD_D1 + MSD_D2) / MSD_M1 * x[2] +
         MSD_D2 / MSD_M1 * x[3] + u[0] / MSD_M1;
  o[3] = MSD_K2 / MSD_M2 * x[0] - MSD_K2 / MSD_M2 * x[1] + MSD_D2 / MSD_M2 * x[2] - MSD_D2 / MSD_M2 * x[3] +
         u[1] / MSD_M2;
}
/* model.hpp:43-48, sf = (10,10,1,1) model.hpp:113 */
static void msd_phix(double* o, const double* x, const double* p) {
  o[0] = -(p[0] - x[0]) * 10.0;
  o[1] = -(p[1] - x[1]) * 10.0;
  o[2] = x[2] * 1.0;
  o[3] = x[3] * 1.0;
}
/* model.hpp:50-55, q = (1,1,10,10) model.hpp:114 */
static void msd_hx(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)u;
  o[0] = -(p[0] - x[0]) * 1.0 - (MSD_K1 + MSD_K2) / MSD_M1 * l[2] + MSD_K2 / MSD_M2 * l[3];
  o[1] = -(p[1] - x[1]) * 1.0 + MSD_K2 / MSD_M1 * l[2] - MSD_K2 / MSD_M2 * l[3];
  o[2] = x[2] * 10.0 + l[0] - (MSD_D1 + MSD_D2) / MSD_M1 * l[2] + MSD_D2 / MSD_M2 * l[3];
  o[3] = x[3] * 10.0 + l[1] + MSD_D2 / MSD_M1 * l[2] - MSD_D2 / MSD_M2 * l[3];
}
/* model.hpp:57-64, r = (0.1,0.1,0.01,0.01) model.hpp:115 */
static void msd_hu(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)x;
  (void)p;
  o[0] = 0.1 * u[0] + l[2] / MSD_M1 + 2.0 * u[4] * (u[0] - MSD_UC);
  o[1] = 0.1 * u[1] + l[3] / MSD_M2 + 2.0 * u[5] * (u[1] - MSD_UC);
  o[2] = -0.01 + 2.0 * u[4] * u[2];
  o[3] = -0.01 + 2.0 * u[5] * u[3];
  o[4] = (u[0] - MSD_UC) * (u[0] - MSD_UC) + u[2] * u[2] - MSD_UR * MSD_UR;
  o[5] = (u[1] - MSD_UC) * (u[1] - MSD_UC) + u[3] * u[3] - MSD_UR * MSD_UR;
}
/* model.hpp:66-108: 6x6, column major, only the listed entries are non-zero */
static void msd_huu(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)x;
  (void)p;
  (void)l;
  for (int i = 0; i < 36; i++) o[i] = 0.0;
  o[0] = 0.1 + 2 * u[4];
  o[4] = 2 * (u[0] - MSD_UC);
  o[7] = 0.1 + 2 * u[5];
  o[11] = 2 * (u[1] - MSD_UC);
  o[14] = 2 * u[4];
  o[16] = 2 * u[2];
  o[21] = 2 * u[5];
  o[23] = 2 * u[3];
  o[24] = 2 * (u[0] - MSD_UC);
  o[26] = 2 * u[2];
  o[31] = 2 * (u[1] - MSD_UC);
  o[33] = 2 * u[3];
}
/* mass_spring_damper/simulator.hpp:13-18 */
static void msd_plant(double* o, const double* x, const double* u) { msd_f(o, x, u, NULL); }

/* ---- arm_type_inverted_pendulum: arm_type_inverted_pendulum/model.hpp:5-99 (== multiple_controller/model2.hpp) ---- */
/* sin/cos: glibc like the reference, or -- in the "ptrig" build of this oracle -- the portable +,-,* implementation
 * that the GPU's exact modes use, so that those modes can be checked bit for bit on this model too. */
#ifdef ORACLE_PORTABLE_TRIG
#include "portable_trig.h"
#define sin(x) opt_sin(x)
#define cos(x) opt_cos(x)
void ORACLE_FN(sincos)(double x, double* s, double* c) { opt_sincos(x, s, c); }
#else
void ORACLE_FN(sincos)(double x, double* s, double* c) {
  *s = sin(x);
  *c = cos(x);
}
#endif
#define ARM_AS 6.25
#define ARM_BS 15.6
#define ARM_A52 39.1111
#define ARM_C22 0.0407448
#define ARM_A32A 5.65635
#define ARM_A32 0.905016
#define ARM_A32B 14.1183
#define ARM_UC ((3.0 + -3.0) / 2.0)
#define ARM_UR ((3.0 - -3.0) / 2.0)

/* model.hpp:37-42 */
static void arm_f(double* o, const double* x, const double* u, const double* p) {
  (void)p;
  o[0] = x[2];
  o[1] = x[3];
  o[2] = -ARM_AS * x[2] + ARM_BS * u[0];
  o[3] = ARM_A32 * x[2] * x[2] * sin(x[0] - x[1]) + ARM_A52 * sin(x[1]) - ARM_A32B * cos(x[0] - x[1]) * u[0] +
         ARM_A32A * cos(x[0] - x[1]) * x[2] + ARM_C22 * (x[2] - x[3]);
}
/* model.hpp:44-49, sf = (3,1,0,0) model.hpp:81 */
static void arm_phix(double* o, const double* x, const double* p) {
  o[0] = (x[0] - p[0]) * 3.0;
  o[1] = (x[1] - p[1]) * 1.0;
  o[2] = x[2] * 0.0;
  o[3] = x[3] * 0.0;
}
/* model.hpp:51-56, q = (1,1,0,0) model.hpp:82 */
static void arm_hx(double* o, const double* x, const double* u, const double* p, const double* l) {
  o[0] = (x[0] - p[0]) * 1.0 + l[3] * (ARM_A32 * x[2] * x[2] * cos(x[0] - x[1]) + ARM_A32B * sin(x[0] - x[1]) * u[0] -
                                       ARM_A32A * sin(x[0] - x[1]) * x[2]);
  o[1] = (x[1] - p[1]) * 1.0 + l[3] * (-ARM_A32 * x[2] * x[2] * cos(x[0] - x[1]) + ARM_A52 * cos(x[1]) -
                                       ARM_A32B * sin(x[0] - x[1]) * u[0] + ARM_A32A * sin(x[0] - x[1]) * x[2]);
  o[2] = x[2] * 0.0 + l[0] - l[2] * ARM_AS +
         l[3] * (0.2e1 * ARM_A32 * x[2] * sin(x[0] - x[1]) + ARM_A32A * cos(x[0] - x[1]) + ARM_C22);
  o[3] = x[3] * 0.0 + l[1] - l[3] * ARM_C22;
}
/* model.hpp:58-62, r = (1, 0.1) model.hpp:83 */
static void arm_hu(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)p;
  o[0] = (1.0 * u[0]) + l[2] * ARM_BS - l[3] * ARM_A32B * cos(x[0] - x[1]) + (double)(u[2] * (2.0 * u[0] - 2.0 * ARM_UC));
  o[1] = -0.5 * 0.1 + (2.0 * u[2] * u[1]);
  o[2] = (u[0] - ARM_UC) * (u[0] - ARM_UC) + u[1] * u[1] - ARM_UR * ARM_UR;
}
/* model.hpp:64-76: 3x3 column major */
static void arm_huu(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)x;
  (void)p;
  (void)l;
  o[0] = 1.0 + 2 * u[2];
  o[1] = 0;
  o[2] = 2 * u[0] - 2 * ARM_UC;
  o[3] = 0;
  o[4] = 2 * u[2];
  o[5] = 2 * u[1];
  o[6] = 2 * u[0] - 2 * ARM_UC;
  o[7] = 2 * u[1];
  o[8] = 0;
}
/* arm_type_inverted_pendulum/simulator.hpp:14-19 */
static void arm_plant(double* o, const double* x, const double* u) { arm_f(o, x, u, NULL); }

#ifdef ORACLE_PORTABLE_TRIG
#undef sin
#undef cos
#endif

/* ---- semiactive_damper: semiactive_damper/model.hpp:4-86 ---- */
#define SAD_A (-1.0)
#define SAD_B (-1.0)
#define SAD_UC ((1.0 + 0.0) / 2.0)
#define SAD_UR ((1.0 - 0.0) / 2.0)

/* model.hpp:36-39 */
static void sad_f(double* o, const double* x, const double* u, const double* p) {
  (void)p;
  o[0] = x[1];
  o[1] = SAD_A * x[0] + SAD_B * u[0] * x[1];
}
/* model.hpp:41-44, sf = (1,10) model.hpp:74 */
static void sad_phix(double* o, const double* x, const double* p) {
  (void)p;
  o[0] = x[0] * 1.0;
  o[1] = x[1] * 10.0;
}
/* model.hpp:46-49, q = (1,10) model.hpp:75 */
static void sad_hx(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)p;
  o[0] = x[0] * 1.0 + SAD_A * l[1];
  o[1] = x[1] * 10.0 + l[0] + SAD_B * u[0] * l[1];
}
/* model.hpp:51-55, r = (1, 0.01) model.hpp:76 */
static void sad_hu(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)p;
  o[0] = 1.0 * u[0] + SAD_B * x[1] * l[1] + 2 * u[2] * (u[0] - SAD_UC);
  o[1] = -0.01 + 2 * u[1] * u[2];
  o[2] = (u[0] - SAD_UC) * (u[0] - SAD_UC) + u[1] * u[1] - SAD_UR * SAD_UR;
}
/* model.hpp:57-69 */
static void sad_huu(double* o, const double* x, const double* u, const double* p, const double* l) {
  (void)x;
  (void)p;
  (void)l;
  o[0] = 1.0 + 2 * u[2];
  o[1] = 0;
  o[2] = 2 * (u[0] - SAD_UC);
  o[3] = 0;
  o[4] = 2 * u[2];
  o[5] = 2 * u[1];
  o[6] = 2 * (u[0] - SAD_UC);
  o[7] = 2 * u[1];
  o[8] = 0;
}
/* semiactive_damper/simulator.hpp:13-16 */
static void sad_plant(double* o, const double* x, const double* u) { sad_f(o, x, u, NULL); }

/* dims / solver parameters: <example>/model.hpp:7-34 of each example; plant dt: simulator.hpp:7 */
static const problem_t PROBLEMS[3] = {
    {4, 6, 2, 50, 5, 2, 0.001, 0.002, 1000.0, 1.0, 0.5, 1e-6, 0.001, msd_f, msd_phix, msd_hx, msd_hu, msd_huu, msd_plant},
    {4, 3, 2, 25, 5, 1, 0.001, 0.002, 1000.0, 0.5, 0.5, 1e-6, 0.001, arm_f, arm_phix, arm_hx, arm_hu, arm_huu, arm_plant},
    {2, 3, 0, 50, 5, 1, 0.001, 0.002, 1000.0, 1.0, 0.5, 1e-6, 0.001, sad_f, sad_phix, sad_hx, sad_hu, sad_huu, sad_plant},
};

static const problem_t* problem_of(int model) { return (model >= 0 && model < 3) ? &PROBLEMS[model] : NULL; }

/* ------------------------------------------------------------------------- */
/* Controller object                                                         */
/* ------------------------------------------------------------------------- */

typedef struct {
  const problem_t* pb;
  int model, L;
  double t;                  /* cgmres.hpp:195 */
  double *U, *dUdt;          /* cgmres.hpp:196-197; dUdt zero-initialised = the de-facto contract */
  double *x_dxh, *ptau, *F1; /* cgmres.hpp:199-202 (F1 == F_dxh_h) */
  /* GMRES workspace, gmres.hpp:11-15 */
  double *V;   /* (k_max+1) columns of length L, column k at L*k */
  double *H;   /* (k_max+1)^2, H[i][j] at (k_max+1)*j + i          */
  double *rho; /* k_max+1 */
  double *g;   /* 3*k_max reflectors */
  double *tmp; /* L */
  /* F_func scratch, cgmres.hpp:116-117 */
  double *xtau, *ltau;
  double *b, *Ubuf;
  int status;
} ctl_t;

/* horizon ramp: cgmres.hpp:32-34 */
static double dtau_of(const problem_t* pb, double t) { return pb->Tf * (1 - exp(-pb->alpha * t)) / (double)pb->dv; }

/* strictly sequential sums: matrix.hpp:140-159 */
static double seq_dot(const double* a, const double* b, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static double seq_norm(const double* a, int n) { return sqrt(seq_dot(a, a, n)); }

/* F(U,x,t): cgmres.hpp:113-162 */
static void eval_F(ctl_t* c, double* out, const double* U, const double* x, double t) {
  const problem_t* pb = c->pb;
  const int nx = pb->dim_x, nu = pb->dim_u, np = pb->dim_p, dv = pb->dv;
  const double dtau = dtau_of(pb, t); /* :127 */
  double* xt = c->xtau;
  double* lt = c->ltau;

  /* forward Euler rollout, :132-140: xtau[i+1] = f*dtau + xtau[i] (mul then add) */
  for (int j = 0; j < nx; j++) xt[j] = x[j];
  for (int i = 0; i < dv; i++) {
    double* nxt = xt + nx * (i + 1);
    const double* cur = xt + nx * i;
    pb->f(nxt, cur, U + nu * i, c->ptau + np * i);
    for (int j = 0; j < nx; j++) nxt[j] = nxt[j] * dtau;
    for (int j = 0; j < nx; j++) nxt[j] = nxt[j] + cur[j];
  }
  /* costate: terminal :145, backward sweep :146-153 */
  pb->phix(lt + nx * dv, xt + nx * dv, c->ptau + np * dv);
  for (int i = dv - 1; i >= 0; i--) {
    double* li = lt + nx * i;
    const double* ln = lt + nx * (i + 1);
    pb->hx(li, xt + nx * i, U + nu * i, c->ptau + np * i, ln);
    for (int j = 0; j < nx; j++) li[j] = li[j] * dtau;
    for (int j = 0; j < nx; j++) li[j] = li[j] + ln[j];
  }
  /* stage-wise dH/du, :156-161 */
  for (int i = 0; i < dv; i++) pb->hu(out + nu * i, xt + nx * i, U + nu * i, c->ptau + np * i, lt + nx * (i + 1));
}

/* forward-difference Jacobian-vector product: cgmres.hpp:164-175; div() multiplies by 1.0/h, matrix.hpp:122-128 */
static void apply_A(ctl_t* c, double* out, const double* v) {
  const problem_t* pb = c->pb;
  const int L = c->L;
  const double inv_h = 1.0 / pb->h;
  for (int i = 0; i < L; i++) c->Ubuf[i] = v[i] * pb->h;
  for (int i = 0; i < L; i++) c->Ubuf[i] = c->Ubuf[i] + c->U[i];
  eval_F(c, out, c->Ubuf, c->x_dxh, c->t + pb->h);
  for (int i = 0; i < L; i++) out[i] = out[i] - c->F1[i];
  for (int i = 0; i < L; i++) out[i] = out[i] * inv_h;
}

/* warm-started GMRES(k_max): gmres.hpp:28-112.  Returns exit code | ncol<<8. */
static int solve_gmres(ctl_t* c, double* x, const double* b) {
  const int L = c->L, km = c->pb->k_max, ld = km + 1;
  const double tol = c->pb->tol;
  double *V = c->V, *H = c->H, *rho = c->rho, *g = c->g;
  int k;
  int code = ORACLE_EXIT_FULL;

  apply_A(c, V, x); /* :33 */
  for (int i = 0; i < L; i++) V[i] = b[i] - V[i]; /* :34 */
  rho[0] = seq_norm(V, L);                        /* :37 */
  if (rho[0] < tol) return ORACLE_EXIT_RHO0;      /* :39-41 */
  {
    const double inv = 1.0 / rho[0]; /* :44 */
    for (int i = 0; i < L; i++) V[i] = V[i] * inv;
  }
  for (k = 0; k < km; k++) {
    double* w = V + L * (k + 1);
    apply_A(c, w, V + L * k); /* :48 */
    /* modified Gram-Schmidt :52-58: tmp = v_i*h ; w -= tmp (two roundings) */
    for (int i = 0; i <= k; i++) {
      const double* vi = V + L * i;
      const double hik = seq_dot(vi, w, L);
      H[ld * k + i] = hik;
      for (int j = 0; j < L; j++) c->tmp[j] = vi[j] * hik;
      for (int j = 0; j < L; j++) w[j] = w[j] - c->tmp[j];
    }
    const double hn = seq_norm(w, L); /* :59-60 */
    H[ld * k + k + 1] = hn;
    if (fabs(hn) < DBL_EPSILON) return ORACLE_EXIT_BREAKDOWN | (k << 8); /* :63-65 (reference prints "Breakdown") */
    {
      const double inv = 1.0 / hn; /* :67 */
      for (int j = 0; j < L; j++) w[j] = w[j] * inv;
    }
    /* apply stored 2x2 Householder reflectors to column k, :71-77 */
    double* hc = H + ld * k;
    for (int i = 0; i < k; i++) {
      const double* gi = g + 3 * i;
      const double s = (gi[0] * hc[i] + gi[1] * hc[i + 1]) * gi[2];
      hc[i] = hc[i] - s * gi[0];
      hc[i + 1] = hc[i + 1] - s * gi[1];
    }
    /* new reflector from (H[k][k], H[k+1][k]), :78-85; sign(0)=+1 matrix.hpp:162 */
    {
      double* gk = g + 3 * k;
      const double sg = (hc[k] < 0.0) ? -1.0 : 1.0;
      const double s = -sg * seq_norm(hc + k, 2);
      gk[0] = hc[k] - s;
      gk[1] = hc[k + 1];
      gk[2] = 2.0 / seq_dot(gk, gk, 2);
      hc[k] = s;
      hc[k + 1] = 0.0;
      /* residual update :88-90 */
      const double r = gk[0] * rho[k] * gk[2];
      rho[k] = rho[k] - r * gk[0];
      rho[k + 1] = -r * gk[1];
    }
    if (fabs(rho[k + 1]) < tol) { /* :93-95: break WITHOUT k++ -> one column dropped */
      code = ORACLE_EXIT_CONVERGED;
      break;
    }
  }
  /* back substitution on the k x k triangle, :100-107 */
  for (int i = k - 1; i >= 0; i--) {
    for (int j = k - 1; j > i; j--) rho[i] -= H[ld * j + i] * rho[j];
    rho[i] /= H[ld * i + i];
  }
  /* x += V[:,0:k]*y staged in column k_max, :110-111 with matrix.hpp:82-91 accumulation order */
  {
    double* s = V + L * km;
    for (int i = 0; i < L; i++) s[i] = 0.0;
    for (int j = 0; j < k; j++)
      for (int i = 0; i < L; i++) s[i] += V[L * j + i] * rho[j];
    for (int i = 0; i < L; i++) x[i] = x[i] + s[i];
  }
  return code | (k << 8);
}

/* Gaussian elimination with partial pivoting, column major, in place: matrix.hpp:166-224 */
static void solve_dense(double* vec, double* mat, int n) {
  for (int k = 0; k < n - 1; k++) {
    int piv = k;
    double best = fabs(mat[n * k + k]);
    for (int i = k + 1; i < n; i++) {
      if (best < fabs(mat[n * k + i])) { /* strict <: first maximum wins */
        best = fabs(mat[n * k + i]);
        piv = i;
      }
    }
    if (piv != k) {
      double s = vec[k];
      vec[k] = vec[piv];
      vec[piv] = s;
      for (int j = k; j < n; j++) {
        s = mat[n * j + k];
        mat[n * j + k] = mat[n * j + piv];
        mat[n * j + piv] = s;
      }
    }
    const double r = 1.0 / mat[n * k + k];
    for (int i = k + 1; i < n; i++) {
      mat[n * k + i] = mat[n * k + i] * r;
      for (int j = k + 1; j < n; j++) mat[n * j + i] -= mat[n * k + i] * mat[n * j + k];
      vec[i] -= mat[n * k + i] * vec[k];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    for (int j = n - 1; j > i; j--) vec[i] -= mat[n * j + i] * vec[j];
    vec[i] /= mat[n * i + i];
  }
}

/* ------------------------------------------------------------------------- */
/* exported object API                                                       */
/* ------------------------------------------------------------------------- */

int ORACLE_FN(model_dims)(int model, int* dims) {
  const problem_t* pb = problem_of(model);
  if (!pb) return -1;
  dims[0] = pb->dim_x;
  dims[1] = pb->dim_u;
  dims[2] = pb->dim_p;
  dims[3] = pb->dv;
  dims[4] = pb->k_max;
  dims[5] = pb->n_ctrl;
  return 0;
}

int ORACLE_FN(model_params)(int model, double* par) {
  const problem_t* pb = problem_of(model);
  if (!pb) return -1;
  par[0] = pb->dt;
  par[1] = pb->h;
  par[2] = pb->zeta;
  par[3] = pb->Tf;
  par[4] = pb->alpha;
  par[5] = pb->tol;
  return 0;
}

static double* zalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

void* ORACLE_FN(create)(int model) {
  const problem_t* pb = problem_of(model);
  if (!pb) return NULL;
  ctl_t* c = (ctl_t*)calloc(1, sizeof(ctl_t));
  c->pb = pb;
  c->model = model;
  c->L = pb->dim_u * pb->dv;
  const int L = c->L, km = pb->k_max;
  c->t = 0.0; /* cgmres.hpp:12 */
  c->U = zalloc(L);
  c->dUdt = zalloc(L);
  c->x_dxh = zalloc(pb->dim_x);
  c->ptau = zalloc((size_t)pb->dim_p * (pb->dv + 1));
  c->F1 = zalloc(L);
  c->V = zalloc((size_t)L * (km + 1));
  c->H = zalloc((size_t)(km + 1) * (km + 1));
  c->rho = zalloc(km + 1);
  c->g = zalloc(3 * km);
  c->tmp = zalloc(L);
  c->xtau = zalloc((size_t)pb->dim_x * (pb->dv + 1));
  c->ltau = zalloc((size_t)pb->dim_x * (pb->dv + 1));
  c->b = zalloc(L);
  c->Ubuf = zalloc(L);
  return c;
}

void ORACLE_FN(destroy)(void* h) {
  ctl_t* c = (ctl_t*)h;
  if (!c) return;
  free(c->U);
  free(c->dUdt);
  free(c->x_dxh);
  free(c->ptau);
  free(c->F1);
  free(c->V);
  free(c->H);
  free(c->rho);
  free(c->g);
  free(c->tmp);
  free(c->xtau);
  free(c->ltau);
  free(c->b);
  free(c->Ubuf);
  free(c);
}

/* cgmres.hpp:36-39 */
void ORACLE_FN(set_ptau)(void* h, const double* ptau_buf) {
  ctl_t* c = (ctl_t*)h;
  memcpy(c->ptau, ptau_buf, sizeof(double) * c->pb->dim_p * (c->pb->dv + 1));
}
/* cgmres.hpp:41-49 */
void ORACLE_FN(set_ptau_repeat)(void* h, const double* p_buf) {
  ctl_t* c = (ctl_t*)h;
  for (int i = 0; i <= c->pb->dv; i++)
    for (int j = 0; j < c->pb->dim_p; j++) c->ptau[c->pb->dim_p * i + j] = p_buf[j];
}
/* cgmres.hpp:51-59 */
void ORACLE_FN(init_u0)(void* h, const double* u0) {
  ctl_t* c = (ctl_t*)h;
  for (int i = 0; i < c->pb->dv; i++)
    for (int j = 0; j < c->pb->dim_u; j++) c->U[c->pb->dim_u * i + j] = u0[j];
}
/* cgmres.hpp:61-76 (mutates u0) */
void ORACLE_FN(init_u0_newton)(void* h, double* u0, const double* x0, const double* p0, int n_loop) {
  ctl_t* c = (ctl_t*)h;
  const problem_t* pb = c->pb;
  double lmd0[8], vec[8], mat[64];
  pb->phix(lmd0, x0, p0);
  for (int it = 0; it < n_loop; it++) {
    pb->hu(vec, x0, u0, p0, lmd0);
    pb->huu(mat, x0, u0, p0, lmd0);
    solve_dense(vec, mat, pb->dim_u);
    for (int j = 0; j < pb->dim_u; j++) u0[j] = u0[j] - vec[j];
  }
  ORACLE_FN(init_u0)(h, u0);
}

/* one control update: cgmres.hpp:78-110 */
void ORACLE_FN(control)(void* h, double* u, const double* x) {
  ctl_t* c = (ctl_t*)h;
  const problem_t* pb = c->pb;
  const int L = c->L, nx = pb->dim_x;
  /* x + dxdt*h, :83-85 */
  pb->f(c->x_dxh, x, c->U, c->ptau);
  for (int j = 0; j < nx; j++) c->x_dxh[j] = c->x_dxh[j] * pb->h;
  for (int j = 0; j < nx; j++) c->x_dxh[j] = c->x_dxh[j] + x[j];
  eval_F(c, c->F1, c->U, c->x_dxh, c->t + pb->h); /* :88 */
  eval_F(c, c->b, c->U, x, c->t);                 /* :91 */
  {                                               /* :94-96 */
    const double c1 = (1 - pb->zeta * pb->h);
    const double inv_h = 1.0 / pb->h;
    for (int i = 0; i < L; i++) c->b[i] = c->b[i] * c1;
    for (int i = 0; i < L; i++) c->b[i] = c->b[i] - c->F1[i];
    for (int i = 0; i < L; i++) c->b[i] = c->b[i] * inv_h;
  }
  c->status = solve_gmres(c, c->dUdt, c->b); /* :99 */
  for (int i = 0; i < L; i++) c->tmp[i] = c->dUdt[i] * pb->dt; /* :102 */
  for (int i = 0; i < L; i++) c->U[i] = c->U[i] + c->tmp[i];  /* :103 */
  c->t = c->t + pb->dt;                                        /* :107 */
  for (int j = 0; j < pb->dim_u; j++) u[j] = c->U[j];          /* :109 */
}

double ORACLE_FN(get_dtau)(void* h, double t) { return dtau_of(((ctl_t*)h)->pb, t); }

void ORACLE_FN(get_state)(void* h, double* t, double* U, double* dUdt) {
  ctl_t* c = (ctl_t*)h;
  if (t) *t = c->t;
  if (U) memcpy(U, c->U, sizeof(double) * c->L);
  if (dUdt) memcpy(dUdt, c->dUdt, sizeof(double) * c->L);
}
void ORACLE_FN(set_state)(void* h, const double* t, const double* U, const double* dUdt) {
  ctl_t* c = (ctl_t*)h;
  if (t) c->t = *t;
  if (U) memcpy(c->U, U, sizeof(double) * c->L);
  if (dUdt) memcpy(c->dUdt, dUdt, sizeof(double) * c->L);
}
int ORACLE_FN(last_status)(void* h) { return ((ctl_t*)h)->status; }

/* Euler plant step with the *returned* u: <example>/main.cpp:74-76 */
void ORACLE_FN(plant_step)(int model, double* x, const double* u) {
  const problem_t* pb = problem_of(model);
  double d[8];
  pb->plant(d, x, u);
  for (int j = 0; j < pb->dim_x; j++) d[j] = d[j] * pb->plant_dt;
  for (int j = 0; j < pb->dim_x; j++) x[j] = x[j] + d[j];
}

/* ------------------------------------------------------------------------- */
/* batch closed loop on host threads                                         */
/* ------------------------------------------------------------------------- */

typedef struct {
  int model, p_full, newton_iters, n_steps, rec_stride, tid, n_threads;
  int64_t n;
  const double *x0, *p, *u0;
  double *x_traj, *u_traj, *x_fin, *u_fin, *U_fin, *dUdt_fin, *ctl_seconds;
  int32_t* exit_hist;
  double loop_seconds; /* wall time this thread spent inside its closed-loop step loops */
} job_t;
static double g_last_loop_seconds = 0.0;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker(void* arg) {
  job_t* jb = (job_t*)arg;
  const problem_t* pb = problem_of(jb->model);
  const int nx = pb->dim_x, nu = pb->dim_u, np = pb->dim_p, L = pb->dim_u * pb->dv;
  const int plen = jb->p_full ? np * (pb->dv + 1) : np;
  for (int64_t n = jb->tid; n < jb->n; n += jb->n_threads) {
    double x[8], u[8], p0[8] = {0};
    ctl_t* c = (ctl_t*)ORACLE_FN(create)(jb->model);
    const double* pn = jb->p ? jb->p + (size_t)plen * n : NULL;
    if (np > 0) {
      if (jb->p_full)
        ORACLE_FN(set_ptau)(c, pn);
      else
        ORACLE_FN(set_ptau_repeat)(c, pn);
      for (int j = 0; j < np; j++) p0[j] = pn[j];
    }
    for (int j = 0; j < nx; j++) x[j] = jb->x0[(size_t)nx * n + j];
    for (int j = 0; j < nu; j++) u[j] = jb->u0[(size_t)nu * n + j];
    ORACLE_FN(init_u0)(c, u);
    ORACLE_FN(init_u0_newton)(c, u, x, p0, jb->newton_iters);
    double acc = 0.0;
    const double loop0 = now_s();
    for (int s = 0; s < jb->n_steps; s++) {
      const double t0 = now_s();
      ORACLE_FN(control)(c, u, x);
      acc += now_s() - t0;
      ORACLE_FN(plant_step)(jb->model, x, u);
      if (jb->exit_hist) jb->exit_hist[4 * n + (c->status & 0xff)]++;
      if (jb->rec_stride > 0 && (s + 1) % jb->rec_stride == 0) {
        const size_t r = (size_t)((s + 1) / jb->rec_stride - 1);
        if (jb->x_traj) memcpy(jb->x_traj + (r * jb->n + n) * nx, x, sizeof(double) * nx);
        if (jb->u_traj) memcpy(jb->u_traj + (r * jb->n + n) * nu, u, sizeof(double) * nu);
      }
    }
    jb->loop_seconds += now_s() - loop0;
    if (jb->x_fin) memcpy(jb->x_fin + (size_t)nx * n, x, sizeof(double) * nx);
    if (jb->u_fin) memcpy(jb->u_fin + (size_t)nu * n, u, sizeof(double) * nu);
    if (jb->U_fin) memcpy(jb->U_fin + (size_t)L * n, c->U, sizeof(double) * L);
    if (jb->dUdt_fin) memcpy(jb->dUdt_fin + (size_t)L * n, c->dUdt, sizeof(double) * L);
    if (jb->ctl_seconds) jb->ctl_seconds[n] = acc;
    ORACLE_FN(destroy)(c);
  }
  return NULL;
}

int ORACLE_FN(run_closed_loop)(int model, int64_t n, const double* x0, const double* p, int p_full,
                               const double* u0, int newton_iters, int n_steps, int rec_stride,
                               double* x_traj, double* u_traj, double* x_fin, double* u_fin,
                               double* U_fin, double* dUdt_fin, int32_t* exit_hist,
                               double* ctl_seconds, int n_threads) {
  const problem_t* pb = problem_of(model);
  if (!pb || n < 0 || !x0 || !u0 || (pb->dim_p > 0 && !p) || n_steps < 0) return -1;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 1024) n_threads = 1024;
  if (exit_hist) memset(exit_hist, 0, sizeof(int32_t) * 4 * (size_t)n);
  job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
  for (int t = 0; t < n_threads; t++) {
    job_t jb = {model, p_full, newton_iters, n_steps, rec_stride, t, n_threads, n, x0, p, u0,
                x_traj, u_traj, x_fin, u_fin, U_fin, dUdt_fin, ctl_seconds, exit_hist, 0.0};
    jobs[t] = jb;
    if (n_threads == 1)
      worker(&jobs[t]);
    else
      pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  if (n_threads > 1)
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  g_last_loop_seconds = 0.0;
  for (int t = 0; t < n_threads; t++)
    if (jobs[t].loop_seconds > g_last_loop_seconds) g_last_loop_seconds = jobs[t].loop_seconds;
  free(jobs);
  free(th);
  return 0;
}

/* max over the worker threads of the wall time spent inside the closed-loop step loops of the last
 * run_closed_loop (controller construction, init_u0_newton and thread start-up excluded) */
double ORACLE_FN(last_loop_seconds)(void) { return g_last_loop_seconds; }
