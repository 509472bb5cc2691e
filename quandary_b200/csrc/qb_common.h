/*
 * qb_common.h -- types shared by the host side of the engine (engine.cu) and the kernel translation units.
 */
#pragma once
#include "../../include/quandary_b200.h"

#include <cuda_runtime.h>
#include <math.h>

constexpr int kMaxOsc = QB_MAX_OSC;
constexpr int kMaxPairs = QB_MAX_PAIRS;
constexpr int kRestart = 30;        /* PETSc's default GMRES restart length */
constexpr int kRedStride = kRestart + 2;
constexpr int kLdh = kRestart + 2;
constexpr int kMaxUpd = 96;         /* gradient updater threads per step: sum_k ncarrier_k * 3 */
constexpr int kMaxThreads = 256;    /* small-N kernels: one thread per element, dim <= 256 */

/* Static system description as the kernels see it. */
struct DevSys {
  int Q, N, dim, lind, npairs;
  int nlev[kMaxOsc], stride[kMaxOsc];
  double detuning[kMaxOsc], selfkerr[kMaxOsc], decay[kMaxOsc], dephase[kMaxOsc];
  double crosskerr[kMaxPairs], Jkl[kMaxPairs], eta[kMaxPairs];
  int ntime;            /* number of implicit-midpoint (sub-)steps of a sweep = time steps x nstage */
  double dt;            /* total_time / (number of time steps), src/timestepper.cpp:35 */
  int nstage;           /* 1 (IMR), 3 (IMR4), 15 (IMR8): CompositionalImplMidpoint, src/timestepper.cpp:731-826 */
  int spred_bias;       /* large-state kernels: added to the predicted basis length of the one-reduction solver (QB_SPRED_BIAS, default 0) */
  int ee;               /* 1: explicit Euler (ExplEuler, src/timestepper.cpp:484-519): the control table has ntime+1 rows at t_n = n*dt */
  double gamma[15];
  int linsolve, maxiter;
  double abstol, reltol;
  int nd;               /* design parameters */
  int osc_off[kMaxOsc + 1];
  int ncar_total;
  int car_off[kMaxOsc + 1];
  int tabstride;        /* doubles per step in the control table */
  int nupd;             /* gradient updater threads */
  int flags;
  /* sqrt(i*j) for level numbers i, j < 8 at [i*8 + j]: systems whose oscillators have at most 7 levels read their ladder-operator
   * factors from the kernel-parameter constant bank instead of a shared-memory table (the large-state kernels are bound by the
   * shared-memory pipe) */
  double sq8[64];
};

/* State-dependent penalty integral (TimeStepper::penaltyIntegral[_diff], src/timestepper.cpp:256-339) in the generic form
 *   penalty_ic = sum_{n=1..ntime} wdt[n] * (c0_ic + <a_ic, x_n> + <d, x_n^2>) + gw[n] * <g, x_n^2>
 * (weighted objective integral + guard-level occupation).  The adjoint sweep adds its derivative to the adjoint state at
 * every step; the forward value is reduced from the stored trajectory by penalty_reduce_kernel. */
struct PenArgs {
  const double2* a;      /* [ninit][dimp] */
  const double2* d;      /* [dimp] */
  const double2* g;      /* [dimp], null without leakage term */
  const double* wdt;     /* [ntime+1] */
  const double* jbar;    /* [ninit] */
  const double* gw;      /* [ntime+1]: 1/(number of time steps) at full time steps, 0 at compositional sub-steps; null without leakage term */
  /* overlap mode (Schroedinger + Jtrace): J(x) = 1 - (Re<a,x>/purity)^2 - (Im<a,x>)^2.  ov[ic][n] = (-2 Re<a,x_n>/purity^2,
   * -2 Im<a,x_n>) from overlap_table_kernel; the source term of the adjoint sweep is then the complex product ov * a */
  const double2* ov;     /* [ninit][ntime+1] or null */
  int on;
  /* second-derivative ("dpdm") penalty of the populations, Schroedinger only (TimeStepper::penaltyDpDm[_diff],
   * src/timestepper.cpp:342-442): sum_n sum_i (|x_{n+1,i}|^2 - 2|x_{n,i}|^2 + |x_{n-1,i}|^2)^2 / dt^4, from the stored trajectory */
  const double2* dp_traj;   /* [ninit][ntime+1][dimp] */
  const double* dp_jbar;    /* [ninit]: adjoint seed (already divided by the number of time steps) */
  double dp_dtinv;          /* 1/dt^4 */
  int dp_on;
};

struct SweepArgs {
  const double* tab;
  const int* gidx;
  const double* gval;
  const double* gcs;
  const int* upd;        /* [nupd][4]: oscillator, global carrier index, carrier index inside oscillator, j */
  double2* state;        /* [ninit][dimp]: in: rho0 (fwd) / rhoT_bar (adj) */
  double2* state_out;    /* out: rhoT (fwd) / rho0_bar (adj) */
  double2* final_state;  /* adj, Schroedinger: rho(T) of the forward sweep [ninit][dimp] */
  double2* traj;         /* [ninit][ntime+1][dimp] or null */
  double2* stage_traj;   /* [ninit][ntime][dimp] stored stages x_n + dt/2 k (QB_FLAG_STORE_STAGE, Lindblad) or null */
  double* grad_ic;       /* [ninit][nd] */
  unsigned long long* counters; /* [0] solver iterations, [1] rhs evals */
  int ninit, mcap, store;
  PenArgs pen;
};

typedef void (*SweepFn)(const DevSys, const SweepArgs);
typedef void (*ApplyFn)(const DevSys, const double*, const double2*, double2*, int);

struct KernelSet {
  SweepFn fwd, adj;
  ApplyFn app;
  int dimp;              /* threads per CTA = padded state dimension */
  unsigned smem_fixed;   /* bytes of shared memory before the Krylov basis */
  unsigned vec_bytes;    /* bytes of one state vector in shared memory */
};

template <int Q>
void qb_pick(bool all2, bool hasj, bool lind, int dim, KernelSet* ks);

/* ---- qubit-system kernels (qb_qubit.cuh) ---- */
namespace qq { struct QArgs; }
typedef void (*QSweepFn)(const DevSys, const qq::QArgs);
struct QKernelSet {
  QSweepFn fwd[2], adj[2], adj_stored[2];   /* [0] GMRES, [1] Neumann; adj_stored: stored stages (QB_FLAG_STORE_STAGE) */
  ApplyFn app;
  int threads, nw, dim, epT;
  unsigned vec_bytes;
};
void qq_pick(int Q, bool lind, int epT, QKernelSet* ks);
size_t qq_smem(const QKernelSet& ks, bool adj, bool with_stage, unsigned recbytes, int nd);

/* ---- large-state kernels (qb_big.cuh), Lindblad, 256 < dim <= 4096 ---- */
namespace qbig { struct BArgs; }
typedef void (*BSweepFn)(const DevSys, const qbig::BArgs);
typedef void (*BApplyFn)(const DevSys, const double*, const double2*, double2*, int, int, int);
struct BKernelSet {
  BSweepFn fwd, adj;
  BApplyFn app;
  int threads, dimp, nred;
};
void qbig_pick_q1(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_q2(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_q3(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_q4(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_q5(bool hasj, int dim, BKernelSet* ks);
size_t qbig_smem(const BKernelSet& ks, int mcap, int L, bool adj, int nd);
