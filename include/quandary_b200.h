/*
 * quandary_b200.h -- C ABI of the B200-native forward/adjoint Lindblad time-stepping engine.
 *
 * This is the drop-in boundary for the hot path of LLNL/quandary (reference paths below are relative to the
 * reference repository root).  The reference has no FFI of its own: its boundary is the C++ class surface
 *   MasterEq      include/mastereq.hpp:91-263    (assemble_RHS, getRHS + MatMult, compute_dRHS_dParams, setControlAmplitudes)
 *   TimeStepper   include/timestepper.hpp:48-197 (solveODE, solveAdjointODE, getState, evolveFWD/evolveBWD)
 *   OptimProblem  include/optimproblem.hpp:123-256 (evalF, evalGradF)
 * The host-side C++ classes of the same names in quandary_b200/host/ keep that surface and call the functions below.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every function returns 0 on success or a QB_ERR_* code, and
 *     qb_last_error() returns a human-readable message for the calling thread.
 *   - all pointers are HOST memory unless the name ends in _dev.  Caller owns inputs/outputs, engine owns device memory.
 *   - state vectors use the reference's layout (src/mastereq.cpp:1322-1323, src/util.cpp:150-152):
 *       x[0..dim) = Re, x[dim..2*dim) = Im, vec(rho) column-major, dim = N^2 (Lindblad) or N (Schroedinger),
 *       oscillator 0 most significant in the composite index.
 *   - one engine per CUDA device and per process; an engine is not thread-safe.
 *   - frequencies in the descriptor are ANGULAR (already multiplied by 2*pi as the reference does on load,
 *     src/oscillator.cpp:15-21, src/mastereq.cpp:29-37); decay/dephase TIMES are passed as in the .cfg
 *     (gamma = 1/T, src/mastereq.cpp:2070-2077).
 */
#ifndef QUANDARY_B200_H
#define QUANDARY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_MAX_OSC 5   /* the reference's matrix-free path stops at 5 oscillators (src/mastereq.cpp:722-737) */
#define QB_MAX_PAIRS 10

enum { QB_OK = 0, QB_ERR_INVALID = 1, QB_ERR_UNSUPPORTED = 2, QB_ERR_CUDA = 3, QB_ERR_STATE = 4, QB_ERR_NOMEM = 5 };

/* LindbladType, include/defs.hpp */
enum { QB_LINDBLAD_NONE = 0, QB_LINDBLAD_DECAY = 1, QB_LINDBLAD_DEPHASE = 2, QB_LINDBLAD_BOTH = 3 };
/* LinearSolverType, include/defs.hpp */
enum { QB_LINSOLVE_GMRES = 0, QB_LINSOLVE_NEUMANN = 1 };
/* ControlType subset with an adjoint in the reference (src/oscillator.cpp:350-355) */
enum { QB_CTRL_BSPLINE2ND = 0, QB_CTRL_BSPLINE0 = 1 };
/* TimeStepper type (src/main.cpp:357-366) */
enum { QB_STEPPER_IMR = 0, QB_STEPPER_EE = 1, QB_STEPPER_IMR4 = 4, QB_STEPPER_IMR8 = 8 };
/* QB_STEPPER_EE = ExplEuler (src/timestepper.cpp:484-519), state dimension <= 256 only (general small-N kernels) */

/* One control segment of one oscillator: ControlBasis (include/controlbasis.hpp, src/controlbasis.cpp:20-96,219-245). */
typedef struct qb_segment {
  int32_t type;       /* QB_CTRL_* */
  int32_t nsplines;   /* number of basis functions per carrier; the segment owns 2*nsplines*ncarrier parameters */
  double tstart;      /* segment is active for tstart <= t <= tstop; first match wins (src/oscillator.cpp:296-322) */
  double tstop;
  int32_t skip;       /* offset of this segment inside the oscillator's parameter block (ControlBasis::setSkip) */
  int32_t enforce_bc; /* control_enforceBC: skip the first/last two splines (src/controlbasis.cpp:54-56) */
} qb_segment;

/* Everything MatShellCtx (include/mastereq.hpp:20-42), Oscillator and ImplMidpoint hold that the hot path reads. */
typedef struct qb_system_desc {
  int32_t nosc;
  int32_t nlevels[QB_MAX_OSC];
  int32_t lindblad;                 /* QB_LINDBLAD_* */
  double detuning[QB_MAX_OSC];      /* 2*pi*(transfreq - rotfreq) */
  double selfkerr[QB_MAX_OSC];      /* 2*pi*xi_k */
  double crosskerr[QB_MAX_PAIRS];   /* 2*pi*xi_kl, pairs in order (0,1),(0,2),...,(1,2),... */
  double Jkl[QB_MAX_PAIRS];         /* 2*pi*J_kl */
  double eta[QB_MAX_PAIRS];         /* 2*pi*(rotfreq_k - rotfreq_l) */
  double decay_time[QB_MAX_OSC];    /* T1 (0 = none) */
  double dephase_time[QB_MAX_OSC];  /* T2 (0 = none) */

  int32_t ntime;                    /* number of time steps */
  double total_time;                /* = ntime*dt of the .cfg (src/main.cpp:189); stepper dt = total_time/ntime */
  int32_t stepper;                  /* QB_STEPPER_* */
  int32_t linsolve_type;            /* QB_LINSOLVE_* */
  int32_t linsolve_maxiter;
  double linsolve_abstol;           /* 1e-10 in the reference (src/timestepper.cpp:535-536) */
  double linsolve_reltol;           /* 1e-20 */

  /* controls: flat arrays, oscillator k owns [off[k], off[k+1]) */
  int32_t seg_off[QB_MAX_OSC + 1];
  const qb_segment* segments;
  int32_t carrier_off[QB_MAX_OSC + 1];
  const double* carrier_freq;       /* angular */
  int32_t pipulse_off[QB_MAX_OSC + 1];
  const double* pipulse;            /* triplets (tstart, tstop, amp) per entry (include/oscillator.hpp PiPulse) */

  int32_t device;                   /* CUDA device ordinal */
  int32_t flags;                    /* QB_FLAG_* */
} qb_system_desc;

/* flags */
#define QB_FLAG_RECOMPUTE_STAGE 1  /* adjoint step re-solves for the primal stage exactly as ImplMidpoint::evolveBWD does
                                      (src/timestepper.cpp:644,674-680) instead of using (x_{n-1}+x_n)/2 */

#define QB_FLAG_STORE_STAGE 2      /* the forward sweep also stores the implicit-midpoint stage x_n + dt/2 k of every step
                                      (a second trajectory buffer) and the adjoint sweep reads it instead of re-solving:
                                      the value QB_FLAG_RECOMPUTE_STAGE computes, one linear solve less per adjoint
                                      step; with store_trajectory = 2 the stages REPLACE the state trajectory (same
                                      memory as the reference's storeFWD).  Lindblad + qubit-system / large-state kernels;
                                      otherwise the engine falls back to QB_FLAG_RECOMPUTE_STAGE. */

#define QB_FLAG_ARNOLDI 4          /* qubit-system kernels: run every GMRES solve with the step-by-step Arnoldi process (classical
                                      Gram-Schmidt, Givens residual recurrence: KSPGMRES as the reference drives it) instead of
                                      the one-reduction formulation of the same iterates; a verification switch, slower */

typedef struct qb_engine qb_engine;

/* Work counters of the last forward/adjoint call, summed over the local initial conditions. */
typedef struct qb_stats {
  int64_t steps_fwd;        /* IMR steps taken by forward sweeps */
  int64_t steps_adj;        /* adjoint IMR steps */
  int64_t rhs_evals;        /* M*x and M^T*x applications + one per dRHSdp pass */
  int64_t linsolve_iters;   /* linear-solver iterations */
  int64_t kernel_launches;  /* CUDA kernels launched by the engine since creation */
  double last_fwd_ms;       /* device time of the last forward sweep kernel (CUDA events on the engine stream) */
  double last_adj_ms;       /* device time of the last adjoint sweep kernel */
  double total_fwd_ms;      /* device time of all forward sweep kernels since creation, and their number */
  double total_adj_ms;
  int64_t n_fwd;
  int64_t n_adj;
} qb_stats;

/* ---- life cycle ------------------------------------------------------------------------------------------- */
/* Replaces the constructors of MasterEq (src/mastereq.cpp:14-155) and ImplMidpoint (src/timestepper.cpp:521-551). */
int qb_engine_create(const qb_system_desc* desc, qb_engine** out);
int qb_engine_destroy(qb_engine* e);
const char* qb_last_error(void);

/* number of design parameters (sum of Oscillator::getNParams), complex state dimension */
int qb_engine_ndesign(const qb_engine* e);
int qb_engine_dim(const qb_engine* e);

/* ---- controls -------------------------------------------------------------------------------------------- */
/* = MasterEq::setControlAmplitudes (src/mastereq.cpp:693-707): oscillators concatenated.  Evaluates the control
 * table p_k(t_{n+1/2}), q_k(t_{n+1/2}) for every step on the device (replaces MasterEq::assemble_RHS,
 * src/mastereq.cpp:657-678, and Oscillator::evalControl, src/oscillator.cpp:281-337). */
int qb_engine_set_params(qb_engine* e, const double* alpha, int ndesign);
/* p_k(t), q_k(t) at arbitrary times, evaluated on the device = Oscillator::evalControl.  out_p/out_q: [nt][nosc]. */
int qb_engine_eval_controls(qb_engine* e, const double* t, int nt, double* out_p, double* out_q);

/* ---- time stepping --------------------------------------------------------------------------------------- */
/* = ninit x TimeStepper::solveODE (src/timestepper.cpp:96-181).  rho0, rhoT: [ninit][2*dim].
 * store_trajectory: 0 keeps nothing; 1 keeps all ntime+1 states of every initial condition on the device (storeFWD, what
 * TimeStepper::store_states holds, src/timestepper.cpp:132) plus, under QB_FLAG_STORE_STAGE, the stages; 2 keeps only what the
 * following adjoint sweep reads (the stages under QB_FLAG_STORE_STAGE on the qubit-system kernels, else the states): half the
 * trajectory memory and HBM writes, but qb_engine_get_state and the penalty terms, which read states, then fail with
 * QB_ERR_STATE. */
int qb_engine_forward(qb_engine* e, const double* rho0, int ninit, int store_trajectory, double* rhoT);
/* = ninit x TimeStepper::solveAdjointODE (src/timestepper.cpp:184-253) + the VecAXPY(G, 1, redgrad) of
 * OptimProblem::evalGradF (src/optimproblem.cpp:442,514).  rhoT_bar: terminal adjoint states [ninit][2*dim]
 * (already weighted); redgrad[ndesign] receives the sum over the local initial conditions (summed in initial-condition
 * order).  Must follow a qb_engine_forward with the same ninit (Lindblad: with store_trajectory). */
int qb_engine_adjoint(qb_engine* e, const double* rhoT_bar, int ninit, double* redgrad, double* rho0_bar /* may be NULL */);
/* = TimeStepper::getState (src/timestepper.cpp:86-94) for initial condition iinit: out[2*dim]. */
int qb_engine_get_state(qb_engine* e, int iinit, int tindex, double* out);
/* = MasterEq::assemble_RHS(t) + MatMult / MatMultTranspose (src/mastereq.cpp:657-678,710-739). x,y: [nbatch][2*dim] */
int qb_engine_rhs_apply(qb_engine* e, double t, int transpose, const double* x, double* y, int nbatch);

/* ---- penalty integral (TimeStepper::penaltyIntegral / penaltyIntegral_diff, src/timestepper.cpp:256-339) ------------- */
/* The reference adds, per initial condition, gamma * sum_n [ w(t_n) * dt * J(x_n) + guard-level occupation(x_n) / ntime ]
 * to the objective (w(t) = exp(-((t-T)/p)^2)/p, J = the final-time objective) and the corresponding source terms to the
 * adjoint sweep.  The engine takes J in the generic form  J(x) = c0 + <a, x> + <d, x.^2>  (covers Jmeasure, Jtrace for
 * Lindblad, Jfrobenius; the caller fills a, c0 per initial condition and the shared d from its OptimTarget):
 *   a     [ninit][2*dim]  per-initial-condition linear part, reference state layout
 *   c0    [ninit]         constant part
 *   d     [2*dim]         diagonal quadratic weights
 *   guard [2*dim] or NULL 1.0 at the diagonal entries of guard levels (real and imaginary part), else 0
 *   wdt   [ntime+1]       w(n*dt)*dt for n = 0..ntime (all zero: no weighted-objective integral)
 *   jbar  [ninit]         adjoint seed of the penalty per initial condition (objective weight * optim_penalty) */
typedef struct qb_penalty_desc {
  const double* a;
  const double* c0;
  const double* d;
  const double* guard;
  const double* wdt;
  const double* jbar;
  int32_t ninit;
  /* Schroedinger solver with Jtrace: the objective is not of the form above but J(x) = 1 - (Re<tau,x>/purity)^2 - (Im<tau,x>)^2
   * (evalJ + finalizeJ, src/optimtarget.cpp:756-760, 864-880).  overlap != 0 selects it: `a` then holds the target states tau_i
   * ([ninit][2*dim]; the unit vector of a pure target), `purity` [ninit] the purities of the initial states (NULL = 1), and c0 / d
   * are ignored.  The guard-level term is unchanged. */
  int32_t overlap;
  const double* purity;
} qb_penalty_desc;
/* Installs (or, with NULL, removes) the penalty for the following sweeps of `ninit` initial conditions. */
int qb_engine_set_penalty(qb_engine* e, const qb_penalty_desc* pen);
/* out[i] = sum_{n=1..ntime} wdt[n]*J_i(x_n) + guard occupation/ntime, from the trajectory stored by the last forward sweep. */
int qb_engine_penalty_integrals(qb_engine* e, int ninit, double* out);

/* Second-derivative penalty of the level populations ("dpdm", TimeStepper::penaltyDpDm / penaltyDpDm_diff,
 * src/timestepper.cpp:342-442; Schroedinger solver and timestepper = IMR only, as in the reference).
 * jbar[ninit]: adjoint seed per initial condition = objective weight * optim_penalty_dpdm / ntime; NULL removes the penalty.
 * qb_engine_dpdm_integrals: out[i] = sum_{n=1..ntime-1} sum_j (|x_{n+1,j}|^2 - 2|x_{n,j}|^2 + |x_{n-1,j}|^2)^2 / dt^4 from the
 * trajectory stored by the last forward sweep (the caller divides by ntime as the reference does, src/timestepper.cpp:160). */
int qb_engine_set_dpdm(qb_engine* e, const double* jbar, int ninit);
int qb_engine_dpdm_integrals(qb_engine* e, int ninit, double* out);

/* ---- final-time objective on the device (OptimTarget::evalJ / evalJ_diff / HilbertSchmidtOverlap[_diff],
 *      src/optimtarget.cpp:343-447, 713-862; the weighted sums of OptimProblem::evalF / evalGradF, src/optimproblem.cpp:263-281,
 *      405-436, 495-512) -------------------------------------------------------------------------------------------------
 * The target states V rho_i(0) V^dagger (gate targets), the purities of rho_i(0) and the weights do not depend on the design
 * vector: the caller computes them once and leaves them on the device.  After a forward sweep the engine then forms, without
 * moving a state to the host,
 *   qb_engine_objective:      out[0..3] = sum_i w_i Re J_i, sum_i w_i Im J_i, fid_scale * sum_i Re HS_i, fid_scale * sum_i Im HS_i
 *                             over the initial conditions [0, ninit) of the last forward sweep (J_i = evalJ(x_i(T)), HS_i =
 *                             HilbertSchmidtOverlap(x_i(T), false)); fixed summation order
 *   qb_engine_objective_bar:  the terminal adjoint states  w_i * (re_bar dReJ_i/dx + im_bar dImJ_i/dx)  (evalJ_diff with the
 *                             seeds finalizeJ_diff returns), written to the buffer the next qb_engine_adjoint_dev starts from.
 *   objective   QB_OBJ_JFROBENIUS | QB_OBJ_JTRACE | QB_OBJ_JMEASURE (Jmeasure: pure targets only, as in the reference)
 *   pure_id     >= 0: pure target state |m><m| with m = pure_id (Hilbert-space index; `target` unused); < 0: per-state targets
 *   target      [ninit][2*dim] target states in the reference layout (gate / file targets), or NULL for pure targets
 *   purity      [ninit] purity of rho_i(0) (Jtrace divides the overlap by it, src/optimtarget.cpp:391-393); NULL = all 1
 *   weight      [ninit] objective weights of this process' initial conditions (already normalised, src/optimproblem.cpp:71-91)
 *   fid_scale   1 / (global number of initial conditions), src/optimproblem.cpp:279 */
#define QB_OBJ_JFROBENIUS 0
#define QB_OBJ_JTRACE 1
#define QB_OBJ_JMEASURE 2
typedef struct qb_objective_desc {
  int32_t objective;
  int32_t pure_id;
  const double* target;
  const double* purity;
  const double* weight;
  double fid_scale;
  int32_t ninit;
} qb_objective_desc;
/* Installs (or, with NULL, removes) the objective for sweeps of `ninit` initial conditions. */
int qb_engine_set_objective(qb_engine* e, const qb_objective_desc* obj);
int qb_engine_objective(qb_engine* e, int ninit, double out[4]);
int qb_engine_objective_bar(qb_engine* e, int ninit, double re_bar, double im_bar);

/* Largest number of initial conditions whose sweep buffers (with store_trajectory: states, and stages under
 * QB_FLAG_STORE_STAGE) fit the free device memory right now; the host layer batches its initial-condition loop with it. */
int qb_engine_max_batch(qb_engine* e, int store_trajectory);

/* ---- device-resident variants (inputs/outputs stay in HBM; used to time the kernels alone) ---------------- */
/* Device buffers hold states in the engine's internal layout; fill them with qb_engine_upload_states. */
/* rho0 and rhoT_bar live in separate device buffers: a forward+adjoint evaluation can be repeated without re-upload. */
int qb_engine_upload_states(qb_engine* e, const double* host_states, int ninit, int which /*0: rho0, 1: rhoT_bar*/);
int qb_engine_forward_dev(qb_engine* e, int ninit, int store_trajectory);
int qb_engine_adjoint_dev(qb_engine* e, int ninit);
int qb_engine_download_states(qb_engine* e, double* host_states, int ninit, int which /*0: rhoT, 1: rho0_bar*/);
int qb_engine_download_grad(qb_engine* e, double* redgrad);
/* device pointer of the reduced gradient buffer [ndesign + 8] (for an in-place NCCL all-reduce by the caller) */
void* qb_engine_grad_dev(qb_engine* e);

/* Enqueue all engine work on the caller's CUDA stream (external != 0; cuda_stream is a cudaStream_t, 0 = the legacy
 * default stream) or back on the engine-owned stream (external == 0).  The reference has no stream notion: this is the
 * hook that lets a host (torch.distributed, NCCL) order its collective after the sweep without a host sync. */
int qb_engine_set_stream(qb_engine* e, void* cuda_stream, int external);
/* Measured FP64 FMA peak of the device in TFLOP/s (the denominator of the FP64 roofline; MEASURED_PEAKS.json has none). */
int qb_measure_fp64_peak(int device, double* tflops);

/* Measured fp64 tensor-pipe peak (mma.sync m8n8k4 f64 = DMMA) in TFLOP/s, next to the DFMA figure above. */
int qb_measure_fp64_mma_peak(int device, double* tflops);
/* A/B hook for the operator formulation at N = 64 (BASELINE config 5 "tensor-core H*rho path"): device time of ONE launch that
 * applies M(t) `reps` times to `nbatch` resident states (uploaded with which = 1), with the stencil kernel (dense = 0) or with
 * the dense DMMA kernel (dense = 1; -i[H,rho] as 8 real 64^3 products per application, docs/mkdocs/user_guide.md:28-58). */
int qb_engine_apply_bench(qb_engine* e, double t, int nbatch, int reps, int dense, double* ms_out);

/* ---- initial-condition sharding across GPUs: one process (or thread) per GPU, one NCCL all-reduce per evaluation ---------
 * Replaces the MPI_Allreduce calls over comm_init of OptimProblem::evalF/evalGradF (src/optimproblem.cpp:292-298,454-460,527).
 * NCCL is loaded at run time (dlopen of libnccl.so.2); nothing here is needed for single-GPU use.
 *   qb_comm_unique_id   rank 0 creates the 128-byte NCCL id and hands it to the other ranks (any host channel)
 *   qb_engine_comm_init every rank joins the communicator on its engine's device
 *   qb_engine_allreduce in-place SUM of n host doubles over the ranks, staged through the engine's device buffer and ordered
 *                       on the engine's stream (n <= ndesign + 8: the reduced gradient plus the 7 objective scalars) */
int qb_comm_unique_id(char id[128]);
int qb_engine_comm_init(qb_engine* e, const char id[128], int rank, int nranks);
int qb_engine_allreduce(qb_engine* e, double* buf, int n);
/* the same exchange without leaving the device: in-place SUM of the reduced-gradient buffer [ndesign + 8] (qb_engine_grad_dev)
 * over the ranks, enqueued on the engine's stream after the adjoint sweep (asynchronous; the gradient of
 * qb_engine_download_grad is then the global one).  = MPI_Allreduce(grad), src/optimproblem.cpp:527 */
int qb_engine_allreduce_grad_dev(qb_engine* e);

int qb_engine_sync(qb_engine* e);
int qb_engine_get_stats(qb_engine* e, qb_stats* out);

#ifdef __cplusplus
}
#endif
#endif
