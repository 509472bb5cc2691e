/*
 * qoracle.h -- CPU restatement ("port") of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the product
 * (quandary_b200/) never links or calls it.  The system description is shared with the product header
 * (plain data only) so that both sides are driven by the same bytes.
 *
 * Pinned: tests/test_oracle_pinned.py checks this port against oracle/_ref (the UNMODIFIED reference sources
 * compiled against the PETSc shim), which oracle/check_ref_goldens.py in turn pins to the reference's golden
 * regression files at the suite's own rtol=1e-7.
 */
#ifndef QORACLE_H
#define QORACLE_H
#include "../include/quandary_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

int qo_dim(const qb_system_desc* d);      /* N^2 or N */
int qo_ndesign(const qb_system_desc* d);

/* Oscillator::evalControl, src/oscillator.cpp:281-337 (alpha = full design vector) */
void qo_eval_control(const qb_system_desc* d, const double* alpha, int osc, double t, double* p, double* q);
/* Oscillator::evalControl_diff, src/oscillator.cpp:339-381; grad = full design gradient */
void qo_eval_control_diff(const qb_system_desc* d, int osc, double t, double* grad, double pbar, double qbar);

/* assemble_RHS + applyRHS_matfree[_transpose], src/mastereq.cpp:657-678,1279-2893 */
void qo_rhs_apply(const qb_system_desc* d, const double* alpha, double t, int transpose, const double* x, double* y);
/* compute_dRHS_dParams_matfree, src/mastereq.cpp:970-1276: grad += alpha_scale * xbar^T dM/dalpha x */
void qo_drhs_dparams(const qb_system_desc* d, double t, const double* x, const double* xbar, double scale, double* grad);

/* TimeStepper::solveODE for one initial condition, src/timestepper.cpp:96-181.
 * traj (may be NULL): [(ntime+1)][2*dim] receives store_states.  Returns linear-solver iterations taken. */
long qo_solve_ode(const qb_system_desc* d, const double* alpha, const double* rho0, double* rhoT, double* traj, long* rhs_evals);
/* TimeStepper::solveAdjointODE, src/timestepper.cpp:184-253.  traj: stored states (Lindblad) or NULL (Schroedinger:
 * primal recomputed backwards from finalstate).  redgrad[ndesign] is overwritten. xbar0 (may be NULL): adjoint at t=0 */
long qo_solve_adjoint(const qb_system_desc* d, const double* alpha, const double* rhoT_bar, const double* finalstate,
                      const double* traj, double* redgrad, double* xbar0, long* rhs_evals);

/* Batched over initial conditions, OpenMP over ICs (the reference's only parallel axis, src/optimproblem.cpp:245-248).
 * rho0,rhoT,rhoT_bar: [ninit][2*dim]; grad[ndesign] = sum over ICs in IC order.  nthreads<=0: all cores.
 * weights_bar callback is not needed: the caller supplies rhoT_bar through cb (called per IC after its forward sweep)
 * or NULL to use rhoT_bar_in directly. */
typedef void (*qo_terminal_cb)(int iinit, const double* rhoT, double* rhoT_bar, void* user);
long qo_batch_forward(const qb_system_desc* d, const double* alpha, const double* rho0, int ninit, double* rhoT, int nthreads,
                      long* rhs_evals);
long qo_batch_gradient(const qb_system_desc* d, const double* alpha, const double* rho0, int ninit, double* rhoT,
                       const double* rhoT_bar_in, qo_terminal_cb cb, void* user, double* grad, int nthreads, long* rhs_evals);
int qo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
