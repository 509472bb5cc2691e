/* optimproblem.hpp -- objective and gradient over all initial conditions; mirrors the reference's OptimProblem
 * (include/optimproblem.hpp:123-256, src/optimproblem.cpp:224-538).  The initial-condition loops of evalF / evalGradF are
 * the seam of the B200 engine: all local initial conditions go through ONE batched forward sweep and ONE batched adjoint sweep. */
#pragma once
#include "optimtarget.hpp"
#include "output.hpp"
#include "timestepper.hpp"

class OptimProblem {
 public:
  TimeStepper* timestepper;
  Output* output;
  OptimTarget* optim_target;
  Vec xinit, xlower, xupper;

  /* rank_init / size_init: position of this process in the initial-condition group (one process per GPU);
   * allreduce: sums a buffer over that group (NCCL / MPI / nothing for a single process) */
  typedef void (*AllreduceFn)(double* buf, int n, void* user);
  OptimProblem(const Config& config, TimeStepper* timestepper_, int ninit_, Output* output_, bool quietmode_, int rank_init = 0, int size_init = 1,
               AllreduceFn allreduce = nullptr, void* allreduce_user = nullptr);
  ~OptimProblem();

  int getNdesign() const { return ndesign; }
  double getObjective() const { return objective; }
  double getCostT() const { return obj_cost; }
  double getRegul() const { return obj_regul; }
  double getPenalty() const { return obj_penal; }
  double getPenaltyDpDm() const { return obj_penal_dpdm; }
  double getPenaltyEnergy() const { return obj_penal_energy; }
  double getPenaltyVariation() const { return obj_penal_variation; }
  double getFidelity() const { return fidelity; }
  double getGnorm() const { return gnorm; }
  int getMaxIter() const { return maxiter; }

  double evalF(const Vec& x);                  /* src/optimproblem.cpp:224-338 */
  void evalGradF(const Vec& x, Vec& G);        /* src/optimproblem.cpp:342-538 */
  void getStartingPoint(Vec& x);               /* src/optimproblem.cpp:545-575 */
  /* Bound-constrained quasi-Newton iteration.  NOT TAO's BQNLS (third-party, not in the reference tree): a projected L-BFGS
   * with Armijo backtracking and the reference's stopping rules (src/optimproblem.cpp:586-660); iterates differ from TAO's. */
  void solve(Vec& x);

 private:
  int ninit, ninit_local, ndesign;
  int mpirank_init, mpisize_init;
  AllreduceFn allreduce_fn;
  void* allreduce_user;
  bool quietmode;
  std::vector<double> obj_weights, initguess_fromfile;
  double gamma_tik, gatol, fatol, inftol, grtol, gamma_penalty, penalty_param, gamma_penalty_energy, gamma_penalty_dpdm, gamma_penalty_variation;
  bool gamma_tik_interpolate;
  int maxiter;
  double objective = 0, obj_cost = 0, obj_regul = 0, obj_penal = 0, obj_penal_dpdm = 0, obj_penal_energy = 0, obj_penal_variation = 0, fidelity = 0, gnorm = 0;
  Vec rho_t0, rho_t0_bar;
  std::vector<double> rho0s, finalstates, rbars;
  std::vector<int> initids;

  int c0 = 0, cn = 0;                    /* current batch of local initial conditions: [c0, c0+cn) */
  int chunkSize(int store) const;       /* largest batch whose trajectory fits the device (QB_MAX_BATCH overrides) */
  void forwardAll(int store);      /* store: 0 none, 1 state trajectory, 2 only what the adjoint sweep reads */
  bool installPenalty();                 /* builds the engine's penalty functional for the local initial conditions */
  std::vector<double> pen_a, pen_c0, pen_d, pen_guard, pen_wdt, pen_jbar, pen_purity;
  /* final-time objective on the device: target states, purities and weights are uploaded once per batch of
   * initial conditions; per evaluation no state crosses PCIe.  QB_HOST_OBJECTIVE=1 keeps the O(N^2) host loops instead. */
  bool device_objective = true;
  void installObjective();
  void objectiveSums(double* s);         /* s[3..6] += the batch's weighted objective / overlap sums */
  int obj_key0 = -1, obj_keyn = -1, pen_key0 = -1, pen_keyn = -1, res_key0 = -1, res_keyn = -1;
  bool pen_cached = false;
  void writeTrajectories();
  void regularisation(const Vec& x);
};
