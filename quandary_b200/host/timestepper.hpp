/* timestepper.hpp -- host-side mirror of the reference's TimeStepper / ImplMidpoint (include/timestepper.hpp:48-197,285).
 * The per-step interface evolveFWD(tstart,tstop,Vec) handles one state per call; the device engine runs whole sweeps for
 * all initial conditions, so besides solveODE / solveAdjointODE (one initial condition, same semantics as the reference)
 * this class offers the batched forms that OptimProblem uses. */
#pragma once
#include "mastereq.hpp"

class Output;
class OptimTarget;

class TimeStepper {
 public:
  MasterEq* mastereq = nullptr;
  int ntime = 0;
  double total_time = 0.0, dt = 0.0;
  bool writeTrajectoryDataFiles = false;
  Vec redgrad;                       /* reduced gradient of the last solveAdjointODE[_batch] */
  OptimTarget* optim_target = nullptr;
  Output* output = nullptr;
  bool storeFWD = false;
  /* penalty_integral of the LAST single-initial-condition solveODE; the batched path reports per-initial-condition
   * values through penaltyIntegrals().  The dpdm (second-derivative) penalty is not implemented on the device. */
  bool addLeakagePrevent = false;
  double penalty_integral = 0.0, penalty_dpdm = 0.0, energy_penalty_integral = 0.0;
  double penalty_param = 0.0, gamma_penalty = 0.0, gamma_penalty_dpdm = 0.0, gamma_penalty_energy = 0.0;

  TimeStepper(MasterEq* mastereq_, int ntime_, double total_time_, Output* output_, bool storeFWD_);
  virtual ~TimeStepper();

  /* = TimeStepper::solveODE (src/timestepper.cpp:96-181), one initial condition */
  Vec solveODE(int initid, const Vec& rho_t0);
  /* = TimeStepper::solveAdjointODE (src/timestepper.cpp:184-253); must follow solveODE of the same initial condition */
  void solveAdjointODE(const Vec& rho_t0_bar, const Vec& finalstate, double Jbar_penalty, double Jbar_penalty_dpdm, double Jbar_energy_penalty);
  /* = TimeStepper::getState (src/timestepper.cpp:86-94) of initial condition iinit of the last forward sweep */
  Vec getState(size_t tindex, int iinit = 0);

  /* all local initial conditions at once: rho0s/finalstates are [ninit][2*dim] */
  void solveODE_batch(const std::vector<double>& rho0s, int ninit, std::vector<double>& finalstates, int store);
  void solveAdjointODE_batch(const std::vector<double>& rho_bars, int ninit);
  /* the same sweeps with states that stay on the device: rho0s == nullptr reuses the initial states already uploaded; the
   * terminal adjoints are the ones qb_engine_objective_bar left on the device */
  void solveODE_batch_dev(const std::vector<double>* rho0s, int ninit, int store);   /* store: 0 none, 1 states (+ stages), 2 only what the adjoint sweep reads */
  void solveAdjointODE_batch_dev(int ninit);
  qb_engine* getEngine() const { return engine; }
  void checkPenalties() const;
  /* w(n*dt)*dt for n = 0..ntime, w(t) = exp(-((t-T)/penalty_param)^2)/penalty_param (src/timestepper.cpp:262-263) */
  std::vector<double> penaltyWeights() const;
  /* = energyPenaltyIntegral summed over the sweep and its derivative (src/timestepper.cpp:444-480): state independent */
  double energyPenalty() const;
  void energyPenalty_diff(double penaltybar, Vec& grad) const;

 protected:
  qb_engine* engine = nullptr;
  void createEngine(int stepper, LinearSolverType linsolve_type, int linsolve_maxiter);
};

/* = ExplEuler (src/timestepper.cpp:484-519; include/timestepper.hpp:206-221): x <- x + dt A(t_n) x */
class ExplEuler : public TimeStepper {
 public:
  ExplEuler(MasterEq* mastereq_, int ntime_, double total_time_, Output* output_, bool storeFWD_, bool create_engine = true);
  ~ExplEuler() {}
};

class ImplMidpoint : public TimeStepper {
 public:
  /* create_engine = false: host objects only (no CUDA call), for `--dry-run` */
  ImplMidpoint(MasterEq* mastereq_, int ntime_, double total_time_, LinearSolverType linsolve_type_, int linsolve_maxiter_, Output* output_, bool storeFWD_,
               int order = 2, bool create_engine = true);
  ~ImplMidpoint() {}
};

/* = CompositionalImplMidpoint (src/timestepper.cpp:729-826): order 4 (3 sub-steps) or 8 (15 sub-steps) of the implicit midpoint rule */
class CompositionalImplMidpoint : public ImplMidpoint {
 public:
  CompositionalImplMidpoint(int order_, MasterEq* mastereq_, int ntime_, double total_time_, LinearSolverType linsolve_type_, int linsolve_maxiter_,
                            Output* output_, bool storeFWD_, bool create_engine = true)
      : ImplMidpoint(mastereq_, ntime_, total_time_, linsolve_type_, linsolve_maxiter_, output_, storeFWD_, order_, create_engine) {}
};
