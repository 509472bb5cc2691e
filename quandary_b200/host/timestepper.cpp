#include "timestepper.hpp"

#include <cstdio>
#include <cstdlib>

static void die() { printf("ERROR: %s\n", qb_last_error()); exit(1); }

TimeStepper::TimeStepper(MasterEq* mastereq_, int ntime_, double total_time_, Output* output_, bool storeFWD_) {
  mastereq = mastereq_;
  ntime = ntime_;
  total_time = total_time_;
  dt = total_time / ntime;            /* src/timestepper.cpp:35 */
  output = output_;
  storeFWD = storeFWD_;
  size_t nd = 0;
  for (size_t k = 0; k < mastereq->getNOscillators(); k++) nd += mastereq->getOscillator(k)->getNParams();
  redgrad.assign(nd, 0.0);
  for (size_t i = 0; i < mastereq->getNOscillators(); i++)
    if (mastereq->nessential[i] < mastereq->nlevels[i]) addLeakagePrevent = true;      /* src/timestepper.cpp:29-32 */
}

TimeStepper::~TimeStepper() {
  if (engine) qb_engine_destroy(engine);
}

void TimeStepper::createEngine(int stepper, LinearSolverType linsolve_type, int linsolve_maxiter) {
  qb_system_desc d;
  mastereq->fillDescriptor(d);
  d.ntime = ntime;
  d.total_time = total_time;
  d.stepper = stepper;
  d.linsolve_type = linsolve_type == LinearSolverType::GMRES ? QB_LINSOLVE_GMRES : QB_LINSOLVE_NEUMANN;
  d.linsolve_maxiter = linsolve_maxiter;
  d.linsolve_abstol = 1e-10;          /* src/timestepper.cpp:535-536 */
  d.linsolve_reltol = 1e-20;
  const char* dev = getenv("QB_DEVICE");
  d.device = dev ? atoi(dev) : 0;
  d.flags = QB_FLAG_STORE_STAGE;      /* same values as the reference's re-solved stage, one solve less per adjoint step */
  if (qb_engine_create(&d, &engine) != QB_OK) die();
  mastereq->attachEngine(engine);
}

void TimeStepper::checkPenalties() const {
  /* dpdm: Schroedinger + IMR (the engine reports anything else); nothing to reject here */
}

std::vector<double> TimeStepper::penaltyWeights() const {
  std::vector<double> w((size_t)ntime + 1, 0.0);
  if (penalty_param > 1e-13)
    for (int n = 0; n <= ntime; n++) {
      const double time = n * dt;
      const double weight = 1. / penalty_param * exp(-pow((time - total_time) / penalty_param, 2));
      w[n] = weight * dt;
    }
  return w;
}

double TimeStepper::energyPenalty() const {
  double pen = 0.0;
  for (int n = 0; n < ntime; n++) {
    const double tstop = (n + 1) * dt;
    for (size_t k = 0; k < mastereq->getNOscillators(); k++) {
      double p, q;
      mastereq->getOscillator(k)->evalControl(tstop, &p, &q);
      pen += (p * p + q * q) / ntime;
    }
  }
  return pen;
}

void TimeStepper::energyPenalty_diff(double penaltybar, Vec& grad) const {
  for (int n = ntime; n > 0; n--) {
    const double tstop = n * dt;
    size_t col_shift = 0;
    for (size_t k = 0; k < mastereq->getNOscillators(); k++) {
      double p, q;
      mastereq->getOscillator(k)->evalControl(tstop, &p, &q);
      const double pbar = penaltybar / ntime * 2.0 * p, qbar = penaltybar / ntime * 2.0 * q;
      mastereq->getOscillator(k)->evalControl_diff(tstop, grad.data() + col_shift, pbar, qbar);
      col_shift += mastereq->getOscillator(k)->getNParams();
    }
  }
}

void TimeStepper::solveODE_batch(const std::vector<double>& rho0s, int ninit, std::vector<double>& finalstates, int store) {
  finalstates.resize(rho0s.size());
  if (qb_engine_forward(engine, rho0s.data(), ninit, store, finalstates.data()) != QB_OK) die();
  penalty_integral = 0.0; penalty_dpdm = 0.0; energy_penalty_integral = 0.0;
}

void TimeStepper::solveAdjointODE_batch(const std::vector<double>& rho_bars, int ninit) {
  if (qb_engine_adjoint(engine, rho_bars.data(), ninit, redgrad.data(), nullptr) != QB_OK) die();
}

void TimeStepper::solveODE_batch_dev(const std::vector<double>* rho0s, int ninit, int store) {
  if (rho0s && qb_engine_upload_states(engine, rho0s->data(), ninit, 0) != QB_OK) die();
  if (qb_engine_forward_dev(engine, ninit, store) != QB_OK) die();
  penalty_integral = 0.0; penalty_dpdm = 0.0; energy_penalty_integral = 0.0;
}

void TimeStepper::solveAdjointODE_batch_dev(int ninit) {
  if (qb_engine_adjoint_dev(engine, ninit) != QB_OK) die();
  if (qb_engine_download_grad(engine, redgrad.data()) != QB_OK) die();
}

Vec TimeStepper::solveODE(int, const Vec& rho_t0) {
  Vec fin;
  solveODE_batch(rho_t0, 1, fin, (storeFWD || writeTrajectoryDataFiles) ? 1 : 0);
  return fin;
}

void TimeStepper::solveAdjointODE(const Vec& rho_t0_bar, const Vec&, double, double, double) { solveAdjointODE_batch(rho_t0_bar, 1); }

Vec TimeStepper::getState(size_t tindex, int iinit) {
  Vec out(2 * (size_t)mastereq->getDim());
  if (qb_engine_get_state(engine, iinit, (int)tindex, out.data()) != QB_OK) {
    printf("ERROR: Time-stepper requested state at time index %zu, but didn't store it.\n", tindex);   /* src/timestepper.cpp:88-91 */
    exit(1);
  }
  return out;
}

ExplEuler::ExplEuler(MasterEq* mastereq_, int ntime_, double total_time_, Output* output_, bool storeFWD_, bool create_engine)
    : TimeStepper(mastereq_, ntime_, total_time_, output_, storeFWD_) {
  if (create_engine) createEngine(QB_STEPPER_EE, LinearSolverType::GMRES, 0);
}

ImplMidpoint::ImplMidpoint(MasterEq* mastereq_, int ntime_, double total_time_, LinearSolverType linsolve_type_, int linsolve_maxiter_, Output* output_, bool storeFWD_,
                           int order, bool create_engine)
    : TimeStepper(mastereq_, ntime_, total_time_, output_, storeFWD_) {
  if (order != 2) printf("Timestepper: Compositional Impl. Midpoint, order %d, %d stages\n", order, order == 8 ? 15 : 3);
  if (create_engine) createEngine(order == 8 ? QB_STEPPER_IMR8 : (order == 4 ? QB_STEPPER_IMR4 : QB_STEPPER_IMR), linsolve_type_, linsolve_maxiter_);
}
