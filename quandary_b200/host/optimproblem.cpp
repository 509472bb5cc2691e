#include "optimproblem.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <iomanip>
#include <iostream>

OptimProblem::OptimProblem(const Config& config, TimeStepper* timestepper_, int ninit_, Output* output_, bool quietmode_, int rank_init, int size_init,
                           AllreduceFn allreduce, void* allreduce_user_) {
  timestepper = timestepper_;
  ninit = ninit_;
  output = output_;
  quietmode = quietmode_;
  mpirank_init = rank_init;
  mpisize_init = size_init;
  allreduce_fn = allreduce;
  allreduce_user = allreduce_user_;
  if (ninit % mpisize_init != 0) {
    printf("ERROR: Number of initial conditions (%d) must be divisible by the number of processes for them (%d).\n", ninit, mpisize_init);   /* src/main.cpp:150-153 */
    exit(1);
  }
  ninit_local = ninit / mpisize_init;
  c0 = 0;
  cn = ninit_local;
  MasterEq* mastereq = timestepper->mastereq;
  ndesign = 0;
  for (size_t k = 0; k < mastereq->getNOscillators(); k++) ndesign += (int)mastereq->getOscillator(k)->getNParams();
  if (!quietmode) std::cout << "Number of control parameters: " << ndesign << std::endl;

  std::vector<std::string> target_str, initcond_str;
  config.GetVecStrParam("optim_target", target_str, "pure");
  std::string objective_str = config.GetStrParam("optim_objective", "Jfrobenius");
  std::vector<double> read_gate_rot;
  config.GetVecDoubleParam("gate_rot_freq", read_gate_rot, 1e20, true, false);
  config.GetVecStrParam("initialcondition", initcond_str, "none", false);
  optim_target = new OptimTarget(target_str, objective_str, initcond_str, mastereq, timestepper->total_time, read_gate_rot, rho_t0, quietmode);
  rho_t0_bar.assign(rho_t0.size(), 0.0);

  /* weights beta_i, normalised to sum 1, this process' block (src/optimproblem.cpp:71-91) */
  config.GetVecDoubleParam("optim_weights", obj_weights, 1.0);
  copyLast(obj_weights, (size_t)ninit);
  double scale = 0.0;
  for (int i = 0; i < ninit; i++) scale += obj_weights[i];
  for (int i = 0; i < ninit; i++) obj_weights[i] = obj_weights[i] / scale;
  std::vector<double> mine(obj_weights.begin() + mpirank_init * ninit_local, obj_weights.begin() + (mpirank_init + 1) * ninit_local);
  obj_weights = mine;

  gamma_tik = config.GetDoubleParam("optim_regul", 1e-4);
  gatol = config.GetDoubleParam("optim_atol", 1e-8);
  fatol = config.GetDoubleParam("optim_ftol", 1e-8);
  inftol = config.GetDoubleParam("optim_inftol", 1e-5);
  grtol = config.GetDoubleParam("optim_rtol", 1e-4);
  maxiter = config.GetIntParam("optim_maxiter", 200);
  gamma_penalty = config.GetDoubleParam("optim_penalty", 0.0);
  penalty_param = config.GetDoubleParam("optim_penalty_param", 0.5);
  gamma_penalty_energy = config.GetDoubleParam("optim_penalty_energy", 0.0);
  if (config.count("optim_regul_tik0") > 0) gamma_tik_interpolate = config.GetBoolParam("optim_regul_tik0", false, false);
  else if (config.count("optim_regul_interpolate") > 0) gamma_tik_interpolate = config.GetBoolParam("optim_regul_interpolate", false, false);
  else gamma_tik_interpolate = false;
  gamma_penalty_dpdm = config.GetDoubleParam("optim_penalty_dpdm", 0.0);
  gamma_penalty_variation = config.GetDoubleParam("optim_penalty_variation", 0.01);
  if (gamma_penalty_dpdm > 1e-13 && mastereq->lindbladtype != LindbladType::NONE) gamma_penalty_dpdm = 0.0;   /* as the reference: not implemented for Lindblad */
  timestepper->penalty_param = penalty_param;
  timestepper->gamma_penalty = gamma_penalty;
  timestepper->gamma_penalty_dpdm = gamma_penalty_dpdm;
  timestepper->gamma_penalty_energy = gamma_penalty_energy;
  timestepper->optim_target = optim_target;
  timestepper->checkPenalties();

  /* bounds (src/optimproblem.cpp:136-164) */
  xlower.assign(ndesign, 0.0);
  xupper.assign(ndesign, 0.0);
  int col = 0;
  for (size_t iosc = 0; iosc < mastereq->getNOscillators(); iosc++) {
    std::vector<std::string> bound_str;
    config.GetVecStrParam("control_bounds" + std::to_string(iosc), bound_str, "10000.0");
    Oscillator* osc = mastereq->getOscillator(iosc);
    for (size_t iseg = 0; iseg < osc->getNSegments(); iseg++) {
      double boundval = bound_str.size() <= iseg ? atof(bound_str.back().c_str()) : atof(bound_str[iseg].c_str());
      boundval = boundval / (sqrt(2) * osc->getNCarrierfrequencies());
      boundval = boundval * 2.0 * M_PI;
      for (int i = 0; i < osc->getNSegParams((int)iseg); i++) { xupper[col + i] = boundval; xlower[col + i] = -1. * boundval; }
      col = col + osc->getNSegParams((int)iseg);
    }
  }
  std::vector<std::string> controlinit_str;
  config.GetVecStrParam("control_initialization0", controlinit_str, "constant, 0.0");
  if (!controlinit_str.empty() && controlinit_str[0] == "file") {
    assert(controlinit_str.size() >= 2);
    initguess_fromfile.assign(ndesign, 0.0);
    read_vector(controlinit_str[1].c_str(), initguess_fromfile.data(), ndesign, quietmode);
  }
  xinit.assign(ndesign, 0.0);
  const char* hostobj = getenv("QB_HOST_OBJECTIVE");
  device_objective = !(hostobj && atoi(hostobj) != 0);
}

OptimProblem::~OptimProblem() { delete optim_target; }

void OptimProblem::getStartingPoint(Vec& x) {
  MasterEq* mastereq = timestepper->mastereq;
  x.assign(ndesign, 0.0);
  if (!initguess_fromfile.empty()) x = initguess_fromfile;
  else {
    size_t shift = 0;
    for (size_t k = 0; k < mastereq->getNOscillators(); k++) { mastereq->getOscillator(k)->getParams(x.data() + shift); shift += mastereq->getOscillator(k)->getNParams(); }
  }
  mastereq->setControlAmplitudes(x);
  output->writeControls(x, mastereq, timestepper->ntime, timestepper->dt);
}

int OptimProblem::chunkSize(int store) const {
  int n = ninit_local;
  if (store) {
    const int fit = qb_engine_max_batch(timestepper->getEngine(), store);
    if (fit < 1) { printf("ERROR: not even one trajectory fits the device memory: %s\n", qb_last_error()); exit(1); }
    n = std::min(n, fit);
  }
  const char* forced = getenv("QB_MAX_BATCH");
  if (forced && atoi(forced) > 0) n = std::min(n, atoi(forced));
  return n;
}

bool OptimProblem::installPenalty() {
  MasterEq* mastereq = timestepper->mastereq;
  qb_engine* eng = timestepper->getEngine();
  /* the penalty functional depends on the initial conditions and the target only, not on the design vector */
  if (pen_key0 == c0 && pen_keyn == cn) return pen_cached;
  pen_key0 = c0; pen_keyn = cn; pen_cached = false;
  if (gamma_penalty_dpdm > 1e-13) {
    std::vector<double> jb(cn);
    for (int i = 0; i < cn; i++) jb[i] = obj_weights[c0 + i] * gamma_penalty_dpdm / timestepper->ntime;     /* src/timestepper.cpp:224 */
    if (qb_engine_set_dpdm(eng, jb.data(), cn) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
  }
  if (!(gamma_penalty > 1e-13)) { qb_engine_set_penalty(eng, nullptr); return false; }
  const size_t n2 = 2 * (size_t)mastereq->getDim();
  pen_a.assign((size_t)cn * n2, 0.0);
  pen_c0.assign(cn, 0.0);
  pen_jbar.assign(cn, 0.0);
  pen_purity.assign(cn, 1.0);
  Vec a, d;
  for (int iinit = 0; iinit < cn; iinit++) {
    optim_target->prepareInitialState(mpirank_init * ninit_local + c0 + iinit, ninit, mastereq->nlevels, mastereq->nessential, rho_t0);
    optim_target->prepareTargetState(rho_t0);
    /* weighted-objective integral only for penalty_param > 0 (src/timestepper.cpp:262); otherwise just the guard-level term */
    if (penalty_param > 1e-13) optim_target->penaltyFunctional(a, pen_c0[iinit], d);
    else { a.assign(n2, 0.0); d.assign(n2, 0.0); pen_c0[iinit] = 0.0; }
    std::copy(a.begin(), a.end(), pen_a.begin() + (size_t)iinit * n2);
    pen_jbar[iinit] = obj_weights[c0 + iinit] * gamma_penalty;
    pen_purity[iinit] = optim_target->getPurity();
  }
  pen_d = d;
  pen_wdt = timestepper->penaltyWeights();
  pen_guard.clear();
  if (timestepper->addLeakagePrevent) {
    pen_guard.assign(n2, 0.0);
    const bool lind = mastereq->lindbladtype != LindbladType::NONE;
    const long N = mastereq->getDimRho(), dim = mastereq->getDim();
    for (long i = 0; i < N; i++) {
      /* guard level = LAST level of an oscillator that is not essential (isGuardLevel, src/util.cpp:236-256) */
      long index = i;
      bool guard = false;
      for (size_t k = 0; k < mastereq->nlevels.size(); k++) {
        long postdim = 1;
        for (size_t j = k + 1; j < mastereq->nlevels.size(); j++) postdim *= (long)mastereq->nlevels[j];
        const long itest = index / postdim;
        if (itest == (long)mastereq->nlevels[k] - 1 && itest >= (long)mastereq->nessential[k]) { guard = true; break; }
        index = index % postdim;
      }
      if (guard) { const long id = lind ? getVecID(i, i, N) : i; pen_guard[id] = 1.0; pen_guard[id + dim] = 1.0; }
    }
  }
  qb_penalty_desc pd;
  pd.a = pen_a.data(); pd.c0 = pen_c0.data(); pd.d = pen_d.data(); pd.guard = pen_guard.empty() ? nullptr : pen_guard.data();
  pd.wdt = pen_wdt.data(); pd.jbar = pen_jbar.data(); pd.ninit = cn;
  pd.overlap = (penalty_param > 1e-13 && optim_target->penaltyIsOverlap()) ? 1 : 0;
  pd.purity = pen_purity.data();
  if (qb_engine_set_penalty(eng, &pd) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
  pen_cached = true;
  return true;
}

void OptimProblem::installObjective() {
  if (!device_objective || (obj_key0 == c0 && obj_keyn == cn)) return;
  MasterEq* mastereq = timestepper->mastereq;
  const size_t n2 = 2 * (size_t)mastereq->getDim();
  const long pure = optim_target->getPureStateID();
  std::vector<double> targets, purity(cn);
  if (pure < 0) targets.assign((size_t)cn * n2, 0.0);
  for (int iinit = 0; iinit < cn; iinit++) {
    optim_target->prepareInitialState(mpirank_init * ninit_local + c0 + iinit, ninit, mastereq->nlevels, mastereq->nessential, rho_t0);
    optim_target->prepareTargetState(rho_t0);
    purity[iinit] = optim_target->getPurity();
    if (pure < 0) std::copy(optim_target->getTargetState().begin(), optim_target->getTargetState().end(), targets.begin() + (size_t)iinit * n2);
  }
  qb_objective_desc od;
  od.objective = optim_target->getObjectiveType() == ObjectiveType::JFROBENIUS ? QB_OBJ_JFROBENIUS
               : optim_target->getObjectiveType() == ObjectiveType::JTRACE ? QB_OBJ_JTRACE : QB_OBJ_JMEASURE;
  od.pure_id = (int)pure;
  od.target = pure < 0 ? targets.data() : nullptr;
  od.purity = purity.data();
  od.weight = obj_weights.data() + c0;
  od.fid_scale = 1. / ninit;
  od.ninit = cn;
  if (qb_engine_set_objective(timestepper->getEngine(), &od) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
  obj_key0 = c0; obj_keyn = cn;
}

void OptimProblem::objectiveSums(double* s) {
  double o[4];
  if (qb_engine_objective(timestepper->getEngine(), cn, o) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
  s[3] += o[0]; s[4] += o[1]; s[5] += o[2]; s[6] += o[3];
}

void OptimProblem::forwardAll(int store) {
  MasterEq* mastereq = timestepper->mastereq;
  const size_t n2 = 2 * (size_t)mastereq->getDim();
  const bool resident = device_objective && res_key0 == c0 && res_keyn == cn;     /* this batch's initial states are on the device already */
  if (!resident) {
    rho0s.resize((size_t)cn * n2);
    initids.resize(cn);
    for (int iinit = 0; iinit < cn; iinit++) {
      const int iinit_global = mpirank_init * ninit_local + c0 + iinit;
      initids[iinit] = optim_target->prepareInitialState(iinit_global, ninit, mastereq->nlevels, mastereq->nessential, rho_t0);
      std::copy(rho_t0.begin(), rho_t0.end(), rho0s.begin() + (size_t)iinit * n2);
    }
  }
  if (!device_objective) { timestepper->solveODE_batch(rho0s, cn, finalstates, store); return; }
  installObjective();          /* before the upload: both may grow the engine's state buffers */
  timestepper->solveODE_batch_dev(resident ? nullptr : &rho0s, cn, store);
  res_key0 = c0; res_keyn = cn;
}

void OptimProblem::writeTrajectories() {
  if (!timestepper->writeTrajectoryDataFiles || !output->writesTrajectories()) return;
  for (int iinit = 0; iinit < cn; iinit++) {
    output->openTrajectoryDataFiles("rho", initids[iinit]);
    for (int n = 0; n <= timestepper->ntime; n++) {
      if (n % output->output_frequency != 0) continue;
      output->writeTrajectoryDataFiles(n, n * timestepper->dt, timestepper->getState((size_t)n, iinit), timestepper->mastereq);
    }
    output->closeTrajectoryDataFiles();
  }
}

void OptimProblem::regularisation(const Vec& x) {
  double xnorm2 = 0.0;
  if (!gamma_tik_interpolate) for (double v : x) xnorm2 += v * v;
  else for (int i = 0; i < ndesign; i++) xnorm2 += (x[i] - xinit[i]) * (x[i] - xinit[i]);
  const double xnorm = sqrt(xnorm2);
  obj_regul = gamma_tik / 2. * pow(xnorm, 2.0);
  double var_reg = 0.0;
  for (size_t k = 0; k < timestepper->mastereq->getNOscillators(); k++) var_reg += timestepper->mastereq->getOscillator(k)->evalControlVariation();
  obj_penal_variation = 0.5 * gamma_penalty_variation * var_reg;
}

double OptimProblem::evalF(const Vec& x) {
  MasterEq* mastereq = timestepper->mastereq;
  if (!quietmode) printf("EVAL F... \n");
  mastereq->setControlAmplitudes(x);
  const size_t n2 = 2 * (size_t)mastereq->getDim();
  const bool want_traj = timestepper->writeTrajectoryDataFiles && output->writesTrajectories();
  const bool have_dpdm = gamma_penalty_dpdm > 1e-13;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};   /* penal, penal_dpdm, penal_energy, cost_re, cost_im, fid_re, fid_im */
  const bool need_store = want_traj || gamma_penalty > 1e-13 || have_dpdm;
  const int chunk = chunkSize(need_store ? 1 : 0);
  for (c0 = 0; c0 < ninit_local; c0 += chunk) {      /* initial conditions in batches that fit the device memory */
  cn = std::min(chunk, ninit_local - c0);
  const bool have_pen = installPenalty();
  forwardAll(need_store ? 1 : 0);
  writeTrajectories();
  if (have_dpdm) {
    std::vector<double> dint(cn);
    if (qb_engine_dpdm_integrals(timestepper->getEngine(), cn, dint.data()) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
    for (int i = 0; i < cn; i++) s[1] += obj_weights[c0 + i] * gamma_penalty_dpdm * (dint[i] / timestepper->ntime);
  }
  if (have_pen) {
    std::vector<double> pint(cn);
    if (qb_engine_penalty_integrals(timestepper->getEngine(), cn, pint.data()) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
    for (int i = 0; i < cn; i++) s[0] += obj_weights[c0 + i] * gamma_penalty * pint[i];
  }
  if (gamma_penalty_energy > 1e-13) {
    const double epen = timestepper->energyPenalty();
    for (int i = 0; i < cn; i++) s[2] += obj_weights[c0 + i] * gamma_penalty_energy * epen;
  }
  Vec fin(n2);
  if (device_objective) objectiveSums(s);
  else
  for (int iinit = 0; iinit < cn; iinit++) {
    std::copy(rho0s.begin() + (size_t)iinit * n2, rho0s.begin() + (size_t)(iinit + 1) * n2, rho_t0.begin());
    optim_target->prepareTargetState(rho_t0);
    std::copy(finalstates.begin() + (size_t)iinit * n2, finalstates.begin() + (size_t)(iinit + 1) * n2, fin.begin());
    double jre = 0, jim = 0, fre = 0, fim = 0;
    optim_target->evalJ(fin, &jre, &jim);
    s[3] += obj_weights[c0 + iinit] * jre; s[4] += obj_weights[c0 + iinit] * jim;
    optim_target->HilbertSchmidtOverlap(fin, false, &fre, &fim);
    s[5] += 1. / ninit * fre; s[6] += 1. / ninit * fim;
  }
  }   /* batches */
  c0 = 0; cn = ninit_local;
  if (allreduce_fn && mpisize_init > 1) allreduce_fn(s, 7, allreduce_user);
  obj_penal = s[0]; obj_penal_dpdm = s[1]; obj_penal_energy = s[2];
  fidelity = mastereq->lindbladtype == LindbladType::NONE ? pow(s[5], 2.0) + pow(s[6], 2.0) : s[5];
  obj_cost = optim_target->finalizeJ(s[3], s[4]);
  regularisation(x);
  objective = obj_cost + obj_regul + obj_penal + obj_penal_dpdm + obj_penal_energy + obj_penal_variation;
  if (!quietmode) {
    std::cout << "Objective = " << std::scientific << std::setprecision(14) << obj_cost << " + " << obj_regul << " + " << obj_penal << " + " << obj_penal_dpdm << " + "
              << obj_penal_energy << " + " << obj_penal_variation << std::endl;
    std::cout << "Fidelity = " << fidelity << std::endl;
  }
  return objective;
}

void OptimProblem::evalGradF(const Vec& x, Vec& G) {
  MasterEq* mastereq = timestepper->mastereq;
  const bool lind = mastereq->lindbladtype != LindbladType::NONE;
  if (!quietmode) std::cout << "EVAL GRAD F... " << std::endl;
  mastereq->setControlAmplitudes(x);
  G.assign(ndesign, 0.0);
  if (mpirank_init == 0) {   /* regulariser gradient on one process only (src/optimproblem.cpp:356-372) */
    for (int i = 0; i < ndesign; i++) G[i] += gamma_tik * x[i];
    if (gamma_tik_interpolate) for (int i = 0; i < ndesign; i++) G[i] += -1.0 * gamma_tik * xinit[i];
    const double var_reg_bar = 0.5 * gamma_penalty_variation;
    int skip = 0;
    for (size_t k = 0; k < mastereq->getNOscillators(); k++) {
      mastereq->getOscillator(k)->evalControlVariationDiff(G, var_reg_bar, skip);
      skip += (int)mastereq->getOscillator(k)->getNParams();
    }
  }
  const size_t n2 = 2 * (size_t)mastereq->getDim();
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  /* Lindblad: forward + adjoint per batch of initial conditions that fits the device memory (each initial condition is
   * independent, src/optimproblem.cpp:386-444).  Schroedinger needs all overlaps before any adjoint and stores no trajectory. */
  /* what the forward sweep keeps (src/main.cpp:339-344 sets storeFWD for Lindblad gradients only): the state trajectory when a
   * penalty, the dpdm term or trajectory output reads it; otherwise only the adjoint sweep's input (Lindblad: the stored
   * stages; Schroedinger: nothing, the adjoint sweep recomputes the primal backwards as the reference does) */
  const bool want_traj = timestepper->writeTrajectoryDataFiles && output->writesTrajectories();
  const bool need_states = want_traj || gamma_penalty > 1e-13 || gamma_penalty_dpdm > 1e-13;
  const int store = need_states ? 1 : (lind ? 2 : 0);
  int chunk = ninit_local;
  if (lind) chunk = chunkSize(store);
  for (c0 = 0; c0 < ninit_local; c0 += chunk) {
  cn = std::min(chunk, ninit_local - c0);
  const bool have_pen = installPenalty();
  forwardAll(store);
  writeTrajectories();
  if (!device_objective) rbars.assign((size_t)cn * n2, 0.0);
  if (gamma_penalty_dpdm > 1e-13) {
    std::vector<double> dint(cn);
    if (qb_engine_dpdm_integrals(timestepper->getEngine(), cn, dint.data()) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
    for (int i = 0; i < cn; i++) s[1] += obj_weights[c0 + i] * gamma_penalty_dpdm * (dint[i] / timestepper->ntime);
  }
  if (have_pen) {
    std::vector<double> pint(cn);
    if (qb_engine_penalty_integrals(timestepper->getEngine(), cn, pint.data()) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
    for (int i = 0; i < cn; i++) s[0] += obj_weights[c0 + i] * gamma_penalty * pint[i];
  }
  if (gamma_penalty_energy > 1e-13) {
    const double epen = timestepper->energyPenalty();
    for (int i = 0; i < cn; i++) {
      s[2] += obj_weights[c0 + i] * gamma_penalty_energy * epen;
      timestepper->energyPenalty_diff(obj_weights[c0 + i] * gamma_penalty_energy, G);     /* energyPenaltyIntegral_diff per initial condition */
    }
  }
  Vec fin(n2);
  if (device_objective) {
    /* overlaps, terminal adjoints and the adjoint sweep without moving a state to the host.  Lindblad: the seeds of
     * finalizeJ_diff are constants; Schroedinger: they need the overlap summed over ALL initial conditions first
     * (src/optimproblem.cpp:454-460, 495-512), hence the early all-reduce */
    objectiveSums(s);
    if (!lind && allreduce_fn && mpisize_init > 1) allreduce_fn(s, 7, allreduce_user);
    double re_bar, im_bar;
    optim_target->finalizeJ_diff(s[3], s[4], &re_bar, &im_bar);
    if (qb_engine_objective_bar(timestepper->getEngine(), cn, re_bar, im_bar) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
    timestepper->solveAdjointODE_batch_dev(cn);
    for (int i = 0; i < ndesign; i++) G[i] += timestepper->redgrad[i];
    continue;
  }
  for (int iinit = 0; iinit < cn; iinit++) {
    std::copy(rho0s.begin() + (size_t)iinit * n2, rho0s.begin() + (size_t)(iinit + 1) * n2, rho_t0.begin());
    optim_target->prepareTargetState(rho_t0);
    std::copy(finalstates.begin() + (size_t)iinit * n2, finalstates.begin() + (size_t)(iinit + 1) * n2, fin.begin());
    double jre = 0, jim = 0, fre = 0, fim = 0;
    optim_target->evalJ(fin, &jre, &jim);
    s[3] += obj_weights[c0 + iinit] * jre; s[4] += obj_weights[c0 + iinit] * jim;
    optim_target->HilbertSchmidtOverlap(fin, false, &fre, &fim);
    s[5] += 1. / ninit * fre; s[6] += 1. / ninit * fim;
    if (lind) {   /* terminal adjoint per initial condition, src/optimproblem.cpp:427-436 */
      std::fill(rho_t0_bar.begin(), rho_t0_bar.end(), 0.0);
      double re_bar, im_bar;
      optim_target->finalizeJ_diff(s[3], s[4], &re_bar, &im_bar);
      optim_target->evalJ_diff(fin, rho_t0_bar, obj_weights[c0 + iinit] * re_bar, obj_weights[c0 + iinit] * im_bar);
      std::copy(rho_t0_bar.begin(), rho_t0_bar.end(), rbars.begin() + (size_t)iinit * n2);
    }
  }
  /* Schroedinger needs the summed overlaps before the terminal adjoints can be formed (src/optimproblem.cpp:454-460,495-512);
   * Lindblad does not, so its 7 scalars travel with the gradient in ONE all-reduce at the end */
  if (!lind && allreduce_fn && mpisize_init > 1) allreduce_fn(s, 7, allreduce_user);
  if (!lind) {   /* Schroedinger: terminal adjoints need the summed overlap, src/optimproblem.cpp:495-512 */
    for (int iinit = 0; iinit < cn; iinit++) {
      std::copy(rho0s.begin() + (size_t)iinit * n2, rho0s.begin() + (size_t)(iinit + 1) * n2, rho_t0.begin());
      optim_target->prepareTargetState(rho_t0);
      std::copy(finalstates.begin() + (size_t)iinit * n2, finalstates.begin() + (size_t)(iinit + 1) * n2, fin.begin());
      std::fill(rho_t0_bar.begin(), rho_t0_bar.end(), 0.0);
      double re_bar, im_bar;
      optim_target->finalizeJ_diff(s[3], s[4], &re_bar, &im_bar);
      optim_target->evalJ_diff(fin, rho_t0_bar, obj_weights[c0 + iinit] * re_bar, obj_weights[c0 + iinit] * im_bar);
      std::copy(rho_t0_bar.begin(), rho_t0_bar.end(), rbars.begin() + (size_t)iinit * n2);
    }
  }
  timestepper->solveAdjointODE_batch(rbars, cn);
  for (int i = 0; i < ndesign; i++) G[i] += timestepper->redgrad[i];
  }   /* batches */
  c0 = 0; cn = ninit_local;
  if (allreduce_fn && mpisize_init > 1) {
    if (lind) {   /* [grad | 7 scalars] in one exchange */
      std::vector<double> buf(G);
      buf.insert(buf.end(), s, s + 7);
      allreduce_fn(buf.data(), ndesign + 7, allreduce_user);
      std::copy(buf.begin(), buf.begin() + ndesign, G.begin());
      std::copy(buf.begin() + ndesign, buf.end(), s);
    } else {
      allreduce_fn(G.data(), ndesign, allreduce_user);
    }
  }
  obj_penal = s[0]; obj_penal_dpdm = s[1]; obj_penal_energy = s[2];
  fidelity = !lind ? pow(s[5], 2.0) + pow(s[6], 2.0) : s[5];
  obj_cost = optim_target->finalizeJ(s[3], s[4]);
  regularisation(x);
  objective = obj_cost + obj_regul + obj_penal + obj_penal_dpdm + obj_penal_energy + obj_penal_variation;
  double g2 = 0.0;
  for (double v : G) g2 += v * v;
  gnorm = sqrt(g2);
}

void OptimProblem::solve(Vec& x) {
  const int m = 10;
  auto project = [&](Vec& v) { for (int i = 0; i < ndesign; i++) v[i] = std::min(std::max(v[i], xlower[i]), xupper[i]); };
  auto pgnorm = [&](const Vec& v, const Vec& g) {
    double s = 0.0;
    for (int i = 0; i < ndesign; i++) {
      double gi = g[i];
      if ((v[i] <= xlower[i] && gi > 0) || (v[i] >= xupper[i] && gi < 0)) gi = 0.0;
      s += gi * gi;
    }
    return sqrt(s);
  };
  project(x);
  Vec g(ndesign), xn(ndesign), gn(ndesign), d(ndesign);
  evalGradF(x, g);
  double f = objective;
  std::deque<Vec> S, Y;
  double step = 0.0, pg0 = 0.0;
  for (int iter = 0;; iter++) {
    const double pg = pgnorm(x, g);
    if (iter == 0) pg0 = pg;
    /* stopping tests first, so that the last iterate is always logged (TaoMonitor's lastIter, src/optimproblem.cpp:607-634):
     * infidelity, final-time cost, iteration cap, ||g|| <= optim_atol, and TAO's gttol ||g|| <= optim_rtol * ||g_0||
     * (TaoSetTolerances(tao, gatol, PETSC_DEFAULT, grtol), src/optimproblem.cpp:182) */
    const bool stop = iter >= maxiter || pg <= gatol || pg <= grtol * pg0 || 1.0 - fidelity <= inftol || obj_cost <= fatol;
    const bool logged = stop || iter % output->optim_monitor_freq == 0;
    if (logged) {
      output->writeOptimFile(iter, f, pg, step, fidelity, obj_cost, obj_regul, obj_penal, obj_penal_dpdm, obj_penal_energy, obj_penal_variation);
      output->writeGradient(g);
    }
    if (!quietmode) printf("%5d  objective %1.14e  ||Pr(grad)|| %1.6e  fidelity %1.10f\n", iter, f, pg, fidelity);
    if (stop) break;
    /* two-loop recursion */
    d = g;
    std::vector<double> al(S.size());
    for (int i = (int)S.size() - 1; i >= 0; i--) {
      double sy = 0, sd = 0;
      for (int j = 0; j < ndesign; j++) { sy += S[i][j] * Y[i][j]; sd += S[i][j] * d[j]; }
      al[i] = sd / sy;
      for (int j = 0; j < ndesign; j++) d[j] -= al[i] * Y[i][j];
    }
    if (!S.empty()) {
      double sy = 0, yy = 0;
      for (int j = 0; j < ndesign; j++) { sy += S.back()[j] * Y.back()[j]; yy += Y.back()[j] * Y.back()[j]; }
      for (int j = 0; j < ndesign; j++) d[j] *= sy / yy;
    }
    for (size_t i = 0; i < S.size(); i++) {
      double sy = 0, yd = 0;
      for (int j = 0; j < ndesign; j++) { sy += S[i][j] * Y[i][j]; yd += Y[i][j] * d[j]; }
      const double be = yd / sy;
      for (int j = 0; j < ndesign; j++) d[j] += (al[i] - be) * S[i][j];
    }
    /* projected Armijo backtracking along -d */
    step = 1.0;
    double fn = f;
    bool ok = false;
    for (int ls = 0; ls < 30; ls++) {
      for (int j = 0; j < ndesign; j++) xn[j] = x[j] - step * d[j];
      project(xn);
      double dec = 0.0;
      for (int j = 0; j < ndesign; j++) dec += g[j] * (x[j] - xn[j]);
      evalGradF(xn, gn);
      fn = objective;
      if (fn <= f - 1e-4 * dec && dec > 0) { ok = true; break; }
      step *= 0.5;
    }
    if (!ok) {
      if (S.empty()) {      /* line search failed along the steepest-descent direction: stop, and log this iterate if it was not */
        evalGradF(x, g);    /* restores objective / fidelity members of the current iterate */
        if (!logged) {
          output->writeOptimFile(iter, f, pg, step, fidelity, obj_cost, obj_regul, obj_penal, obj_penal_dpdm, obj_penal_energy, obj_penal_variation);
          output->writeGradient(g);
        }
        break;
      }
      S.clear(); Y.clear(); evalGradF(x, g); continue;
    }
    Vec s(ndesign), y(ndesign);
    double sy = 0;
    for (int j = 0; j < ndesign; j++) { s[j] = xn[j] - x[j]; y[j] = gn[j] - g[j]; sy += s[j] * y[j]; }
    if (sy > 1e-14) { S.push_back(s); Y.push_back(y); if ((int)S.size() > m) { S.pop_front(); Y.pop_front(); } }
    x = xn; g = gn; f = fn;
  }
  /* final iterate: controls and trajectories, as TaoMonitor does on the last iteration (src/optimproblem.cpp:645-650) */
  output->writeControls(x, timestepper->mastereq, timestepper->ntime, timestepper->dt);
  timestepper->writeTrajectoryDataFiles = true;
  evalF(x);
}
