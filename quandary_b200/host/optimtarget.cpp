#include "optimtarget.hpp"

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>

static void fill_from_file(const std::string& filename, MasterEq* mastereq, long dim_ess, long dim_rho, bool lind, bool quietmode, Vec& out) {
  const long nelems = lind ? 2 * dim_ess * dim_ess : 2 * dim_ess;
  std::vector<double> vec((size_t)nelems);
  read_vector(filename.c_str(), vec.data(), (int)nelems, quietmode);
  const long dim = mastereq->getDim();
  if (lind) {
    for (long i = 0; i < dim_ess * dim_ess; i++) {
      long k = i % dim_ess, j = i / dim_ess;
      if (dim_ess * dim_ess < dim) { k = mapEssToFull(k, mastereq->nlevels, mastereq->nessential); j = mapEssToFull(j, mastereq->nlevels, mastereq->nessential); }
      const long id = getVecID(k, j, dim_rho);
      out[id] = vec[i]; out[id + dim] = vec[i + dim_ess * dim_ess];
    }
  } else {
    for (long i = 0; i < dim_ess; i++) {
      long k = i;
      if (dim_ess < dim) k = mapEssToFull(i, mastereq->nlevels, mastereq->nessential);
      out[k] = vec[i]; out[k + dim] = vec[i + dim_ess];
    }
  }
}

OptimTarget::OptimTarget(std::vector<std::string> target_str, const std::string& objective_str, std::vector<std::string> initcond_str, MasterEq* mastereq,
                         double total_time, std::vector<double> read_gate_rot, Vec& rho_t0, bool quietmode_) {
  dim = mastereq->getDim();
  dim_rho = mastereq->getDimRho();
  dim_ess = mastereq->getDimEss();
  noscillators = mastereq->getNOscillators();
  quietmode = quietmode_;
  lindbladtype = mastereq->lindbladtype;
  const bool lind = lindbladtype != LindbladType::NONE;
  rho_t0.assign(2 * (size_t)dim, 0.0);

  if (initcond_str[0] == "file") initcond_type = InitialConditionType::FROMFILE;
  else if (initcond_str[0] == "pure") initcond_type = InitialConditionType::PURE;
  else if (initcond_str[0] == "ensemble") initcond_type = InitialConditionType::ENSEMBLE;
  else if (initcond_str[0] == "performance") initcond_type = InitialConditionType::PERFORMANCE;
  else if (initcond_str[0] == "3states") initcond_type = InitialConditionType::THREESTATES;
  else if (initcond_str[0] == "Nplus1") initcond_type = InitialConditionType::NPLUSONE;
  else if (initcond_str[0] == "diagonal") initcond_type = InitialConditionType::DIAGONAL;
  else if (initcond_str[0] == "basis") initcond_type = InitialConditionType::BASIS;
  else { printf("ERROR: Unknown initial condition type %s!\n", initcond_str[0].c_str()); exit(1); }
  if (!lind) {
    if (initcond_type == InitialConditionType::ENSEMBLE || initcond_type == InitialConditionType::THREESTATES || initcond_type == InitialConditionType::NPLUSONE) {
      printf("\n\n ERROR for initial condition setting: \n When running Schroedingers solver (collapse_type == NONE), the initial condition needs to be either 'pure' or 'from file' or 'diagonal' or 'basis'.\n\n");
      exit(1);
    } else if (initcond_type == InitialConditionType::BASIS) initcond_type = InitialConditionType::DIAGONAL;
  }
  std::string initfile = (initcond_type == InitialConditionType::FROMFILE && initcond_str.size() >= 2) ? initcond_str[1] : "";
  if (initcond_str.size() < 2)
    for (size_t j = 0; j < noscillators; j++) initcond_str.push_back(std::to_string(j));
  for (size_t i = 1; i < initcond_str.size(); i++) initcond_IDs.push_back((size_t)atoi(initcond_str[i].c_str()));

  if (initcond_type == InitialConditionType::PURE) {
    if (initcond_IDs.size() != noscillators) { printf("ERROR during pure-state initialization: List of IDs must contain %zu elements!\n", noscillators); exit(1); }
    long diag_id = 0;
    for (size_t k = 0; k < initcond_IDs.size(); k++) {
      if (initcond_IDs[k] > mastereq->getOscillator(k)->getNLevels() - 1) {
        printf("ERROR in config setting. The requested pure state initialization |%zu> exceeds the number of allowed levels for that oscillator (%zu).\n", initcond_IDs[k], mastereq->getOscillator(k)->getNLevels());
        exit(1);
      }
      long dim_postkron = 1;
      for (size_t m = k + 1; m < initcond_IDs.size(); m++) dim_postkron *= (long)mastereq->getOscillator(m)->getNLevels();
      diag_id += (long)initcond_IDs[k] * dim_postkron;
    }
    rho_t0[lind ? getVecID(diag_id, diag_id, dim_rho) : diag_id] = 1.0;
  } else if (initcond_type == InitialConditionType::FROMFILE) {
    fill_from_file(initfile, mastereq, dim_ess, dim_rho, lind, quietmode, rho_t0);
  } else if (initcond_type == InitialConditionType::ENSEMBLE) {
    long dimpost = 1, dimsub = 1;
    for (size_t i = 0; i < noscillators; i++) {
      if (initcond_IDs[0] <= i && i <= initcond_IDs.back()) dimsub *= (long)mastereq->nessential[i];
      else dimpost *= (long)mastereq->nessential[i];
    }
    for (long i = 0; i < dimsub; i++)
      for (long j = i; j < dimsub; j++) {
        long ifull = i * dimpost, jfull = j * dimpost;
        if (dim_ess < dim_rho) { ifull = mapEssToFull(ifull, mastereq->nlevels, mastereq->nessential); jfull = mapEssToFull(jfull, mastereq->nlevels, mastereq->nessential); }
        if (i == j) rho_t0[getVecID(ifull, jfull, dim_rho)] = 1. / dimsub;
        else {
          long id = getVecID(ifull, jfull, dim_rho);
          rho_t0[id] = 0.5 / (dimsub * dimsub); rho_t0[id + dim] = 0.5 / (dimsub * dimsub);
          id = getVecID(jfull, ifull, dim_rho);
          rho_t0[id] = 0.5 / (dimsub * dimsub); rho_t0[id + dim] = -0.5 / (dimsub * dimsub);
        }
      }
  }

  purestateID = -1;
  if (target_str[0] == "gate") {
    target_type = TargetType::GATE;
    if (target_str.size() < 2) { printf("ERROR: You want to optimize for a gate, but didn't specify which one. Check your config for 'optim_target'!\n"); exit(1); }
    copyLast(read_gate_rot, noscillators);
    std::vector<double> gate_rot_freq(noscillators, 0.0);
    for (size_t i = 0; i < noscillators; i++) gate_rot_freq[i] = read_gate_rot[0] < 1e20 ? read_gate_rot[i] : 0.0;
    targetgate = new Gate(target_str[1], target_str.size() > 2 ? target_str[2] : "", mastereq->nlevels, mastereq->nessential, total_time, gate_rot_freq, lindbladtype, quietmode);
  } else if (target_str[0] == "pure") {
    target_type = TargetType::PURE;
    purestateID = 0;
    if (target_str.size() < 2) printf("# Warning: You want to prepare a pure state, but didn't specify which one. Taking default: ground-state |0...0> \n");
    else {
      if (target_str.size() - 1 < noscillators) copyLast(target_str, noscillators + 1);
      for (size_t i = 0; i < noscillators; i++) {
        const size_t Qi = (size_t)atoi(target_str[i + 1].c_str());
        if (Qi >= mastereq->getOscillator(i)->getNLevels()) {
          printf("ERROR in config setting. The requested pure state target |%zu> exceeds the number of modeled levels for that oscillator (%zu).\n", Qi, mastereq->getOscillator(i)->getNLevels());
          exit(1);
        }
        purestateID += (long)Qi * mastereq->getOscillator(i)->dim_postOsc;
      }
    }
  } else if (target_str[0] == "file") {
    target_type = TargetType::FROMFILE;
    assert(target_str.size() >= 2);
  } else { printf("\n\n ERROR: Unknown optimization target: %s\n", target_str[0].c_str()); exit(1); }

  if (objective_str == "Jfrobenius") objective_type = ObjectiveType::JFROBENIUS;
  else if (objective_str == "Jtrace") objective_type = ObjectiveType::JTRACE;
  else if (objective_str == "Jmeasure") objective_type = ObjectiveType::JMEASURE;
  else { printf("\n\n ERROR: Unknown objective function: %s\n", objective_str.c_str()); exit(1); }

  if (target_type == TargetType::GATE || target_type == TargetType::FROMFILE) targetstate.assign(2 * (size_t)dim, 0.0);
  if (target_type == TargetType::FROMFILE) fill_from_file(target_str[1], mastereq, dim_ess, dim_rho, lind, quietmode, targetstate);
}

OptimTarget::~OptimTarget() { delete targetgate; }

int OptimTarget::prepareInitialState(int iinit, int ninit, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential, Vec& rho0) {
  const bool lind = lindbladtype != LindbladType::NONE;
  int initID = 0;
  long dim_post;
  switch (initcond_type) {
    case InitialConditionType::FROMFILE:
    case InitialConditionType::PURE:
    case InitialConditionType::ENSEMBLE:
      break;     /* already stored */
    case InitialConditionType::PERFORMANCE: {
      /* src/optimtarget.cpp:460-481, as written there: psi = (1 + i)/sqrt(2N) in every component (Schroedinger); for the Lindblad
       * solver the first dim_rho entries of the real part are set to 1/dim_rho (the reference indexes the vector with i, not
       * with the diagonal position getVecID(i,i) it computes) */
      std::fill(rho0.begin(), rho0.end(), 0.0);
      const long n = (long)dim_rho;
      if (!lind) {
        const double val = 1. / sqrt(2. * n);
        for (long i = 0; i < n; i++) { rho0[i] = val; rho0[i + n] = val; }
      } else {
        for (long i = 0; i < n; i++) rho0[i] = 1. / n;
      }
      break;
    }
    case InitialConditionType::THREESTATES:
      assert(lind);
      std::fill(rho0.begin(), rho0.end(), 0.0);
      if (iinit == 0) {
        initID = 1;
        for (long i = 0; i < dim_rho; i++) rho0[getVecID(i, i, dim_rho)] = 2. * (dim_rho - i) / (dim_rho * (dim_rho + 1));
      } else if (iinit == 1) {
        initID = 2;
        for (long i = 0; i < dim_rho; i++)
          for (long j = 0; j < dim_rho; j++) rho0[getVecID(i, j, dim_rho)] = 1. / dim_rho;
      } else if (iinit == 2) {
        initID = 3;
        for (long i = 0; i < dim_rho; i++) rho0[getVecID(i, i, dim_rho)] = 1. / dim_rho;
      } else { printf("ERROR: Wrong initial condition setting! Should never happen.\n"); exit(1); }
      break;
    case InitialConditionType::NPLUSONE:
      assert(lind);
      if (iinit < dim_rho) {
        std::fill(rho0.begin(), rho0.end(), 0.0);
        rho0[getVecID(iinit, iinit, dim_rho)] = 1.0;
      } else if (iinit == dim_rho) {
        for (long i = 0; i < dim_rho; i++)
          for (long j = 0; j < dim_rho; j++) rho0[getVecID(i, j, dim_rho)] = 1.0 / dim_rho;   /* imaginary parts keep their previous values, as in the reference */
      } else { printf("Wrong initial condition index. Should never happen!\n"); exit(1); }
      initID = iinit;
      break;
    case InitialConditionType::DIAGONAL: {
      std::fill(rho0.begin(), rho0.end(), 0.0);
      dim_post = 1;
      for (size_t k = initcond_IDs.back() + 1; k < nessential.size(); k++) dim_post *= (long)nessential[k];
      long diagelem = iinit * dim_post;
      if (dim_ess < dim_rho) diagelem = mapEssToFull(diagelem, nlevels, nessential);
      rho0[lind ? getVecID(diagelem, diagelem, dim_rho) : diagelem] = 1.0;
      initID = lind ? iinit * ninit + iinit : iinit;
      break;
    }
    case InitialConditionType::BASIS: {
      assert(lind);
      std::fill(rho0.begin(), rho0.end(), 0.0);
      dim_post = 1;
      for (size_t k = initcond_IDs.back() + 1; k < nessential.size(); k++) dim_post *= (long)nessential[k];
      const int sq = (int)sqrt((double)ninit);
      long k = iinit % sq, j = iinit / sq;
      initID = (int)(j * sq + k);
      k *= dim_post; j *= dim_post;
      if (dim_ess < dim_rho) { k = mapEssToFull(k, nlevels, nessential); j = mapEssToFull(j, nlevels, nessential); }
      if (k == j) rho0[getVecID(k, k, dim_rho)] = 1.0;
      else {
        rho0[getVecID(k, k, dim_rho)] = 0.5; rho0[getVecID(j, j, dim_rho)] = 0.5;
        if (k < j) { rho0[getVecID(k, j, dim_rho)] = 0.5; rho0[getVecID(j, k, dim_rho)] = 0.5; }
        else { rho0[getVecID(k, j, dim_rho) + dim] = -0.5; rho0[getVecID(j, k, dim_rho) + dim] = 0.5; }
      }
      break;
    }
  }
  return initID;
}

void OptimTarget::prepareTargetState(const Vec& rho_t0) {
  if (target_type == TargetType::GATE && !targetgate->isDummy()) targetgate->applyGate(rho_t0, targetstate);
  double nrm2 = 0.0;
  for (double v : rho_t0) nrm2 += v * v;
  purity_rho0 = sqrt(nrm2);
  purity_rho0 = purity_rho0 * purity_rho0;
}

void OptimTarget::HilbertSchmidtOverlap(const Vec& state, bool scalebypurity, double* HS_re_ptr, double* HS_im_ptr) {
  const bool lind = lindbladtype != LindbladType::NONE;
  double HS_re = 0.0, HS_im = 0.0;
  if (target_type == TargetType::PURE) {
    const long idm = lind ? getVecID(purestateID, purestateID, dim_rho) : purestateID;
    HS_re = state[idm]; HS_im = state[idm + dim];
  } else if (lind) {
    for (size_t i = 0; i < state.size(); i++) HS_re += targetstate[i] * state[i];
  } else {
    for (long i = 0; i < dim; i++) {
      HS_re += targetstate[i] * state[i] + targetstate[i + dim] * state[i + dim];
      HS_im += -targetstate[i + dim] * state[i] + targetstate[i] * state[i + dim];
    }
  }
  if (scalebypurity) HS_re = HS_re / purity_rho0;
  *HS_re_ptr = HS_re; *HS_im_ptr = HS_im;
}

void OptimTarget::HilbertSchmidtOverlap_diff(Vec& statebar, bool scalebypurity, double HS_re_bar, double HS_im_bar) {
  const bool lind = lindbladtype != LindbladType::NONE;
  const double scale = scalebypurity ? 1. / purity_rho0 : 1.0;
  if (target_type == TargetType::PURE) {
    const long idm = lind ? getVecID(purestateID, purestateID, dim_rho) : purestateID;
    statebar[idm] += HS_re_bar * scale; statebar[idm + dim] += HS_im_bar;
  } else if (lind) {
    for (size_t i = 0; i < statebar.size(); i++) statebar[i] += HS_re_bar * scale * targetstate[i];
  } else {
    for (long i = 0; i < dim; i++) {
      statebar[i] += targetstate[i] * HS_re_bar * scale - targetstate[i + dim] * HS_im_bar;
      statebar[i + dim] += targetstate[i + dim] * HS_re_bar * scale + targetstate[i] * HS_im_bar;
    }
  }
}

void OptimTarget::evalJ(const Vec& state, double* J_re_ptr, double* J_im_ptr) {
  const bool lind = lindbladtype != LindbladType::NONE;
  double J_re = 0.0, J_im = 0.0;
  switch (objective_type) {
    case ObjectiveType::JFROBENIUS: {
      double nrm2 = 0.0;
      if (target_type == TargetType::GATE || target_type == TargetType::FROMFILE) {
        for (size_t i = 0; i < state.size(); i++) { const double d = targetstate[i] - state[i]; nrm2 += d * d; }
      } else {
        const long diagID = lind ? getVecID(purestateID, purestateID, dim_rho) : purestateID;
        for (size_t i = 0; i < state.size(); i++) { const double d = state[i] - ((long)i == diagID ? 1.0 : 0.0); nrm2 += d * d; }
      }
      const double norm = sqrt(nrm2);
      J_re = norm * norm / 2.0;
      break;
    }
    case ObjectiveType::JTRACE:
      HilbertSchmidtOverlap(state, true, &J_re, &J_im);
      break;
    case ObjectiveType::JMEASURE: {
      if (target_type != TargetType::PURE) { printf("Error: Wrong setting for objective function. Jmeasure can only be used for 'pure' targets.\n"); exit(1); }
      double sum = 0.0;
      for (long i = 0; i < dim_rho; i++) {
        const double rhoii = lind ? state[getVecID(i, i, dim_rho)] : pow(state[i], 2.0) + pow(state[i + dim], 2.0);
        sum += fabs((double)(i - purestateID)) * rhoii;
      }
      J_re = sum;
      break;
    }
  }
  *J_re_ptr = J_re; *J_im_ptr = J_im;
}

void OptimTarget::evalJ_diff(const Vec& state, Vec& statebar, double J_re_bar, double J_im_bar) {
  const bool lind = lindbladtype != LindbladType::NONE;
  switch (objective_type) {
    case ObjectiveType::JFROBENIUS:
      if (target_type == TargetType::GATE || target_type == TargetType::FROMFILE) {
        const double Jbar = J_re_bar / 2.0;
        for (size_t i = 0; i < state.size(); i++) statebar[i] += 2.0 * Jbar * state[i];
        for (size_t i = 0; i < state.size(); i++) statebar[i] += -2.0 * Jbar * targetstate[i];
      } else {
        for (size_t i = 0; i < state.size(); i++) statebar[i] += J_re_bar * state[i];
        statebar[lind ? getVecID(purestateID, purestateID, dim_rho) : purestateID] += -1.0 * J_re_bar;
      }
      break;
    case ObjectiveType::JTRACE:
      HilbertSchmidtOverlap_diff(statebar, true, J_re_bar, J_im_bar);
      break;
    case ObjectiveType::JMEASURE:
      for (long i = 0; i < dim_rho; i++) {
        const double lambdai = fabs((double)(i - purestateID));
        if (lind) statebar[getVecID(i, i, dim_rho)] += lambdai * J_re_bar;
        else { statebar[i] += 2. * J_re_bar * lambdai * state[i]; statebar[i + dim] += 2. * J_re_bar * lambdai * state[i + dim]; }
      }
      break;
  }
}

double OptimTarget::finalizeJ(double obj_cost_re, double obj_cost_im) {
  if (objective_type == ObjectiveType::JTRACE) {
    if (lindbladtype == LindbladType::NONE) return 1.0 - (pow(obj_cost_re, 2.0) + pow(obj_cost_im, 2.0));
    return 1.0 - obj_cost_re;
  }
  return obj_cost_re;
}

void OptimTarget::finalizeJ_diff(double obj_cost_re, double obj_cost_im, double* obj_cost_re_bar, double* obj_cost_im_bar) {
  if (objective_type == ObjectiveType::JTRACE) {
    if (lindbladtype == LindbladType::NONE) { *obj_cost_re_bar = -2. * obj_cost_re; *obj_cost_im_bar = -2. * obj_cost_im; }
    else { *obj_cost_re_bar = -1.0; *obj_cost_im_bar = 0.0; }
  } else { *obj_cost_re_bar = 1.0; *obj_cost_im_bar = 0.0; }
}

void OptimTarget::penaltyFunctional(Vec& a, double& c0, Vec& d) const {
  const bool lind = lindbladtype != LindbladType::NONE;
  a.assign(2 * (size_t)dim, 0.0);
  d.assign(2 * (size_t)dim, 0.0);
  c0 = 0.0;
  const long idm = purestateID < 0 ? 0 : (lind ? getVecID(purestateID, purestateID, dim_rho) : purestateID);
  switch (objective_type) {
    case ObjectiveType::JFROBENIUS:      /* 1/2 ||x - tau||^2 */
      for (auto& v : d) v = 0.5;
      if (target_type == TargetType::PURE) { a[idm] = -1.0; c0 = 0.5; }
      else {
        double t2 = 0.0;
        for (size_t i = 0; i < a.size(); i++) { a[i] = -targetstate[i]; t2 += targetstate[i] * targetstate[i]; }
        c0 = 0.5 * t2;
      }
      break;
    case ObjectiveType::JTRACE:          /* 1 - <tau, x>/purity (Lindblad) */
      if (!lind) {   /* 1 - (Re<tau,x>/purity)^2 - (Im<tau,x>)^2: the engine's overlap form, a = tau (penaltyIsOverlap) */
        if (target_type == TargetType::PURE) a[idm] = 1.0;
        else a = targetstate;
        break;
      }
      c0 = 1.0;
      if (target_type == TargetType::PURE) a[idm] = -1.0 / purity_rho0;
      else for (size_t i = 0; i < a.size(); i++) a[i] = -targetstate[i] / purity_rho0;
      break;
    case ObjectiveType::JMEASURE:        /* sum_i |i - m| rho_ii  or  sum_i |i - m| |psi_i|^2 */
      for (long i = 0; i < dim_rho; i++) {
        const double lambdai = fabs((double)(i - purestateID));
        if (lind) a[getVecID(i, i, dim_rho)] = lambdai;
        else { d[i] = lambdai; d[i + dim] = lambdai; }
      }
      break;
  }
}
