/*
 * ref_hooks.cpp -- extern "C" function-level hooks into the UNMODIFIED reference classes (compiled from
 * /root/reference/src against the PETSc shim into oracle/_ref/libquandary_ref.so).  TEST INFRASTRUCTURE ONLY.
 *
 * The wiring below builds the reference's objects the way its own entry point does (src/main.cpp:40-370) and then
 * exposes single calls of the hot-path functions so that the C restatement (qoracle.c) and the CUDA engine can be
 * compared with the real reference at function granularity:
 *   qref_rhs_apply       MasterEq::assemble_RHS + MatMult/MatMultTranspose   src/mastereq.cpp:657-678,710-739
 *   qref_eval_control    Oscillator::evalControl                              src/oscillator.cpp:281-337
 *   qref_solve_ode       TimeStepper::solveODE                                src/timestepper.cpp:96-181
 *   qref_solve_adjoint   TimeStepper::solveAdjointODE                         src/timestepper.cpp:184-253
 *   qref_eval_f / qref_eval_grad_f   OptimProblem::evalF / evalGradF          src/optimproblem.cpp:224-538
 *   qref_initial_state   OptimTarget::prepareInitialState                     src/optimtarget.cpp:450-698
 */
#include <cstring>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "config.hpp"
#include "mastereq.hpp"
#include "optimproblem.hpp"
#include "oscillator.hpp"
#include "output.hpp"
#include "timestepper.hpp"
#include "util.hpp"

struct qref_handle {
  std::stringstream log;
  Config* config = nullptr;
  std::vector<size_t> nlevels, nessential;
  Oscillator** oscil_vec = nullptr;
  MasterEq* mastereq = nullptr;
  Output* output = nullptr;
  TimeStepper* timestepper = nullptr;
  OptimProblem* optimctx = nullptr;
  int ninit = 1, ntime = 0;
  double dt = 0.0;
  Vec x = nullptr, y = nullptr, xinit = nullptr, grad = nullptr;
};

static bool g_inited = false;

static void vec_set(Vec v, const double* a, int n) {
  double* p;
  VecGetArray(v, &p);
  memcpy(p, a, sizeof(double) * n);
  VecRestoreArray(v, &p);
}
static void vec_get(Vec v, double* a, int n) {
  const double* p;
  VecGetArrayRead(v, &p);
  memcpy(a, p, sizeof(double) * n);
  VecRestoreArrayRead(v, &p);
}

extern "C" {

/* storefwd: -1 = as src/main.cpp:353-355 decides from runtype, 0/1 = force */
void* qref_create(const char* cfgfile, int storefwd) {
  if (!g_inited) {
    int argc = 0; char** argv = nullptr;
    MPI_Init(&argc, &argv);
    PetscInitialize(&argc, &argv, (char*)0, NULL);
    g_inited = true;
  }
  qref_handle* h = new qref_handle();
  bool quietmode = true;
  h->config = new Config(MPI_COMM_WORLD, h->log, quietmode);
  Config& config = *h->config;
  config.ReadFile(cfgfile);

  int rand_seed = config.GetIntParam("rand_seed", -1, false, false);
  if (rand_seed < 0) rand_seed = 1;
  std::mt19937 rand_engine{};
  rand_engine.seed(rand_seed);

  std::vector<int> nlevels_int;
  config.GetVecIntParam("nlevels", nlevels_int, 0);
  h->nlevels.assign(nlevels_int.begin(), nlevels_int.end());
  std::vector<size_t>& nlevels = h->nlevels;
  int ntime = config.GetIntParam("ntime", 1000);
  double dt = config.GetDoubleParam("dt", 0.01);
  h->ntime = ntime; h->dt = dt;
  std::string runtypestr = config.GetStrParam("runtype", "simulation");

  h->nessential.resize(nlevels.size());
  for (size_t i = 0; i < nlevels.size(); i++) h->nessential[i] = nlevels[i];
  std::vector<int> read_nessential;
  config.GetVecIntParam("nessential", read_nessential, -1);
  if (read_nessential[0] > -1) {
    for (size_t i = 0; i < nlevels.size(); i++) {
      if (i < read_nessential.size()) h->nessential[i] = (size_t)read_nessential[i];
      else h->nessential[i] = (size_t)read_nessential[read_nessential.size() - 1];
      if (h->nessential[i] > nlevels[i]) h->nessential[i] = nlevels[i];
    }
  }
  std::vector<size_t>& nessential = h->nessential;

  /* number of initial conditions, src/main.cpp:90-131 */
  int ninit = 1;
  std::vector<std::string> initcondstr;
  config.GetVecStrParam("initialcondition", initcondstr, "none");
  if (initcondstr[0] == "3states") ninit = 3;
  else if (initcondstr[0] == "Nplus1") { ninit = 1; for (size_t i = 0; i < nlevels.size(); i++) ninit *= nlevels[i]; ninit += 1; }
  else if (initcondstr[0] == "diagonal" || initcondstr[0] == "basis") {
    ninit = 1;
    if (initcondstr.size() < 2) for (size_t j = 0; j < nlevels.size(); j++) initcondstr.push_back(std::to_string(j));
    for (size_t i = 1; i < initcondstr.size(); i++) { size_t id = atoi(initcondstr[i].c_str()); if (id < nessential.size()) ninit *= nessential[id]; }
    if (initcondstr[0] == "basis") {
      std::string tmpstr = config.GetStrParam("collapse_type", "none", false);
      if (tmpstr != "none") ninit = (int)pow(ninit, 2.0);
    }
  }
  h->ninit = ninit;

  double total_time = ntime * dt;
  Oscillator** oscil_vec = new Oscillator*[nlevels.size()];
  h->oscil_vec = oscil_vec;
  std::vector<double> trans_freq, rot_freq, selfkerr, decay_time, dephase_time;
  config.GetVecDoubleParam("transfreq", trans_freq, 1e20); copyLast(trans_freq, nlevels.size());
  config.GetVecDoubleParam("rotfreq", rot_freq, 1e20); copyLast(rot_freq, nlevels.size());
  config.GetVecDoubleParam("selfkerr", selfkerr, 0.0); copyLast(selfkerr, nlevels.size());
  std::string lindblad = config.GetStrParam("collapse_type", "none");
  config.GetVecDoubleParam("decay_time", decay_time, 0.0, true, false);
  config.GetVecDoubleParam("dephase_time", dephase_time, 0.0, true, false);
  LindbladType lindbladtype = LindbladType::NONE;
  if (lindblad == "decay") lindbladtype = LindbladType::DECAY;
  else if (lindblad == "dephase") lindbladtype = LindbladType::DEPHASE;
  else if (lindblad == "both") lindbladtype = LindbladType::BOTH;
  copyLast(decay_time, nlevels.size());
  copyLast(dephase_time, nlevels.size());

  std::string default_seg_str = "spline, 10, 0.0, " + std::to_string(total_time);
  std::string default_init_str = "constant, 0.0";
  for (size_t i = 0; i < nlevels.size(); i++) {
    std::vector<double> carrier_freq;
    config.GetVecDoubleParam("carrier_frequency" + std::to_string(i), carrier_freq, 0.0);
    std::vector<std::string> controltype_str, controlinit_str;
    config.GetVecStrParam("control_segments" + std::to_string(i), controltype_str, default_seg_str);
    config.GetVecStrParam("control_initialization" + std::to_string(i), controlinit_str, default_init_str);
    oscil_vec[i] = new Oscillator(config, i, nlevels, controltype_str, controlinit_str, trans_freq[i], selfkerr[i], rot_freq[i],
                                  decay_time[i], dephase_time[i], carrier_freq, total_time, lindbladtype, rand_engine);
    default_seg_str = ""; default_init_str = "";
    for (size_t l = 0; l < controltype_str.size(); l++) default_seg_str += controltype_str[l] += ", ";
    for (size_t l = 0; l < controlinit_str.size(); l++) default_init_str += controlinit_str[l] += ", ";
  }
  std::vector<std::string> pipulse_str;
  config.GetVecStrParam("apply_pipulse", pipulse_str, "none", true, false);
  if (pipulse_str[0] != "none") {
    size_t k = 0;
    while (k + 3 < pipulse_str.size()) {
      size_t id = atoi(pipulse_str[k].c_str());
      for (size_t i = 0; i < nlevels.size(); i++) {
        oscil_vec[i]->pipulse.tstart.push_back(atof(pipulse_str[k + 1].c_str()));
        oscil_vec[i]->pipulse.tstop.push_back(atof(pipulse_str[k + 2].c_str()));
        oscil_vec[i]->pipulse.amp.push_back(i == id ? atof(pipulse_str[k + 3].c_str()) : 0.0);
      }
      k += 4;
    }
  }
  std::vector<double> crosskerr, Jkl;
  config.GetVecDoubleParam("crosskerr", crosskerr, 0.0);
  config.GetVecDoubleParam("Jkl", Jkl, 0.0);
  copyLast(crosskerr, (nlevels.size() - 1) * nlevels.size() / 2);
  copyLast(Jkl, (nlevels.size() - 1) * nlevels.size() / 2);
  bool usematfree = config.GetBoolParam("usematfree", false);
  std::vector<double> eta(nlevels.size() * (nlevels.size() - 1) / 2);
  int idx = 0;
  for (size_t i = 0; i < nlevels.size(); i++) for (size_t j = i + 1; j < nlevels.size(); j++) eta[idx++] = rot_freq[i] - rot_freq[j];
  std::string hsys = config.GetStrParam("hamiltonian_file_Hsys", "none", true, false);
  std::string hc = config.GetStrParam("hamiltonian_file_Hc", "none", true, false);
  if (hsys != "none" || hc != "none") usematfree = false;
  h->mastereq = new MasterEq(nlevels, nessential, oscil_vec, crosskerr, Jkl, eta, lindbladtype, usematfree, hsys, hc, quietmode);
  h->output = new Output(config, MPI_COMM_WORLD, MPI_COMM_WORLD, nlevels.size(), quietmode);

  LinearSolverType linsolvetype = LinearSolverType::GMRES;
  std::string linsolvestr = config.GetStrParam("linearsolver_type", "gmres");
  int linsolve_maxiter = config.GetIntParam("linearsolver_maxiter", 10);
  if (linsolvestr == "neumann") linsolvetype = LinearSolverType::NEUMANN;
  bool storeFWD = false;
  if (lindbladtype != LindbladType::NONE && (runtypestr == "gradient" || runtypestr == "optimization")) storeFWD = true;
  if (storefwd >= 0) storeFWD = storefwd != 0;
  std::string ts = config.GetStrParam("timestepper", "IMR");
  if (ts == "IMR") h->timestepper = new ImplMidpoint(h->mastereq, ntime, total_time, linsolvetype, linsolve_maxiter, h->output, storeFWD);
  else if (ts == "IMR4") h->timestepper = new CompositionalImplMidpoint(4, h->mastereq, ntime, total_time, linsolvetype, linsolve_maxiter, h->output, storeFWD);
  else if (ts == "IMR8") h->timestepper = new CompositionalImplMidpoint(8, h->mastereq, ntime, total_time, linsolvetype, linsolve_maxiter, h->output, storeFWD);
  else h->timestepper = new ExplEuler(h->mastereq, ntime, total_time, h->output, storeFWD);
  h->optimctx = new OptimProblem(config, h->timestepper, MPI_COMM_WORLD, MPI_COMM_WORLD, ninit, h->output, quietmode);

  int n2 = 2 * h->mastereq->getDim();
  VecCreateSeq(PETSC_COMM_SELF, n2, &h->x); VecSetFromOptions(h->x);
  VecCreateSeq(PETSC_COMM_SELF, n2, &h->y); VecSetFromOptions(h->y);
  VecCreateSeq(PETSC_COMM_SELF, h->optimctx->getNdesign(), &h->xinit); VecSetFromOptions(h->xinit);
  VecCreateSeq(PETSC_COMM_SELF, h->optimctx->getNdesign(), &h->grad); VecSetFromOptions(h->grad);
  h->optimctx->getStartingPoint(h->xinit);
  VecCopy(h->xinit, h->optimctx->xinit);
  h->timestepper->writeTrajectoryDataFiles = false;
  return h;
}

int qref_dim(void* hv) { return ((qref_handle*)hv)->mastereq->getDim(); }
int qref_ndesign(void* hv) { return ((qref_handle*)hv)->optimctx->getNdesign(); }
int qref_ninit(void* hv) { return ((qref_handle*)hv)->ninit; }
int qref_ntime(void* hv) { return ((qref_handle*)hv)->ntime; }

void qref_get_params(void* hv, double* alpha) { qref_handle* h = (qref_handle*)hv; vec_get(h->xinit, alpha, h->optimctx->getNdesign()); }
void qref_set_params(void* hv, const double* alpha) {
  qref_handle* h = (qref_handle*)hv;
  vec_set(h->xinit, alpha, h->optimctx->getNdesign());
  h->mastereq->setControlAmplitudes(h->xinit);
}

void qref_eval_control(void* hv, int osc, double t, double* p, double* q) { ((qref_handle*)hv)->mastereq->getOscillator(osc)->evalControl(t, p, q); }

void qref_rhs_apply(void* hv, double t, int transpose, const double* x, double* y) {
  qref_handle* h = (qref_handle*)hv;
  int n2 = 2 * h->mastereq->getDim();
  vec_set(h->x, x, n2);
  h->mastereq->assemble_RHS(t);
  if (transpose) MatMultTranspose(h->mastereq->getRHS(), h->x, h->y);
  else MatMult(h->mastereq->getRHS(), h->x, h->y);
  vec_get(h->y, y, n2);
}

void qref_drhs_dparams(void* hv, double t, const double* x, const double* xbar, double scale, double* grad) {
  qref_handle* h = (qref_handle*)hv;
  int n2 = 2 * h->mastereq->getDim();
  vec_set(h->x, x, n2);
  vec_set(h->y, xbar, n2);
  VecZeroEntries(h->grad);
  h->mastereq->compute_dRHS_dParams(t, h->x, h->y, scale, h->grad);
  vec_get(h->grad, grad, h->optimctx->getNdesign());
}

void qref_initial_state(void* hv, int iinit, double* rho0) {
  qref_handle* h = (qref_handle*)hv;
  h->timestepper->optim_target->prepareInitialState(iinit, h->ninit, h->mastereq->nlevels, h->mastereq->nessential, h->x);
  vec_get(h->x, rho0, 2 * h->mastereq->getDim());
}

/* traj (may be NULL): [(ntime+1)][2*dim], only if the stepper stores (Lindblad + gradient runtype or forced) */
void qref_solve_ode(void* hv, const double* rho0, double* rhoT, double* traj) {
  qref_handle* h = (qref_handle*)hv;
  int n2 = 2 * h->mastereq->getDim();
  vec_set(h->x, rho0, n2);
  Vec fin = h->timestepper->solveODE(0, h->x);
  vec_get(fin, rhoT, n2);
  if (traj && h->timestepper->storeFWD)
    for (int n = 0; n <= h->ntime; n++) vec_get(h->timestepper->getState(n), traj + (size_t)n * n2, n2);
}

/* must follow qref_solve_ode of the same initial condition (uses the stepper's stored trajectory if Lindblad) */
void qref_solve_adjoint(void* hv, const double* rhoT_bar, const double* finalstate, double* redgrad, double* xbar0) {
  qref_handle* h = (qref_handle*)hv;
  int n2 = 2 * h->mastereq->getDim();
  vec_set(h->x, rhoT_bar, n2);
  vec_set(h->y, finalstate, n2);
  h->timestepper->solveAdjointODE(h->x, h->y, 0.0, 0.0, 0.0);
  vec_get(h->timestepper->redgrad, redgrad, h->optimctx->getNdesign());
  (void)xbar0;
}

double qref_eval_f(void* hv, const double* alpha, double* fidelity) {
  qref_handle* h = (qref_handle*)hv;
  vec_set(h->xinit, alpha, h->optimctx->getNdesign());
  double obj = h->optimctx->evalF(h->xinit);
  if (fidelity) *fidelity = h->optimctx->getFidelity();
  return obj;
}

double qref_eval_grad_f(void* hv, const double* alpha, double* grad, double* fidelity) {
  qref_handle* h = (qref_handle*)hv;
  vec_set(h->xinit, alpha, h->optimctx->getNdesign());
  h->optimctx->evalGradF(h->xinit, h->grad);
  vec_get(h->grad, grad, h->optimctx->getNdesign());
  if (fidelity) *fidelity = h->optimctx->getFidelity();
  return h->optimctx->getObjective();
}

}  // extern "C"
