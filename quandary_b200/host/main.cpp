/* main.cpp -- `.cfg`-driven entry point of the B200 engine: `quandary_b200 config.cfg [--quiet]`.
 * Same configuration file, same object wiring and same output files as the reference's src/main.cpp:24-487 (runtypes
 * simulation, gradient, optimization, evalcontrols), with the time stepping of all initial conditions on the GPU. */
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "optimproblem.hpp"

/* One process per GPU (the reference: one MPI rank per group of initial conditions, src/main.cpp:133-175): the parent forks
 * ngpus-1 children BEFORE any CUDA call; rank r drives device r.  The NCCL id travels through an anonymous shared page. */
struct SharedBoot { volatile int id_ready; volatile int failed; char id[128]; };
/* rank 0 on an error path after the fork: collect the other ranks (they leave by themselves once `failed` is set) */
static int reap(const std::vector<pid_t>& children, int status) {
  for (pid_t c : children) { int st = 0; waitpid(c, &st, 0); }
  return status;
}
static void nccl_allreduce(double* buf, int n, void* user) {
  if (qb_engine_allreduce((qb_engine*)user, buf, n) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
}

static double wtime() {
  struct timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

int main(int argc, char** argv) {
  bool quietmode = false;
  std::string cfgfile;
  int ngpus = getenv("QB_GPUS") ? atoi(getenv("QB_GPUS")) : 1;
  bool dryrun = false;      /* parse, build the host objects, write params.dat / control<k>.dat / config_log.dat; no GPU needed */
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--quiet") quietmode = true;
    else if (a == "--gpus" && i + 1 < argc) ngpus = atoi(argv[++i]);
    else if (a == "--dry-run") dryrun = true;
    else if (a == "--version") { printf("quandary_b200 (B200 engine for Quandary 4.4-dev configuration files)\n"); return 0; }
    else if (a == "--help" || a == "-h") { printf("usage: %s <config.cfg> [--quiet] [--gpus G] [--dry-run]\n", argv[0]); return 0; }
    else if (a == "--petsc-options") { i++; if (!quietmode) printf("# Warning: --petsc-options is ignored (no PETSc on this path).\n"); }
    else if (cfgfile.empty()) cfgfile = a;
  }
  if (cfgfile.empty()) { printf("\nUSAGE: %s <config.cfg> [--quiet] [--gpus G]\n", argv[0]); return 1; }
  if (ngpus < 1 || dryrun) ngpus = 1;
  int mpirank = 0;
  /* one non-reproducible seed for ALL ranks, drawn before the processes are forked: the reference draws it on rank 0 and
   * broadcasts it (src/main.cpp:45-51), so that every rank builds the same random initial controls */
  const int drawn_seed = (int)(std::random_device{}() & 0x7fffffff);
  SharedBoot* boot = nullptr;
  std::vector<pid_t> children;
  if (ngpus > 1) {
    boot = (SharedBoot*)mmap(nullptr, sizeof(SharedBoot), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (boot == MAP_FAILED) { perror("mmap"); return 1; }
    boot->id_ready = 0;
    boot->failed = 0;
    fflush(stdout);
    for (int r = 1; r < ngpus; r++) {
      pid_t pid = fork();
      if (pid < 0) { perror("fork"); return 1; }
      if (pid == 0) { mpirank = r; children.clear(); break; }
      children.push_back(pid);
    }
    if (mpirank != 0) quietmode = true;
    static char devbuf[32];
    snprintf(devbuf, sizeof(devbuf), "QB_DEVICE=%d", mpirank);
    putenv(devbuf);
  }

  std::stringstream log;
  Config config(log, quietmode);
  config.ReadFile(cfgfile);

  int rand_seed = config.GetIntParam("rand_seed", -1, false, false);
  if (rand_seed < 0) rand_seed = drawn_seed;
  std::mt19937 rand_engine{};
  rand_engine.seed((unsigned)rand_seed);
  if (mpirank == 0) log << "rand_seed = " << rand_seed << std::endl;     /* the seed actually used (export_param, src/main.cpp:53) */

  std::vector<int> nlevels_int;
  config.GetVecIntParam("nlevels", nlevels_int, 0);
  std::vector<size_t> nlevels(nlevels_int.begin(), nlevels_int.end());
  const int ntime = config.GetIntParam("ntime", 1000);
  const double dt = config.GetDoubleParam("dt", 0.01);
  RunType runtype;
  std::string runtypestr = config.GetStrParam("runtype", "simulation");
  if (runtypestr == "simulation") runtype = RunType::SIMULATION;
  else if (runtypestr == "gradient") runtype = RunType::GRADIENT;
  else if (runtypestr == "optimization") runtype = RunType::OPTIMIZATION;
  else if (runtypestr == "evalcontrols") runtype = RunType::EVALCONTROLS;
  else { printf("\n\n WARNING: Unknown runtype: %s.\n\n", runtypestr.c_str()); runtype = RunType::NONE; }

  std::vector<size_t> nessential(nlevels);
  std::vector<int> read_nessential;
  config.GetVecIntParam("nessential", read_nessential, -1);
  if (read_nessential[0] > -1)
    for (size_t i = 0; i < nlevels.size(); i++) {
      nessential[i] = (size_t)(i < read_nessential.size() ? read_nessential[i] : read_nessential.back());
      if (nessential[i] > nlevels[i]) nessential[i] = nlevels[i];
    }

  /* number of initial conditions (src/main.cpp:90-131) */
  int ninit = 1;
  std::vector<std::string> initcondstr;
  config.GetVecStrParam("initialcondition", initcondstr, "none");
  const std::string collapse = config.GetStrParam("collapse_type", "none");
  if (initcondstr[0] == "file" || initcondstr[0] == "pure" || initcondstr[0] == "performance" || initcondstr[0] == "ensemble") ninit = 1;
  else if (initcondstr[0] == "3states") ninit = 3;
  else if (initcondstr[0] == "Nplus1") { ninit = 1; for (size_t n : nlevels) ninit *= (int)n; ninit += 1; }
  else if (initcondstr[0] == "diagonal" || initcondstr[0] == "basis") {
    ninit = 1;
    if (initcondstr.size() < 2) for (size_t j = 0; j < nlevels.size(); j++) initcondstr.push_back(std::to_string(j));
    for (size_t i = 1; i < initcondstr.size(); i++) { const size_t id = (size_t)atoi(initcondstr[i].c_str()); if (id < nessential.size()) ninit *= (int)nessential[id]; }
    if (initcondstr[0] == "basis" && collapse != "none") ninit = (int)pow(ninit, 2.0);
  } else { printf("\n\n ERROR: Wrong setting for initial condition.\n"); return 1; }
  if (!quietmode) printf("Number of initial conditions: %d\n", ninit);
  if (ninit % ngpus != 0) {
    if (mpirank == 0) printf("\n ERROR: Number of initial conditions (%d) must be divisible by the number of GPUs (%d)!\n", ninit, ngpus);   /* src/main.cpp:150-153 */
    return 1;
  }

  const double total_time = ntime * dt;
  std::vector<double> trans_freq, rot_freq, selfkerr, decay_time, dephase_time;
  config.GetVecDoubleParam("transfreq", trans_freq, 1e20); copyLast(trans_freq, nlevels.size());
  config.GetVecDoubleParam("rotfreq", rot_freq, 1e20); copyLast(rot_freq, nlevels.size());
  config.GetVecDoubleParam("selfkerr", selfkerr, 0.0); copyLast(selfkerr, nlevels.size());
  config.GetVecDoubleParam("decay_time", decay_time, 0.0, true, false); copyLast(decay_time, nlevels.size());
  config.GetVecDoubleParam("dephase_time", dephase_time, 0.0, true, false); copyLast(dephase_time, nlevels.size());
  LindbladType lindbladtype;
  if (collapse == "none") lindbladtype = LindbladType::NONE;
  else if (collapse == "decay") lindbladtype = LindbladType::DECAY;
  else if (collapse == "dephase") lindbladtype = LindbladType::DEPHASE;
  else if (collapse == "both") lindbladtype = LindbladType::BOTH;
  else { printf("\n\n ERROR: Unnown lindblad type: %s.\n", collapse.c_str()); printf(" Choose either 'none', 'decay', 'dephase', or 'both'\n"); return 1; }

  /* oscillators; per-oscillator defaults propagate from the previous one (src/main.cpp:226-250) */
  Oscillator** oscil_vec = new Oscillator*[nlevels.size()];
  std::string default_seg_str = "spline, 10, 0.0, " + std::to_string(total_time);
  std::string default_init_str = "constant, 0.0";
  for (size_t i = 0; i < nlevels.size(); i++) {
    std::vector<double> carrier_freq;
    config.GetVecDoubleParam("carrier_frequency" + std::to_string(i), carrier_freq, 0.0);
    std::vector<std::string> controltype_str, controlinit_str;
    config.GetVecStrParam("control_segments" + std::to_string(i), controltype_str, default_seg_str);
    config.GetVecStrParam("control_initialization" + std::to_string(i), controlinit_str, default_init_str);
    oscil_vec[i] = new Oscillator(config, i, nlevels, controltype_str, controlinit_str, trans_freq[i], selfkerr[i], rot_freq[i], decay_time[i], dephase_time[i],
                                  carrier_freq, total_time, lindbladtype, rand_engine);
    default_seg_str = ""; default_init_str = "";
    for (auto& s : controltype_str) default_seg_str += s + ", ";
    for (auto& s : controlinit_str) default_init_str += s + ", ";
  }
  std::vector<std::string> pipulse_str;
  config.GetVecStrParam("apply_pipulse", pipulse_str, "none", true, false);
  if (pipulse_str[0] != "none") {
    size_t k = 0;
    while (k + 3 < pipulse_str.size()) {
      const size_t id = (size_t)atoi(pipulse_str[k].c_str());
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
  const size_t npairs = (nlevels.size() - 1) * nlevels.size() / 2;
  copyLast(crosskerr, npairs);
  copyLast(Jkl, npairs);
  std::vector<double> eta(npairs);
  size_t idx = 0;
  for (size_t i = 0; i < nlevels.size(); i++)
    for (size_t j = i + 1; j < nlevels.size(); j++) eta[idx++] = rot_freq[i] - rot_freq[j];
  bool usematfree = config.GetBoolParam("usematfree", false);
  const std::string hsys = config.GetStrParam("hamiltonian_file_Hsys", "none", true, false), hc = config.GetStrParam("hamiltonian_file_Hc", "none", true, false);
  if (hsys != "none" || hc != "none") { printf("ERROR: Hamiltonians from file need the sparse-matrix operator, which is out of scope of the B200 engine.\n"); return 1; }
  if (!usematfree) {
    if (!quietmode) printf("# Note: usematfree = false requested; the B200 engine always applies the matrix-free operator (same mathematics).\n");
    usematfree = true;
  }
  MasterEq* mastereq = new MasterEq(nlevels, nessential, oscil_vec, crosskerr, Jkl, eta, lindbladtype, usematfree, quietmode);
  Output* output = new Output(config, (int)nlevels.size(), quietmode, mpirank);

  LinearSolverType linsolvetype;
  const std::string linsolvestr = config.GetStrParam("linearsolver_type", "gmres");
  const int linsolve_maxiter = config.GetIntParam("linearsolver_maxiter", 10);
  if (linsolvestr == "gmres") linsolvetype = LinearSolverType::GMRES;
  else if (linsolvestr == "neumann") linsolvetype = LinearSolverType::NEUMANN;
  else { printf("\n\n ERROR: Unknown linear solver type: %s.\n\n", linsolvestr.c_str()); return 1; }
  bool storeFWD = lindbladtype != LindbladType::NONE && (runtype == RunType::GRADIENT || runtype == RunType::OPTIMIZATION);   /* src/main.cpp:353-355 */
  const std::string ts = config.GetStrParam("timestepper", "IMR");
  TimeStepper* timestepper = nullptr;
  if (ts == "IMR") timestepper = new ImplMidpoint(mastereq, ntime, total_time, linsolvetype, linsolve_maxiter, output, storeFWD, 2, !dryrun);
  else if (ts == "IMR4") timestepper = new CompositionalImplMidpoint(4, mastereq, ntime, total_time, linsolvetype, linsolve_maxiter, output, storeFWD, !dryrun);
  else if (ts == "IMR8") timestepper = new CompositionalImplMidpoint(8, mastereq, ntime, total_time, linsolvetype, linsolve_maxiter, output, storeFWD, !dryrun);
  else if (ts == "EE") timestepper = new ExplEuler(mastereq, ntime, total_time, output, storeFWD, !dryrun);     /* src/main.cpp:362 */
  else { printf("\n\n ERROR: Unknow timestepping type: %s.\n\n", ts.c_str()); return 1; }

  if (ngpus > 1) {   /* join the NCCL communicator of the initial-condition group */
    if (mpirank == 0) {
      if (qb_comm_unique_id(boot->id) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); boot->failed = 1; return reap(children, 1); }
      __sync_synchronize();
      boot->id_ready = 1;
    } else {
      /* wait for rank 0's id; give up if it failed, died (we get re-parented) or stays silent for two minutes */
      const pid_t parent = getppid();
      int waited_ms = 0;
      while (!boot->id_ready) {
        if (boot->failed || getppid() != parent || waited_ms > 120000) { printf("ERROR: rank %d: rank 0 never published the NCCL id\n", mpirank); return 1; }
        usleep(1000);
        waited_ms++;
      }
      __sync_synchronize();
    }
    if (qb_engine_comm_init(timestepper->getEngine(), boot->id, mpirank, ngpus) != QB_OK) {
      printf("ERROR: %s\n", qb_last_error());
      if (mpirank == 0) { boot->failed = 1; return reap(children, 1); }
      return 1;
    }
  }
  OptimProblem* optimctx = new OptimProblem(config, timestepper, ninit, output, quietmode, mpirank, ngpus, ngpus > 1 ? nccl_allreduce : nullptr,
                                            timestepper->getEngine());

  if (mpirank == 0) {  /* config_log.dat: echo of every parsed option (src/main.cpp:382-393) */
    std::ofstream logfile(output->datadir + "/config_log.dat");
    if (logfile.is_open()) logfile << log.str();
  }

  Vec xinit, grad(optimctx->getNdesign(), 0.0);
  if (dryrun) {
    optimctx->getStartingPoint(xinit);
    if (!quietmode) printf("Dry run: %d initial conditions, %d design parameters, state dimension %ld; wrote params.dat and control files.\n", ninit,
                           optimctx->getNdesign(), mastereq->getDim());
    return 0;
  }
  const double StartTime = wtime();
  double objective = 0.0, gnorm = 0.0;
  if (runtype == RunType::SIMULATION) {
    optimctx->getStartingPoint(xinit);
    optimctx->xinit = xinit;
    if (!quietmode) printf("\nStarting primal solver... \n");
    timestepper->writeTrajectoryDataFiles = true;
    objective = optimctx->evalF(xinit);
    if (!quietmode) printf("\nTotal objective = %1.14e, \n", objective);
  }
  if (runtype == RunType::GRADIENT) {
    optimctx->getStartingPoint(xinit);
    optimctx->xinit = xinit;
    if (!quietmode) printf("\nStarting adjoint solver...\n");
    timestepper->writeTrajectoryDataFiles = true;
    optimctx->evalGradF(xinit, grad);
    /* benchmarking aid (not in the reference): QB_REPEAT_EVAL=n repeats the evaluation n more times and reports the mean wall
     * time of the repeats, i.e. the steady-state cost of one optimiser evaluation without first-touch allocations */
    if (const char* rep = getenv("QB_REPEAT_EVAL")) {
      const int nrep = atoi(rep);
      if (nrep > 0) {
        timestepper->writeTrajectoryDataFiles = false;
        const double t0 = wtime();
        for (int i = 0; i < nrep; i++) optimctx->evalGradF(xinit, grad);
        if (mpirank == 0) printf("Evaluation time: %.6f seconds (mean of %d repeated gradient evaluations)\n", (wtime() - t0) / nrep, nrep);
      }
    }
    gnorm = optimctx->getGnorm();
    if (!quietmode) printf("\nGradient norm: %1.14e\n", gnorm);
    output->writeGradient(grad);
  }
  if (runtype == RunType::OPTIMIZATION) {
    optimctx->getStartingPoint(xinit);
    optimctx->xinit = xinit;
    if (!quietmode) printf("\nStarting Optimization solver ... \n");
    timestepper->writeTrajectoryDataFiles = false;
    optimctx->solve(xinit);
  }
  if (runtype == RunType::EVALCONTROLS) {
    optimctx->getStartingPoint(xinit);
    if (!quietmode) printf("\nEvaluating current controls ... \n");
    output->writeControls(xinit, mastereq, ntime, dt);
  }
  if (runtype != RunType::OPTIMIZATION)
    output->writeOptimFile(0, optimctx->getObjective(), gnorm, 0.0, optimctx->getFidelity(), optimctx->getCostT(), optimctx->getRegul(), optimctx->getPenalty(),
                           optimctx->getPenaltyDpDm(), optimctx->getPenaltyEnergy(), optimctx->getPenaltyVariation());

  const double UsedTime = wtime() - StartTime;
  struct rusage r_usage;
  getrusage(RUSAGE_SELF, &r_usage);
  const double myMB = (double)r_usage.ru_maxrss / 1024.0;
  if (!quietmode) {
    printf("\n");
    printf(" Used Time:        %.2f seconds\n", UsedTime);
    printf(" Processors used:  %d (one per GPU)\n", ngpus);
    printf(" Global Memory:    %.2f MB    [~ %.2f MB per proc]\n", myMB, myMB);
    printf("\n");
  }
  if (mpirank == 0) {
    std::string fn = output->datadir + "/timing.dat";
    FILE* timefile = fopen(fn.c_str(), "w");
    if (timefile) { fprintf(timefile, "%d  %1.8e\n", ngpus, UsedTime); fclose(timefile); }
  }
  delete optimctx;
  delete timestepper;
  delete output;
  delete mastereq;
  for (size_t i = 0; i < nlevels.size(); i++) delete oscil_vec[i];
  delete[] oscil_vec;
  int status = 0;
  for (pid_t c : children) { int st = 0; waitpid(c, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) status = 1; }
  return status;
}
