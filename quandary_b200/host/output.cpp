#include "output.hpp"

#include <sys/stat.h>

#include <cmath>
#include <cstdlib>

Output::Output(const Config& config, int noscillators, bool quietmode_, int rank) {
  quietmode = quietmode_;
  mpirank = rank;
  datadir = config.GetStrParam("datadir", "./data_out");
  mkdir(datadir.c_str(), 0777);
  optim_monitor_freq = config.GetIntParam("optim_monitor_frequency", 10);
  output_frequency = config.GetIntParam("output_frequency", 1);
  char filename[255];
  snprintf(filename, 254, "%s/optim_history.dat", datadir.c_str());
  optimfile = fopen(mpirank == 0 ? filename : "/dev/null", "w");
  if (!optimfile) { printf("ERROR: cannot write to %s\n", filename); exit(1); }
  fprintf(optimfile, "#\"iter\"    \"Objective\"           \"||Pr(grad)||\"           \"LS step\"           \"F_avg\"           \"Terminal cost\"         \"Tikhonov-regul\"        \"Penalty-term\"          \"State variation\"        \"Energy-term\"           \"Control variation\"\n");
  expectedfile.assign(noscillators, nullptr);
  populationfile.assign(noscillators, nullptr);
  writeExpectedEnergy.assign(noscillators, false);
  writePopulation.assign(noscillators, false);
  for (int i = 0; i < noscillators; i++) {
    std::vector<std::string> strs;
    config.GetVecStrParam("output" + std::to_string(i), strs, "none");
    for (const auto& s : strs) {
      if (s == "expectedEnergy") writeExpectedEnergy[i] = true;
      if (s == "expectedEnergyComposite") writeExpectedEnergy_comp = true;
      if (s == "population") writePopulation[i] = true;
      if (s == "populationComposite") writePopulation_comp = true;
      if (s == "fullstate") writeFullState = true;
    }
  }
}

Output::~Output() {
  if (!quietmode && mpirank == 0) printf("Output directory: %s\n", datadir.c_str());
  if (optimfile) fclose(optimfile);
}

bool Output::writesTrajectories() const {
  bool any = writeFullState || writeExpectedEnergy_comp || writePopulation_comp;
  for (bool b : writeExpectedEnergy) any = any || b;
  for (bool b : writePopulation) any = any || b;
  return any;
}

void Output::writeOptimFile(int optim_iter, double objective, double gnorm, double stepsize, double Favg, double costT, double tikh_regul, double penalty,
                            double penalty_dpdm, double penalty_energy, double penalty_variation) {
  fprintf(optimfile, "%05d  %1.14e  %1.14e  %.8f  %1.14e  %1.14e  %1.14e  %1.14e  %1.14e  %1.14e  %1.14e\n", optim_iter, objective, gnorm, stepsize, Favg, costT,
          tikh_regul, penalty, penalty_dpdm, penalty_energy, penalty_variation);
  fflush(optimfile);
}

void Output::writeGradient(const Vec& grad) {
  if (mpirank != 0) return;
  char filename[255];
  snprintf(filename, 254, "%s/grad.dat", datadir.c_str());
  FILE* file = fopen(filename, "w");
  for (double g : grad) fprintf(file, "%1.14e\n", g);
  fclose(file);
}

void Output::writeControls(const Vec& params, MasterEq* mastereq, int ntime, double dt) {
  if (mpirank != 0) { mastereq->setControlAmplitudes(params); return; }
  char filename[255];
  snprintf(filename, 254, "%s/params.dat", datadir.c_str());
  FILE* file = fopen(filename, "w");
  for (double p : params) fprintf(file, "%1.14e\n", p);
  fclose(file);
  if (!quietmode) printf("File written: %s\n", filename);
  mastereq->setControlAmplitudes(params);
  /* p(t), q(t) of every oscillator at the output times (= Oscillator::evalControl): evaluated on the device, by the kernel
   * that feeds the sweeps; without an engine (--dry-run) by the host restatement */
  std::vector<double> times;
  for (int i = 0; i <= ntime; i += output_frequency) times.push_back(i * dt);
  const size_t Q = mastereq->getNOscillators();
  std::vector<double> p(times.size() * Q), q(times.size() * Q);
  if (mastereq->getEngine()) {
    if (qb_engine_eval_controls(mastereq->getEngine(), times.data(), (int)times.size(), p.data(), q.data()) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
  } else {
    for (size_t i = 0; i < times.size(); i++)
      for (size_t k = 0; k < Q; k++) mastereq->getOscillator(k)->evalControl(times[i], &p[i * Q + k], &q[i * Q + k]);
  }
  for (size_t k = 0; k < Q; k++) {
    snprintf(filename, 254, "%s/control%zu.dat", datadir.c_str(), k);
    FILE* fc = fopen(filename, "w");
    fprintf(fc, "#\"time\"         \"p(t) (rotating)\"          \"q(t) (rotating)\"         \"f(t) (labframe)\"\n");
    const double gf = mastereq->getOscillator(k)->ground_freq;
    for (size_t i = 0; i < times.size(); i++) {
      const double t = times[i], ReI = p[i * Q + k], ImI = q[i * Q + k];
      const double LabI = 2.0 * ReI * cos(gf * t) - 2.0 * ImI * sin(gf * t);      /* evalControl_Labframe, src/oscillator.cpp:383-428 */
      fprintf(fc, "% 1.8f   % 1.14e   % 1.14e   % 1.14e \n", t, ReI / (2.0 * M_PI), ImI / (2.0 * M_PI), LabI / (2.0 * M_PI));
    }
    fclose(fc);
    if (!quietmode) printf("File written: %s\n", filename);
  }
}

void Output::openTrajectoryDataFiles(const std::string& prefix, int initid) {
  char filename[255];
  for (size_t i = 0; i < expectedfile.size(); i++)
    if (writeExpectedEnergy[i]) {
      snprintf(filename, 254, "%s/expected%zu.iinit%04d.dat", datadir.c_str(), i, initid);
      expectedfile[i] = fopen(filename, "w");
      fprintf(expectedfile[i], "#\"time\"      \"expected energy level\"\n");
    }
  if (writeExpectedEnergy_comp) {
    snprintf(filename, 254, "%s/expected_composite.iinit%04d.dat", datadir.c_str(), initid);
    expectedfile_comp = fopen(filename, "w");
    fprintf(expectedfile_comp, "#\"time\"      \"expected energy level\"\n");
  }
  for (size_t i = 0; i < populationfile.size(); i++)
    if (writePopulation[i]) {
      snprintf(filename, 254, "%s/population%zu.iinit%04d.dat", datadir.c_str(), i, initid);
      populationfile[i] = fopen(filename, "w");
      fprintf(populationfile[i], "#\"time\"      \"diagonal of the density matrix\"\n");
    }
  if (writePopulation_comp) {
    snprintf(filename, 254, "%s/population_composite.iinit%04d.dat", datadir.c_str(), initid);
    populationfile_comp = fopen(filename, "w");
    fprintf(populationfile_comp, "#\"time\"      \"population\"\n");
  }
  if (writeFullState) {
    snprintf(filename, 254, "%s/%s_Re.iinit%04d.dat", datadir.c_str(), prefix.c_str(), initid);
    ufile = fopen(filename, "w");
    snprintf(filename, 254, "%s/%s_Im.iinit%04d.dat", datadir.c_str(), prefix.c_str(), initid);
    vfile = fopen(filename, "w");
  }
}

void Output::writeTrajectoryDataFiles(int timestep, double time, const Vec& state, MasterEq* mastereq) {
  if (timestep % output_frequency != 0) return;
  for (size_t k = 0; k < expectedfile.size(); k++)
    if (writeExpectedEnergy[k]) fprintf(expectedfile[k], "%.8f %1.14e\n", time, mastereq->getOscillator(k)->expectedEnergy(state));
  if (writeExpectedEnergy_comp) fprintf(expectedfile_comp, "%.8f %1.14e\n", time, mastereq->expectedEnergy(state));
  for (size_t k = 0; k < populationfile.size(); k++)
    if (writePopulation[k]) {
      std::vector<double> pop;
      mastereq->getOscillator(k)->population(state, pop);
      fprintf(populationfile[k], "%.8f ", time);
      for (double v : pop) fprintf(populationfile[k], " %1.14e", v);
      fprintf(populationfile[k], "\n");
    }
  if (writePopulation_comp) {
    std::vector<double> pop;
    mastereq->population(state, pop);
    fprintf(populationfile_comp, "%.8f  ", time);
    for (double v : pop) fprintf(populationfile_comp, "%1.14e  ", v);
    fprintf(populationfile_comp, "\n");
  }
  if (writeFullState) {
    fprintf(ufile, "%.8f  ", time);
    fprintf(vfile, "%.8f  ", time);
    const long dim = mastereq->getDim();
    for (long i = 0; i < dim; i++) { fprintf(ufile, "%1.10e  ", state[i]); fprintf(vfile, "%1.10e  ", state[i + dim]); }
    fprintf(ufile, "\n");
    fprintf(vfile, "\n");
  }
}

void Output::closeTrajectoryDataFiles() {
  if (ufile) { fclose(ufile); ufile = nullptr; }
  if (vfile) { fclose(vfile); vfile = nullptr; }
  for (auto& f : expectedfile) if (f) { fclose(f); f = nullptr; }
  if (expectedfile_comp) { fclose(expectedfile_comp); expectedfile_comp = nullptr; }
  for (auto& f : populationfile) if (f) { fclose(f); f = nullptr; }
  if (populationfile_comp) { fclose(populationfile_comp); populationfile_comp = nullptr; }
}
