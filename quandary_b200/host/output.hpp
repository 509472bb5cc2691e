/* output.hpp -- the reference's output files (src/output.cpp): optim_history.dat, grad.dat, params.dat, control<k>.dat and the
 * per-initial-condition trajectory files.  Formats are reproduced verbatim: they are the wire format of the reference's
 * regression tests (tests/regression/regression_test.py). */
#pragma once
#include <cstdio>

#include "config.hpp"
#include "mastereq.hpp"

class Output {
 public:
  std::string datadir;
  int output_frequency = 1, optim_monitor_freq = 10;
  /* rank: position in the initial-condition group; optim_history/grad/params/control files are written by rank 0 only, the
   * per-initial-condition trajectory files by the rank that owns the initial condition (as the reference does) */
  Output(const Config& config, int noscillators, bool quietmode_, int rank = 0);
  ~Output();
  void writeOptimFile(int optim_iter, double objective, double gnorm, double stepsize, double Favg, double costT, double tikh_regul, double penalty,
                      double penalty_dpdm, double penalty_energy, double penalty_variation);
  void writeGradient(const Vec& grad);
  void writeControls(const Vec& params, MasterEq* mastereq, int ntime, double dt);
  bool writesTrajectories() const;
  void openTrajectoryDataFiles(const std::string& prefix, int initid);
  void writeTrajectoryDataFiles(int timestep, double time, const Vec& state, MasterEq* mastereq);
  void closeTrajectoryDataFiles();

 private:
  bool quietmode;
  int mpirank = 0;
  FILE* optimfile = nullptr;
  FILE *ufile = nullptr, *vfile = nullptr, *expectedfile_comp = nullptr, *populationfile_comp = nullptr;
  std::vector<FILE*> expectedfile, populationfile;
  std::vector<bool> writeExpectedEnergy, writePopulation;
  bool writeFullState = false, writeExpectedEnergy_comp = false, writePopulation_comp = false;
};
