/* mastereq.hpp -- host-side mirror of the reference's MasterEq (include/mastereq.hpp:91-263): holds the system description
 * and forwards the operator calls to the device engine.  The matrix-free operator, assemble_RHS and compute_dRHS_dParams
 * run inside the engine's sweep kernels; the single-call forms below exist for tests and outputs. */
#pragma once
#include "../../include/quandary_b200.h"
#include "oscillator.hpp"

class MasterEq {
 public:
  std::vector<size_t> nlevels, nessential;
  bool usematfree = true;
  LindbladType lindbladtype = LindbladType::NONE;
  std::vector<double> crosskerr, Jkl, eta;      /* angular, pair order (0,1),(0,2),...,(1,2),... (src/mastereq.cpp:29-37) */

  MasterEq(const std::vector<size_t>& nlevels_, const std::vector<size_t>& nessential_, Oscillator** oscil_vec_, const std::vector<double>& crosskerr_,
           const std::vector<double>& Jkl_, const std::vector<double>& eta_, LindbladType lindbladtype_, bool usematfree_, bool quietmode_);
  ~MasterEq() {}

  size_t getNOscillators() const { return noscillators; }
  Oscillator* getOscillator(size_t i) const { return oscil_vec[i]; }
  long getDim() const { return dim; }          /* N^2 (Lindblad) or N */
  long getDimRho() const { return dim_rho; }   /* N */
  long getDimEss() const { return dim_ess; }

  /* fills everything of the engine descriptor that MasterEq and the oscillators own; the backing arrays live in this object */
  void fillDescriptor(qb_system_desc& d);
  /* the engine is created by the time stepper (it needs ntime and the linear solver) and attached here */
  void attachEngine(qb_engine* e) { engine = e; }
  qb_engine* getEngine() const { return engine; }

  /* = MasterEq::setControlAmplitudes (src/mastereq.cpp:693-707): oscillators concatenated */
  void setControlAmplitudes(const Vec& x);
  /* = assemble_RHS(t) + MatMult / MatMultTranspose on the shell matrix (src/mastereq.cpp:657-678,710-739) */
  void applyRHS(double t, const Vec& x, Vec& y, bool transpose = false);
  /* <N> and populations of the composite system (src/mastereq.cpp:2896-2975) */
  double expectedEnergy(const Vec& x) const;
  void population(const Vec& x, std::vector<double>& pop) const;

 private:
  size_t noscillators;
  Oscillator** oscil_vec;
  long dim, dim_rho, dim_ess;
  bool quietmode;
  qb_engine* engine = nullptr;
  std::vector<qb_segment> segs_;
  std::vector<double> carriers_, pipulses_;
};
