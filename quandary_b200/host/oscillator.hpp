/* oscillator.hpp -- per-subsystem constants and control parameterisation; mirrors the reference's Oscillator
 * (include/oscillator.hpp, src/oscillator.cpp).  evalControl / evalControl_diff themselves run on the device inside
 * the engine (control table); this class holds the parameters, parses the .cfg control options exactly as the reference
 * does and serves the host-side outputs (expectedEnergy, population). */
#pragma once
#include <random>

#include "config.hpp"
#include "defs.hpp"

struct ControlBasis {            /* one control segment = the reference's ControlBasis subclass instance */
  ControlType type;
  int nsplines;
  double tstart, tstop;
  int skip;
  bool enforceZeroBoundary;
  int getNparams() const { return 2 * nsplines; }     /* BSpline2nd / BSpline0: 2*nsplines per carrier */
};

struct PiPulse { std::vector<double> tstart, tstop, amp; };

class Oscillator {
 public:
  size_t myid = 0;
  size_t nlevels = 0;
  double Tfinal = 0.0;
  double ground_freq = 0.0, selfkerr = 0.0, detuning_freq = 0.0;   /* angular */
  double decay_time = 0.0, dephase_time = 0.0;
  std::vector<double> carrier_freq;                                /* angular */
  std::vector<ControlBasis> basisfunctions;
  std::vector<double> params;
  bool control_enforceBC = true;
  LindbladType lindbladtype = LindbladType::NONE;
  long dim_preOsc = 1, dim_postOsc = 1;
  PiPulse pipulse;

  Oscillator() {}
  /* rand_engine is taken BY VALUE as in the reference (include/oscillator.hpp:99): every oscillator draws the same sequence */
  Oscillator(const Config& config, size_t id, const std::vector<size_t>& nlevels_all, std::vector<std::string>& controlsegments,
             std::vector<std::string>& controlinitializations, double ground_freq_, double selfkerr_, double rotational_freq_, double decay_time_,
             double dephase_time_, const std::vector<double>& carrier_freq_, double Tfinal_, LindbladType lindbladtype_, std::mt19937 rand_engine);

  size_t getNParams() const { return params.size(); }
  size_t getNLevels() const { return nlevels; }
  size_t getNSegments() const { return basisfunctions.size(); }
  size_t getNCarrierfrequencies() const { return carrier_freq.size(); }
  int getNSegParams(int segmentID) const;
  double getSelfkerr() const { return selfkerr; }
  double getDetuning() const { return detuning_freq; }
  double getDecayTime() const { return decay_time; }
  double getDephaseTime() const { return dephase_time; }
  void setParams(const double* x);
  void getParams(double* x) const;
  /* p(t), q(t) and their derivative wrt the parameters on the host (src/oscillator.cpp:281-381, src/controlbasis.cpp:20-96,
   * 219-255); the sweeps evaluate the same functions on the device -- these serve the state-independent energy penalty. */
  int evalControl(double t, double* Re_ptr, double* Im_ptr) const;
  int evalControl_diff(double t, double* grad, double pbar, double qbar) const;
  double evalControlVariation() const;
  void evalControlVariationDiff(Vec& G, double var_reg_bar, int skip_to_oscillator) const;
  /* <N_k> and the level populations of this subsystem for a state in the reference layout (src/oscillator.cpp:430-566) */
  double expectedEnergy(const Vec& x) const;
  void population(const Vec& x, std::vector<double>& pop) const;
};
