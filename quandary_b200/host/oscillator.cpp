#include "oscillator.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

Oscillator::Oscillator(const Config& config, size_t id, const std::vector<size_t>& nlevels_all, std::vector<std::string>& controlsegments,
                       std::vector<std::string>& controlinitializations, double ground_freq_, double selfkerr_, double rotational_freq_,
                       double decay_time_, double dephase_time_, const std::vector<double>& carrier_freq_, double Tfinal_, LindbladType lindbladtype_,
                       std::mt19937 rand_engine) {
  myid = id;
  nlevels = nlevels_all[id];
  Tfinal = Tfinal_;
  ground_freq = ground_freq_ * 2.0 * M_PI;            /* src/oscillator.cpp:15-21 */
  selfkerr = selfkerr_ * 2.0 * M_PI;
  detuning_freq = 2.0 * M_PI * (ground_freq_ - rotational_freq_);
  carrier_freq = carrier_freq_;
  for (auto& f : carrier_freq) f *= 2.0 * M_PI;
  decay_time = decay_time_;
  dephase_time = dephase_time_;
  lindbladtype = lindbladtype_;
  control_enforceBC = config.GetBoolParam("control_enforceBC", true);

  /* control segments: "spline, n [, tstart, tstop]" | "spline0, n [, tstart, tstop]" (src/oscillator.cpp:45-140) */
  size_t idstr = 0;
  int nparams_per_seg = 0;
  while (idstr < controlsegments.size()) {
    const std::string& kind = controlsegments[idstr];
    if (kind == "spline" || kind == "spline0") {
      idstr++;
      if (controlsegments.size() <= idstr) { printf("ERROR: Wrong setting for control segments: Number of splines not found.\n"); exit(1); }
      ControlBasis b;
      b.type = kind == "spline" ? ControlType::BSPLINE : ControlType::BSPLINE0;
      b.nsplines = atoi(controlsegments[idstr].c_str()); idstr++;
      b.tstart = 0.0; b.tstop = Tfinal;
      if (controlsegments.size() >= idstr + 2) { b.tstart = atof(controlsegments[idstr].c_str()); idstr++; b.tstop = atof(controlsegments[idstr].c_str()); idstr++; }
      b.enforceZeroBoundary = control_enforceBC;
      b.skip = nparams_per_seg;
      nparams_per_seg += b.getNparams() * (int)carrier_freq.size();
      basisfunctions.push_back(b);
    } else if (kind == "step" || kind == "spline_amplitude") {
      printf("ERROR: control segment type '%s' has no adjoint on the B200 engine (the reference has none for spline_amplitude either, src/oscillator.cpp:350-355).\n", kind.c_str());
      exit(1);
    } else {
      idstr++;   /* non-controllable */
    }
  }

  /* initial parameters: "constant, amp" | "random, amp" | "file, name" per segment; amplitudes x 2*pi (src/oscillator.cpp:142-198) */
  size_t idini = 0;
  for (size_t seg = 0; seg < basisfunctions.size(); seg++) {
    if (controlinitializations.size() < idini + 2) { controlinitializations.push_back("constant"); controlinitializations.push_back("0.0"); }
    const double initval = atof(controlinitializations[idini + 1].c_str()) * 2.0 * M_PI;
    const size_t n = (size_t)basisfunctions[seg].getNparams();
    if (controlinitializations[idini] == "constant") {
      for (size_t f = 0; f < carrier_freq.size(); f++)
        for (size_t i = 0; i < n; i++) params.push_back(initval);
    } else if (controlinitializations[idini] == "random") {
      std::uniform_real_distribution<double> unit_dist(0.0, 1.0);
      for (size_t f = 0; f < carrier_freq.size(); f++)
        for (size_t i = 0; i < n; i++) {
          const double randval = unit_dist(rand_engine);
          double val = initval * randval;
          val = 2 * val - initval;
          params.push_back(val);
        }
    } else {
      for (size_t i = 0; i < n * carrier_freq.size(); i++) params.push_back(0.0);
    }
    idini += 2;
  }
  /* boundary conditions on the initial guess (BSpline2nd::enforceBoundary / BSpline0::enforceBoundary) */
  if (!params.empty() && control_enforceBC) {
    for (const auto& b : basisfunctions)
      for (size_t f = 0; f < carrier_freq.size(); f++) {
        const int ns = b.nsplines;
        if (b.type == ControlType::BSPLINE) {
          for (int l = 0; l < ns; l++)
            if (l <= 1 || l >= ns - 2) { params[b.skip + f * ns * 2 + l] = 0.0; params[b.skip + f * ns * 2 + l + ns] = 0.0; }
        } else {
          params[b.skip + 2 * f * ns] = 0.0; params[b.skip + 2 * f * ns + ns - 1] = 0.0;
          params[b.skip + (2 * f + 1) * ns] = 0.0; params[b.skip + (2 * f + 1) * ns + ns - 1] = 0.0;
        }
      }
  }
  for (size_t j = 0; j < nlevels_all.size(); j++) {
    if (j < id) dim_preOsc *= (long)nlevels_all[j];
    if (j > id) dim_postOsc *= (long)nlevels_all[j];
  }
}

int Oscillator::getNSegParams(int segmentID) const {
  if (params.empty()) return 0;
  return basisfunctions[segmentID].getNparams() * (int)carrier_freq.size();
}
void Oscillator::setParams(const double* x) { for (size_t i = 0; i < params.size(); i++) params[i] = x[i]; }
void Oscillator::getParams(double* x) const { for (size_t i = 0; i < params.size(); i++) x[i] = params[i]; }

static inline double SQR(double x) { return x * x; }

/* BSpline0::computeVariation (src/controlbasis.cpp:260-280); BSpline2nd has none (returns 0) */
double Oscillator::evalControlVariation() const {
  double var = 0.0;
  if (params.empty()) return var;
  for (const auto& b : basisfunctions) {
    if (b.type != ControlType::BSPLINE0) continue;
    const int ns = b.nsplines;
    for (size_t f = 0; f < carrier_freq.size(); f++) {
      for (int lc = 1; lc < ns; lc++) var += SQR(params[b.skip + 2 * f * ns + lc] - params[b.skip + 2 * f * ns + lc - 1]);
      for (int lc = 1; lc < ns; lc++) var += SQR(params[b.skip + (2 * f + 1) * ns + lc] - params[b.skip + (2 * f + 1) * ns + lc - 1]);
      if (b.enforceZeroBoundary) {
        var += SQR(params[b.skip + 2 * f * ns]); var += SQR(params[b.skip + 2 * f * ns + ns - 1]);
        var += SQR(params[b.skip + (2 * f + 1) * ns]); var += SQR(params[b.skip + (2 * f + 1) * ns + ns - 1]);
      }
    }
  }
  return var;
}
void Oscillator::evalControlVariationDiff(Vec& G, double var_bar, int skip_to_oscillator) const {
  if (params.empty()) return;
  double* grad = G.data() + skip_to_oscillator;
  const double fact = 2.0 * var_bar;
  for (const auto& b : basisfunctions) {
    if (b.type != ControlType::BSPLINE0) continue;
    const int ns = b.nsplines;
    for (size_t f = 0; f < carrier_freq.size(); f++) {
      for (int part = 0; part < 2; part++) {
        const int o = b.skip + (2 * (int)f + part) * ns;
        grad[o] += fact * (params[o] - params[o + 1]);
        for (int lc = 1; lc < ns - 1; lc++) grad[o + lc] += fact * (2 * params[o + lc] - params[o + lc - 1] - params[o + lc + 1]);
        grad[o + ns - 1] += fact * (params[o + ns - 1] - params[o + ns - 2]);
        if (b.enforceZeroBoundary) { grad[o] += fact * params[o]; grad[o + ns - 1] += fact * params[o + ns - 1]; }
      }
    }
  }
}

double Oscillator::expectedEnergy(const Vec& x) const {
  const long dim = (long)x.size() / 2;
  const bool lind = lindbladtype != LindbladType::NONE;
  const long dimmat = lind ? (long)std::llround(std::sqrt((double)dim)) : dim;
  double expected = 0.0;
  for (long i = 0; i < dimmat; i++) {
    long num_diag = (i % ((long)nlevels * dim_postOsc)) / dim_postOsc;
    if (lind) expected += num_diag * x[getVecID(i, i, dimmat)];
    else expected += num_diag * x[i] * x[i] + num_diag * x[i + dim] * x[i + dim];
  }
  return expected;
}
void Oscillator::population(const Vec& x, std::vector<double>& pop) const {
  const long dim = (long)x.size() / 2;
  const bool lind = lindbladtype != LindbladType::NONE;
  const long dimN = lind ? (long)std::llround(std::sqrt((double)dim)) : dim;
  pop.assign(nlevels, 0.0);
  for (long i = 0; i < dimN; i++) {
    const long lev = (i % ((long)nlevels * dim_postOsc)) / dim_postOsc;
    if (lind) pop[lev] += x[getVecID(i, i, dimN)];
    else pop[lev] += x[i] * x[i] + x[i + dim] * x[i + dim];
  }
}

/* BSpline2nd::basisfunction, src/controlbasis.cpp:81-96 */
static double bspline2_basis(const ControlBasis& b, int id, double t) {
  const double dtknot = (b.tstop - b.tstart) / (double)(b.nsplines - 2);
  const double width = 3.0 * dtknot;
  const double tcenter = b.tstart + dtknot * ((id + 1) - 1.5);
  const double tau = (t - tcenter) / width;
  if (tau < -1. / 2. || tau >= 1. / 2.) return 0.0;
  double val = 0.0;
  if (-1. / 2. <= tau && tau < -1. / 6.) val = 9. / 8. + 9. / 2. * tau + 9. / 2. * pow(tau, 2);
  else if (-1. / 6. <= tau && tau < 1. / 6.) val = 3. / 4. - 9. * pow(tau, 2);
  else if (1. / 6. <= tau && tau < 1. / 2.) val = 9. / 8. - 9. / 2. * tau + 9. / 2. * pow(tau, 2);
  return val;
}

int Oscillator::evalControl(double t, double* Re_ptr, double* Im_ptr) const {
  if (t > Tfinal) { printf("ERROR: accessing spline outside of [0,T] at %f. Should never happen! Bug.\n", t); exit(1); }
  *Re_ptr = 0.0; *Im_ptr = 0.0;
  if (!params.empty()) {
    for (const auto& b : basisfunctions) {
      if (b.tstart <= t && b.tstop >= t) {
        double sum_p = 0.0, sum_q = 0.0;
        const int ns = b.nsplines;
        for (size_t f = 0; f < carrier_freq.size(); f++) {
          double B1 = 0.0, B2 = 0.0;
          if (b.type == ControlType::BSPLINE) {
            for (int l = 0; l < ns; l++) {
              if (b.enforceZeroBoundary && (l <= 1 || l >= ns - 2)) continue;
              const double Blt = bspline2_basis(b, l, t);
              B1 += params[b.skip + f * ns * 2 + l] * Blt;
              B2 += params[b.skip + f * ns * 2 + l + ns] * Blt;
            }
          } else {
            const double dtknot = (b.tstop - b.tstart) / (ns - 1.0);
            const int id = (int)ceil((t - b.tstart) / dtknot - 0.5);
            if (id >= 0 && id < ns) { B1 = params[b.skip + f * ns * 2 + id]; B2 = params[b.skip + f * ns * 2 + id + ns]; }
          }
          const double cos_omt = cos(carrier_freq[f] * t), sin_omt = sin(carrier_freq[f] * t);
          sum_p += cos_omt * B1 - sin_omt * B2;
          sum_q += sin_omt * B1 + cos_omt * B2;
        }
        *Re_ptr = sum_p; *Im_ptr = sum_q;
        break;
      }
    }
  }
  for (size_t i = 0; i < pipulse.tstart.size(); i++)
    if (pipulse.tstart[i] <= t && t <= pipulse.tstop[i]) { *Re_ptr = pipulse.amp[i] / sqrt(2.0); *Im_ptr = pipulse.amp[i] / sqrt(2.0); }
  return 0;
}

int Oscillator::evalControl_diff(double t, double* grad, double pbar, double qbar) const {
  if (!params.empty()) {
    for (const auto& b : basisfunctions) {
      if (b.tstart <= t && b.tstop >= t) {
        const int ns = b.nsplines;
        for (size_t f = 0; f < carrier_freq.size(); f++) {
          const double cos_omt = cos(carrier_freq[f] * t), sin_omt = sin(carrier_freq[f] * t);
          const double B1bar = sin_omt * qbar + cos_omt * pbar, B2bar = cos_omt * qbar - sin_omt * pbar;
          if (b.type == ControlType::BSPLINE) {
            for (int l = 0; l < ns; l++) {
              if (b.enforceZeroBoundary && (l <= 1 || l >= ns - 2)) continue;
              const double Blt = bspline2_basis(b, l, t);
              grad[b.skip + f * ns * 2 + l] += Blt * B1bar;
              grad[b.skip + f * ns * 2 + l + ns] += Blt * B2bar;
            }
          } else {
            const double dtknot = (b.tstop - b.tstart) / (ns - 1.0);
            const int id = (int)ceil((t - b.tstart) / dtknot - 0.5);
            if (id >= 0 && id < ns) { grad[b.skip + f * ns * 2 + id] += B1bar; grad[b.skip + f * ns * 2 + id + ns] += B2bar; }
          }
        }
        break;
      }
    }
  }
  for (size_t i = 0; i < pipulse.tstart.size(); i++)
    if (pipulse.tstart[i] <= t && t <= pipulse.tstop[i]) { printf("ERROR: Derivative of pipulse not implemented. Sorry! But also, this should never happen!\n"); exit(1); }
  return 0;
}
