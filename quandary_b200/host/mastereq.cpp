#include "mastereq.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

MasterEq::MasterEq(const std::vector<size_t>& nlevels_, const std::vector<size_t>& nessential_, Oscillator** oscil_vec_, const std::vector<double>& crosskerr_,
                   const std::vector<double>& Jkl_, const std::vector<double>& eta_, LindbladType lindbladtype_, bool usematfree_, bool quietmode_) {
  nlevels = nlevels_;
  nessential = nessential_;
  noscillators = nlevels.size();
  oscil_vec = oscil_vec_;
  crosskerr = crosskerr_;
  Jkl = Jkl_;
  eta = eta_;
  lindbladtype = lindbladtype_;
  usematfree = usematfree_;
  quietmode = quietmode_;
  for (auto& v : crosskerr) v *= 2.0 * M_PI;     /* src/mastereq.cpp:29-37 */
  for (auto& v : Jkl) v *= 2.0 * M_PI;
  for (auto& v : eta) v *= 2.0 * M_PI;
  dim_rho = 1; dim_ess = 1;
  for (size_t i = 0; i < noscillators; i++) { dim_rho *= (long)nlevels[i]; dim_ess *= (long)nessential[i]; }
  dim = lindbladtype == LindbladType::NONE ? dim_rho : dim_rho * dim_rho;
  if (!usematfree) {
    printf("ERROR: the B200 engine implements the matrix-free operator only (usematfree = true); the sparse-matrix path is out of scope.\n");
    exit(1);
  }
  if (noscillators > QB_MAX_OSC) {
    printf("ERROR: matrix-free solver supports at most %d oscillators (as the reference, src/mastereq.cpp:722-737).\n", QB_MAX_OSC);
    exit(1);
  }
}

void MasterEq::fillDescriptor(qb_system_desc& d) {
  memset(&d, 0, sizeof(d));
  const size_t Q = noscillators;
  d.nosc = (int)Q;
  d.lindblad = (int)lindbladtype;
  segs_.clear(); carriers_.clear(); pipulses_.clear();
  for (size_t k = 0; k < Q; k++) {
    const Oscillator* o = oscil_vec[k];
    d.nlevels[k] = (int)o->getNLevels();
    d.detuning[k] = o->getDetuning();
    d.selfkerr[k] = o->getSelfkerr();
    d.decay_time[k] = o->getDecayTime();
    d.dephase_time[k] = o->getDephaseTime();
    d.seg_off[k] = (int)segs_.size();
    d.carrier_off[k] = (int)carriers_.size();
    d.pipulse_off[k] = (int)pipulses_.size() / 3;
    if (o->getNParams() > 0)
      for (const auto& b : o->basisfunctions) {
        qb_segment s;
        s.type = b.type == ControlType::BSPLINE0 ? QB_CTRL_BSPLINE0 : QB_CTRL_BSPLINE2ND;
        s.nsplines = b.nsplines; s.tstart = b.tstart; s.tstop = b.tstop; s.skip = b.skip; s.enforce_bc = b.enforceZeroBoundary ? 1 : 0;
        segs_.push_back(s);
      }
    for (double f : o->carrier_freq) carriers_.push_back(f);
    for (size_t i = 0; i < o->pipulse.tstart.size(); i++) { pipulses_.push_back(o->pipulse.tstart[i]); pipulses_.push_back(o->pipulse.tstop[i]); pipulses_.push_back(o->pipulse.amp[i]); }
  }
  d.seg_off[Q] = (int)segs_.size();
  d.carrier_off[Q] = (int)carriers_.size();
  d.pipulse_off[Q] = (int)pipulses_.size() / 3;
  d.segments = segs_.data();
  d.carrier_freq = carriers_.data();
  d.pipulse = pipulses_.data();
  for (size_t i = 0; i < Q * (Q - 1) / 2; i++) {
    d.crosskerr[i] = i < crosskerr.size() ? crosskerr[i] : 0.0;
    d.Jkl[i] = i < Jkl.size() ? Jkl[i] : 0.0;
    d.eta[i] = i < eta.size() ? eta[i] : 0.0;
  }
}

void MasterEq::setControlAmplitudes(const Vec& x) {
  size_t shift = 0;
  for (size_t k = 0; k < noscillators; k++) {
    oscil_vec[k]->setParams(x.data() + shift);
    shift += oscil_vec[k]->getNParams();
  }
  if (engine && qb_engine_set_params(engine, x.data(), (int)x.size()) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
}

void MasterEq::applyRHS(double t, const Vec& x, Vec& y, bool transpose) {
  y.resize(x.size());
  if (qb_engine_rhs_apply(engine, t, transpose ? 1 : 0, x.data(), y.data(), 1) != QB_OK) { printf("ERROR: %s\n", qb_last_error()); exit(1); }
}

double MasterEq::expectedEnergy(const Vec& x) const {
  const bool lind = lindbladtype != LindbladType::NONE;
  const long d2 = (long)x.size() / 2;
  double expected = 0.0;
  for (long i = 0; i < dim_rho; i++) {
    const double xd = lind ? x[getVecID(i, i, dim_rho)] : x[i] * x[i] + x[i + d2] * x[i + d2];
    expected += i * xd;
  }
  return expected;
}
void MasterEq::population(const Vec& x, std::vector<double>& pop) const {
  const bool lind = lindbladtype != LindbladType::NONE;
  const long d2 = (long)x.size() / 2;
  pop.assign(dim_rho, 0.0);
  for (long i = 0; i < dim_rho; i++) pop[i] = lind ? x[getVecID(i, i, dim_rho)] : x[i] * x[i] + x[i + d2] * x[i + d2];
}
