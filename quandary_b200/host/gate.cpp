#include "gate.hpp"

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>

typedef std::complex<double> cplx;

Gate::Gate(const std::string& kind, const std::string& filename, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential, double time,
           const std::vector<double>& gate_rot_freq_, LindbladType lindbladtype_, bool quietmode) {
  lindbladtype = lindbladtype_;
  if (kind == "none") return;      /* dummy gate */
  dim_ess = 1; dim_rho = 1;
  for (size_t i = 0; i < nessential.size(); i++) { dim_ess *= (long)nessential[i]; dim_rho *= (long)nlevels[i]; }
  const long n = dim_ess;
  std::vector<cplx> Ve((size_t)n * n, cplx(0.0, 0.0));
  auto at = [&](long r, long c) -> cplx& { return Ve[(size_t)r * n + c]; };
  if (kind == "xgate") { assert(n == 2); at(0, 1) = 1.0; at(1, 0) = 1.0; }
  else if (kind == "ygate") { assert(n == 2); at(0, 1) = cplx(0, -1); at(1, 0) = cplx(0, 1); }
  else if (kind == "zgate") { assert(n == 2); at(0, 0) = cplx(0, 1); at(1, 1) = cplx(0, -1); }   /* the reference fills V_im here (src/gate.cpp:330-331) */
  else if (kind == "hadamard") { assert(n == 2); const double v = 1. / sqrt(2); at(0, 0) = v; at(0, 1) = v; at(1, 0) = v; at(1, 1) = -v; }
  else if (kind == "cnot") { assert(n == 4); at(0, 0) = 1.0; at(1, 1) = 1.0; at(2, 3) = 1.0; at(3, 2) = 1.0; }
  else if (kind == "swap") { assert(n == 4); at(0, 0) = 1.0; at(1, 2) = 1.0; at(2, 1) = 1.0; at(3, 3) = 1.0; }
  else if (kind == "swap0q") {
    const int Q = (int)nlevels.size();
    const long h = 1L << (Q - 2), f = 1L << (Q - 1);
    for (long i = 0; i < h; i++) at(2 * i, 2 * i) = 1.0;
    for (long i = h; i < f; i++) at(2 * i + 1, 2 * i + 1) = 1.0;
    for (long i = 0; i < h; i++) { at(2 * i + 1, 2 * i + f) = 1.0; at(2 * i + f, 2 * i + 1) = 1.0; }
  } else if (kind == "cqnot") {
    for (long k = 0; k < n - 2; k++) at(k, k) = 1.0;
    at(n - 2, n - 1) = 1.0; at(n - 1, n - 2) = 1.0;
  } else if (kind == "qft") {
    const double sq = sqrt((double)n);
    for (long j = 0; j < n; j++)
      for (long k = 0; k < n; k++) at(j, k) = cplx(cos(2.0 * M_PI * j * k / n) / sq, sin(2.0 * M_PI * j * k / n) / sq);
  } else if (kind == "file") {
    std::vector<double> vec(2 * (size_t)n * n);
    read_vector(filename.c_str(), vec.data(), (int)vec.size(), quietmode);
    for (long i = 0; i < n * n; i++) at(i % n, i / n) = cplx(vec[i], vec[i + n * n]);
  } else {
    printf("ERROR. Could not find target gate. Exiting now.\n");
    exit(1);
  }
  /* rotating frame: row r scaled by exp(i*freq_r*T), freq_r = sum_k r_k * 2*pi*gate_rot_freq_k (src/gate.cpp:88-131) */
  for (long row = 0; row < n; row++) {
    long r = row;
    double freq = 0.0;
    for (size_t iosc = 0; iosc < nlevels.size(); iosc++) {
      long dim_post = 1;
      for (size_t josc = iosc + 1; josc < nlevels.size(); josc++) dim_post *= (long)nessential[josc];
      const long rk = r / dim_post;
      freq = freq + rk * (gate_rot_freq_[iosc] * 2. * M_PI);
      r = r % dim_post;
    }
    const double ra = cos(freq * time), rb = sin(freq * time);
    for (long c = 0; c < n; c++) {
      const cplx v = at(row, c);
      at(row, c) = cplx(ra * v.real() - rb * v.imag(), ra * v.imag() + rb * v.real());
    }
  }
  /* lift to the full dimension: identity on the non-essential levels */
  V.assign((size_t)dim_rho * dim_rho, cplx(0.0, 0.0));
  for (long rf = 0; rf < dim_rho; rf++) {
    if (isEssential(rf, nlevels, nessential)) {
      const long re = mapFullToEss(rf, nlevels, nessential);
      for (long ce = 0; ce < n; ce++) V[(size_t)rf * dim_rho + mapEssToFull(ce, nlevels, nessential)] = at(re, ce);
    } else {
      V[(size_t)rf * dim_rho + rf] = 1.0;
    }
  }
}

void Gate::applyGate(const Vec& state, Vec& VrhoV) const {
  if (dim_rho == 0) return;
  const long N = dim_rho;
  VrhoV.assign(state.size(), 0.0);
  if (lindbladtype == LindbladType::NONE) {
    for (long r = 0; r < N; r++) {
      cplx s(0.0, 0.0);
      for (long c = 0; c < N; c++) {
        const cplx v = V[(size_t)r * N + c];
        if (v != cplx(0.0, 0.0)) s += v * cplx(state[c], state[c + N]);
      }
      VrhoV[r] = s.real(); VrhoV[r + N] = s.imag();
    }
    return;
  }
  const long dim = N * N;
  std::vector<cplx> T((size_t)dim, cplx(0.0, 0.0));   /* T = V rho */
  for (long r = 0; r < N; r++)
    for (long k = 0; k < N; k++) {
      const cplx v = V[(size_t)r * N + k];
      if (v == cplx(0.0, 0.0)) continue;
      for (long c = 0; c < N; c++) T[(size_t)r * N + c] += v * cplx(state[getVecID(k, c, N)], state[getVecID(k, c, N) + dim]);
    }
  for (long r = 0; r < N; r++)                          /* (T V^dagger)_{rc} = sum_k T_{rk} conj(V_{ck}) */
    for (long c = 0; c < N; c++) {
      cplx s(0.0, 0.0);
      for (long k = 0; k < N; k++) {
        const cplx v = V[(size_t)c * N + k];
        if (v != cplx(0.0, 0.0)) s += T[(size_t)r * N + k] * std::conj(v);
      }
      VrhoV[getVecID(r, c, N)] = s.real(); VrhoV[getVecID(r, c, N) + dim] = s.imag();
    }
}
