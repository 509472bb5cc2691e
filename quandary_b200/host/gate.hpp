/* gate.hpp -- target gates; mirrors the reference's Gate hierarchy (include/gate.hpp, src/gate.cpp:1-571).  The reference
 * assembles the sparse superoperator conj(V) (x) V; here the (small) unitary is kept dense and V rho V^dagger is applied directly. */
#pragma once
#include <complex>

#include "defs.hpp"

class Gate {
 public:
  Gate() {}
  /* kind: none, xgate, ygate, zgate, hadamard, cnot, swap, swap0q, cqnot, qft, file (src/gate.cpp:535-560) */
  Gate(const std::string& kind, const std::string& filename, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential, double time,
       const std::vector<double>& gate_rot_freq, LindbladType lindbladtype, bool quietmode);
  /* VrhoV = V state V^dagger (Lindblad) or V state (Schroedinger), in the [Re | Im] layout (src/gate.cpp:260-283) */
  void applyGate(const Vec& state, Vec& VrhoV) const;
  bool isDummy() const { return dim_rho == 0; }

 private:
  long dim_ess = 0, dim_rho = 0;
  LindbladType lindbladtype = LindbladType::NONE;
  std::vector<std::complex<double>> V;   /* full-dimension gate, row-major dim_rho x dim_rho, rotating frame applied */
};
