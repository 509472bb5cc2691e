/* qb_qubit_v0.cu -- instantiates the qubit-system sweep kernels (qb_qubit.cuh) for 3 qubits, 2 elements per thread. */
#include "qb_qubit.cuh"

void qq_set_v0(QKernelSet* ks) {
  using C = qq::K<3, true, 2>;
  ks->fwd[0] = qq::qfwd_kernel<3, true, 2, true>;
  ks->fwd[1] = qq::qfwd_kernel<3, true, 2, false>;
  ks->adj[0] = qq::qadj_kernel<3, true, 2, false, true>;
  ks->adj[1] = qq::qadj_kernel<3, true, 2, false, false>;
  ks->adj_stored[0] = qq::qadj_kernel<3, true, 2, true, true>;
  ks->adj_stored[1] = qq::qadj_kernel<3, true, 2, true, false>;
  ks->app = qq::qapply_kernel<3, true, 2>;
  ks->threads = C::T;
  ks->nw = C::NW;
  ks->vec_bytes = C::VEC;
  ks->dim = C::DIM;
  ks->epT = 2;
}
