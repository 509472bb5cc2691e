/* qb_qubit_v1.cu -- instantiates the qubit-system sweep kernels (qb_qubit.cuh) for 4 qubits, 4 elements per thread. */
#include "qb_qubit.cuh"

void qq_set_v1(QKernelSet* ks) {
  using C = qq::K<4, true, 4>;
  ks->fwd[0] = qq::qfwd_kernel<4, true, 4, true>;
  ks->fwd[1] = qq::qfwd_kernel<4, true, 4, false>;
  ks->adj[0] = qq::qadj_kernel<4, true, 4, false, true>;
  ks->adj[1] = qq::qadj_kernel<4, true, 4, false, false>;
  ks->adj_stored[0] = qq::qadj_kernel<4, true, 4, true, true>;
  ks->adj_stored[1] = qq::qadj_kernel<4, true, 4, true, false>;
  ks->app = qq::qapply_kernel<4, true, 4>;
  ks->threads = C::T;
  ks->nw = C::NW;
  ks->vec_bytes = C::VEC;
  ks->dim = C::DIM;
  ks->epT = 4;
}
