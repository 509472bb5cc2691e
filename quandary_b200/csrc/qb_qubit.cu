/* qb_qubit.cu -- instantiates the qubit-system sweep kernels (qb_qubit.cuh). */
#include "qb_qubit.cuh"

template <int Q, bool LIND, int E>
static void set(QKernelSet* ks) {
  using C = qq::K<Q, LIND, E>;
  ks->fwd = qq::qfwd_kernel<Q, LIND, E>;
  ks->adj = qq::qadj_kernel<Q, LIND, E, false>;
  ks->adj_stored = qq::qadj_kernel<Q, LIND, E, true>;
  ks->app = qq::qapply_kernel<Q, LIND, E>;
  ks->threads = C::T;
  ks->nw = C::NW;
  ks->vec_bytes = C::VEC;
  ks->nred = C::NRED;
  ks->dim = C::DIM;
  ks->epT = E;
}

/* Kernels exist for Lindblad systems of 3 and 4 qubits (dim 64 and 256).  epT: preferred elements per thread (0 = default). */
void qq_pick(int Q, bool lind, int epT, QKernelSet* ks) {
  ks->fwd = nullptr; ks->adj = nullptr; ks->app = nullptr;
  if (!lind) return;
  if (Q == 3) { set<3, true, 2>(ks); return; }
  if (Q == 4) {
    if (epT == 8) set<4, true, 8>(ks);
    else if (epT == 2) set<4, true, 2>(ks);
    else if (epT == 1) set<4, true, 1>(ks);
    else set<4, true, 4>(ks);
  }
}

size_t qq_smem(const QKernelSet& ks, int mcap, int nvec, bool adj, bool with_traj, bool with_stage, unsigned recbytes, int nd) {
  return qq::qlayout(ks.nw, ks.vec_bytes, ks.nred, mcap, nvec, adj, with_traj, with_stage, recbytes, nd).total;
}
