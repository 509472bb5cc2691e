/* qb_bigs_q5.cu -- large-state kernels, Schroedinger solver (collapse_type = none, 256 < N <= 4096), 5 oscillator(s) (qb_big.cuh) */
#include "qb_big.cuh"
template <int Q, bool HASJ, int E>
static void set(BKernelSet* ks) {
  ks->fwd = qbig::bfwd_kernel<Q, false, HASJ, E>;
  ks->adj = qbig::badj_kernel<Q, false, HASJ, E>;
  ks->app = qbig::bapply_kernel<Q, false, HASJ, E>;
  ks->threads = qbig::T;
  ks->dimp = E * qbig::T;
  ks->nred = qbig::nred<Q>();
}
void qbig_pick_s5(bool hasj, int dim, BKernelSet* ks) {
  constexpr int Q = 5;
  ks->fwd = nullptr;
  if (dim > 4 * qbig::T) return;
  const bool big = dim > qbig::T;
  if constexpr (Q >= 2) {
    if (hasj) { if (big) set<Q, true, 4>(ks); else set<Q, true, 1>(ks); return; }
  }
  if (big) set<Q, false, 4>(ks); else set<Q, false, 1>(ks);
}
