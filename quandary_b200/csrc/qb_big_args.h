/* qb_big_args.h -- launch arguments and shared-memory layout of the large-state kernels (host + device). */
#pragma once
#include "qb_common.h"

namespace qbig {
constexpr int T = 1024;
constexpr int NW = T / 32;
constexpr int kGiv = 6 * 32 * 8;
constexpr int kBMax = 9;     /* longest Krylov basis of the one-reduction solver (qb_big.cuh gmres_mono); longer solves: Arnoldi fallback */
constexpr int kBGram = (kBMax + 1) * (kBMax + 2) / 2;

struct BLayout {
  unsigned red, giv, hh, sq2, upd, gp, gt, diag, xb0, xb1, grad, total;
};
__host__ __device__ inline BLayout blayout(int nred, int mcap, int L, int dimp, bool adj, int nd) {
  BLayout B;
  unsigned o = 0;
  B.red = o; o += 2u * nred * NW * 8u;
  B.giv = o; o += kGiv;
  B.hh = o; o += (unsigned)(mcap * mcap) * 8u; o = (o + 15u) & ~15u;
  B.sq2 = o; o += (unsigned)((L + 1) * (L + 1)) * 8u; o = (o + 15u) & ~15u;
  B.upd = o; o += adj ? kMaxUpd * 16u : 0u;
  B.gp = o; o += (unsigned)(kBGram * NW) * 8u;              /* per-warp partial sums of the Gram entries */
  B.gt = o; o += (unsigned)(kBGram + 3 * (kBMax + 2) + (kBMax + 1) * kBMax / 2 + 16) * 8u; o = (o + 15u) & ~15u;   /* totals, L D L^T work arrays, results */
  B.diag = o; o += (unsigned)dimp * 16u;
  B.xb0 = o; o += (unsigned)dimp * 16u;
  B.xb1 = o; o += (unsigned)dimp * 16u;
  B.grad = o; o += adj ? (unsigned)nd * 8u : 0u;
  B.total = (o + 15u) & ~15u;
  return B;
}

struct BArgs {
  const double* tab;
  const int* gidx;
  const double* gval;
  const double* gcs;
  const int* upd;
  double2* state;
  double2* state_out;
  double2* final_state;
  double2* traj;         /* [ninit][ntime+1][dimp] or null */
  double2* stage_traj;   /* [ninit][ntime][dimp] or null */
  double2* scratch;      /* [ninit][nvec + 3][dimp]: Krylov basis, rhs stash, adjoint-state stash, stage stash */
  double* grad_ic;
  unsigned long long* counters;
  PenArgs pen;
  int ninit, mcap, nvec, store, dimp, L;
};

}  // namespace qbig
