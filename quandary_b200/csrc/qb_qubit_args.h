/* qb_qubit_args.h -- launch arguments and shared-memory layout of the qubit-system kernels (host + device). */
#pragma once
#include "qb_common.h"

namespace qq {
constexpr int kNR = 4;       /* control-record ring slots */
constexpr int kNT = 2;       /* trajectory ring slots (adjoint): the primal input is requested one full step ahead */
constexpr int kGivW = 6 * 32 * 8;   /* warp-private (fallback Arnoldi): cc ss grs rdiag nrs hcol, 32 doubles each */
constexpr int kSMax = 6;     /* = qmono::kSMax (qb_mono.cuh; this header is also compiled by the host side) */
constexpr int kSBlk = 4;     /* = qmono::kSBlk */
constexpr int kNRed = 28;    /* values per reduction buffer half: (kSMax+1)(kSMax+2)/2 = 28 Gram entries */

/* shared-memory layout, shared by host (size) and device (offsets) */
struct QLayout {
  unsigned bars, red, giv, upd, xb, sb, u, traj, rec, grad, total;
};
__host__ __device__ inline QLayout qlayout(int NW, unsigned VEC, bool adj, bool with_traj, bool with_stage, unsigned recbytes, int nd) {
  QLayout L;
  unsigned o = 0;
  L.bars = o; o += 128;
  L.red = o; o += 2u * kNRed * NW * 8u;
  L.giv = o; o += (unsigned)NW * kGivW;
  L.upd = o; o += adj ? kMaxUpd * 16u : 0u;
  L.xb = o; o += adj ? VEC : 2u * VEC;          /* forward: double-buffered source of the state bulk stores */
  L.sb = o; o += with_stage ? 2u * VEC : 0u;    /* forward: double-buffered source of the stage bulk stores */
  L.u = o; o += (unsigned)(kSMax + 1) * VEC;    /* Krylov vectors u_j = (alpha M)^j b of the current solve */
  L.traj = o; o += (adj && with_traj) ? kNT * VEC : 0u;
  L.rec = o; o += kNR * recbytes;
  L.grad = o; o += adj ? (unsigned)nd * 8u : 0u;
  L.total = (o + 15u) & ~15u;
  return L;
}

struct QArgs {
  const double* rec;     /* [ntime][recstride] per-step records */
  int recstride;         /* doubles, even */
  unsigned rec_fwd_bytes, rec_bytes;   /* multiples of 16 */
  int rec_gcs, rec_gval, rec_gidx;     /* offsets (doubles) inside a record */
  const int* upd;
  double2* state;
  double2* state_out;
  double2* final_state;
  double2* traj;         /* [ninit][ntime+1][dim] states, or null (forward: not stored; adjoint: stages only) */
  double2* stage_traj;   /* [ninit][ntime][dim] stored stages (QB_FLAG_STORE_STAGE) or null */
  double2* scratch;      /* [ninit][mcap+1][dim] Arnoldi basis of the fallback solver (own elements only: thread-private) */
  double* hh;            /* [ninit][NW][mcap*mcap] Hessenberg columns of the fallback solver */
  double* grad_ic;
  unsigned long long* counters;
  PenArgs pen;
  int ninit, mcap;
};

}  // namespace qq
