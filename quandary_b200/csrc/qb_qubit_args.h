/* qb_qubit_args.h -- launch arguments and shared-memory layout of the qubit-system kernels (host + device). */
#pragma once
#include "qb_common.h"

namespace qq {
constexpr int kNR = 4;       /* control-record ring slots */
constexpr int kNT = 2;       /* trajectory ring slots (adjoint): x_{n-1} is requested one full step ahead */
constexpr int kGivW = 6 * 32 * 8;   /* warp-private: cc ss grs rdiag nrs hcol, 32 doubles each */

/* shared-memory layout, shared by host (size) and device (offsets) */
struct QLayout {
  unsigned bars, red, giv, hh, upd, xb, sb, v, traj, rec, grad, total;
};
__host__ __device__ inline QLayout qlayout(int NW, unsigned VEC, int nred, int mcap, int nvec, bool adj, bool with_traj, bool with_stage, unsigned recbytes, int nd) {
  QLayout L;
  unsigned o = 0;
  L.bars = o; o += 128;
  L.red = o; o += 2u * nred * NW * 8u; o = (o + 15u) & ~15u;
  L.giv = o; o += (unsigned)NW * kGivW;
  L.hh = o; o += (unsigned)NW * (unsigned)(mcap * mcap) * 8u; o = (o + 15u) & ~15u;
  L.upd = o; o += adj ? kMaxUpd * 16u : 0u;
  L.xb = o; o += 2u * VEC;
  L.sb = o; o += with_stage ? 2u * VEC : 0u;
  L.v = o; o += (unsigned)nvec * VEC;
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
  double2* traj;
  double2* stage_traj;   /* [ninit][ntime][dim] stored stages (QB_FLAG_STORE_STAGE) or null */
  double* grad_ic;
  unsigned long long* counters;
  PenArgs pen;
  int ninit, mcap, nvec, store;
};

}  // namespace qq
