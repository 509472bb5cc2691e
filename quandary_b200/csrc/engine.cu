/*
 * engine.cu -- B200 (sm_100a) forward/adjoint Lindblad time-stepping engine behind include/quandary_b200.h.
 *
 * What runs here replaces, for ALL initial conditions of one optimizer evaluation at once, the reference's
 *   TimeStepper::solveODE / solveAdjointODE        src/timestepper.cpp:96-253
 *   ImplMidpoint::evolveFWD / evolveBWD / Neumann   src/timestepper.cpp:584-727   (+ PETSc KSPGMRES, not in tree)
 *   applyRHS_matfree / _transpose                   src/mastereq.cpp:1279-2893, include/mastereq.hpp:316-912
 *   compute_dRHS_dParams_matfree                    src/mastereq.cpp:970-1276
 *   MasterEq::assemble_RHS, Oscillator::evalControl[_diff], BSpline2nd/BSpline0   src/mastereq.cpp:657-678,
 *                                                   src/oscillator.cpp:281-381, src/controlbasis.cpp:20-96,219-255
 *
 * Design:
 *   - one CTA per initial condition; the whole sweep (all ntime steps) is ONE kernel launch.  Three kernel families:
 *     qb_qubit.cuh (qubit systems, E elements per thread, TMA bulk copies), qb_kernels.cuh (any level tuple, dim <= 256,
 *     one thread per element), qb_big.cuh (256 < dim <= 4096).  State, Krylov basis and gradient accumulator live in
 *     shared memory / registers; the only per-step HBM traffic is the trajectory (16*dim B written per forward step,
 *     read once per adjoint step, prefetched one step ahead).
 *   - every thread keeps the stencil of ITS element (neighbour offsets, sqrt factors, diagonal) in registers for
 *     the whole sweep; the real-vectorised operator is  M = H + D + L1  with H skew-symmetric, so M^T = -H + D + L1^T.
 *   - the controls p_k(t_{n+1/2}), q_k(t_{n+1/2}), the carrier phases and the non-zero B-spline basis values of
 *     every step are evaluated once per parameter set by a device kernel (they are shared by all initial conditions).
 *   - the linear solve of each implicit-midpoint step is done in-kernel: GMRES (same algorithm and stopping rule as
 *     the reference's KSPGMRES set-up) or the reference's Neumann iteration.
 *   - reductions are fixed-tree (xor-butterfly in the warp, warp partials summed in warp order) => bitwise reproducible.
 */
#include "qb_qubit_args.h"
#include "qb_big_args.h"

#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x)                                                                                        \
  do {                                                                                                     \
    cudaError_t err_ = (x);                                                                                \
    if (err_ != cudaSuccess) return fail(QB_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(err_), __FILE__, __LINE__); \
  } while (0)

/* ---- NCCL through dlopen: no link-time dependency, nothing loaded unless a communicator is requested ---- */
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, char[128], int) = nullptr;   /* ncclUniqueId is a 128-byte struct passed by value */
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char internal[128]; };
NcclApi g_nccl;
int nccl_load() {
  if (g_nccl.lib) return QB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return fail(QB_ERR_UNSUPPORTED, "NCCL library not found (libnccl.so.2): %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, char[128], int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) return fail(QB_ERR_UNSUPPORTED, "NCCL symbols missing");
  return QB_OK;
}

/* ---------------------------------------------------------------------------------------------------------
 * Control table.  One thread per time step.  Per step n (tstart = n*dt, tstop = (n+1)*dt as in
 * src/timestepper.cpp:128-129; tmid = (tstart+tstop)/2, dtn = tstop-tstart as in :587-590):
 *   tab[n*tabstride + 0]            dtn
 *   tab[... + 1 + k]                p_k(tmid)          Oscillator::evalControl, src/oscillator.cpp:281-337
 *   tab[... + 1 + Q + k]            q_k(tmid)
 *   tab[... + 1 + 2Q + id]          cos(eta_id*tmid)   src/mastereq.cpp:672-675
 *   tab[... + 1 + 2Q + npairs + id] sin(eta_id*tmid)
 * and for the gradient (Oscillator::evalControl_diff, src/oscillator.cpp:339-381):
 *   gcs[(n*ncar_total + c)*2 + {0,1}]  cos/sin(Omega_c * tmid)
 *   gidx[(n*Q + k)*4 + j], j<3         parameter index (inside oscillator k's block, carrier 0, real part) of the
 *                                      j-th non-zero basis function of the active segment, -1 if none
 *   gidx[(n*Q + k)*4 + 3]              nsplines of the active segment
 *   gval[(n*Q + k)*3 + j]              its value B_l(tmid)
 * --------------------------------------------------------------------------------------------------------- */
struct CtrlArgs {
  const qb_segment* segs;
  int seg_off[kMaxOsc + 1];
  const double* carrier;
  const double* pipulse;
  int pp_off[kMaxOsc + 1];
  const double* alpha;
};

/* BSpline2nd::basisfunction, src/controlbasis.cpp:81-96 */
__device__ __forceinline__ double bspline2_basis(const qb_segment& s, int id, double t) {
  double dtknot = (s.tstop - s.tstart) / (double)(s.nsplines - 2);
  double width = 3.0 * dtknot;
  double tcenter = s.tstart + dtknot * ((id + 1) - 1.5);
  double tau = (t - tcenter) / width;
  if (tau < -1. / 2. || tau >= 1. / 2.) return 0.0;
  double val = 0.0;
  if (-1. / 2. <= tau && tau < -1. / 6.) val = 9. / 8. + 9. / 2. * tau + 9. / 2. * (tau * tau);
  else if (-1. / 6. <= tau && tau < 1. / 6.) val = 3. / 4. - 9. * (tau * tau);
  else if (1. / 6. <= tau && tau < 1. / 2.) val = 9. / 8. - 9. / 2. * tau + 9. / 2. * (tau * tau);
  return val;
}

/* Oscillator::evalControl for oscillator k at time t; optionally records the gradient-table entries. */
__device__ void eval_control_dev(const DevSys& S, const CtrlArgs& A, int k, double t, double* p_out, double* q_out, int* gidx,
                                 double* gval, double* gcs /* [ncar_k][2] or null */) {
  double p = 0.0, q = 0.0;
  const double* params = A.alpha + S.osc_off[k];
  const int nc = S.car_off[k + 1] - S.car_off[k];
  const double* om = A.carrier + S.car_off[k];
  if (gidx) { gidx[0] = gidx[1] = gidx[2] = -1; gidx[3] = 0; gval[0] = gval[1] = gval[2] = 0.0; }
  if (gcs) for (int f = 0; f < nc; f++) { double sn, cs; sincos(om[f] * t, &sn, &cs); gcs[2 * f] = cs; gcs[2 * f + 1] = sn; }
  if (S.osc_off[k + 1] > S.osc_off[k]) {
    for (int si = A.seg_off[k]; si < A.seg_off[k + 1]; si++) {
      const qb_segment s = A.segs[si];
      if (s.tstart <= t && s.tstop >= t) {
        const int ns = s.nsplines;
        double sum_p = 0.0, sum_q = 0.0;
        for (int f = 0; f < nc; f++) {
          double B1 = 0.0, B2 = 0.0;
          if (s.type == QB_CTRL_BSPLINE2ND) {
            int cnt = 0;
            for (int l = 0; l < ns; l++) {
              if (s.enforce_bc && (l <= 1 || l >= ns - 2)) continue;
              double Blt = bspline2_basis(s, l, t);
              B1 += params[s.skip + f * ns * 2 + l] * Blt;
              B2 += params[s.skip + f * ns * 2 + l + ns] * Blt;
              if (f == 0 && gidx && Blt != 0.0 && cnt < 3) { gidx[cnt] = s.skip + l; gval[cnt] = Blt; cnt++; }
            }
          } else { /* BSpline0::evaluate, src/controlbasis.cpp:229-243 */
            double dtknot = (s.tstop - s.tstart) / (ns - 1.0);
            int id = (int)ceil((t - s.tstart) / dtknot - 0.5);
            if (id >= 0 && id < ns) {
              B1 = params[s.skip + f * ns * 2 + id];
              B2 = params[s.skip + f * ns * 2 + id + ns];
              if (f == 0 && gidx) { gidx[0] = s.skip + id; gval[0] = 1.0; }
            }
          }
          double sn, cs;
          sincos(om[f] * t, &sn, &cs);
          sum_p += cs * B1 - sn * B2;
          sum_q += sn * B1 + cs * B2;
        }
        if (gidx) gidx[3] = ns;
        p = sum_p; q = sum_q;
        break;
      }
    }
  }
  for (int i = A.pp_off[k]; i < A.pp_off[k + 1]; i++) {
    const double* pp = A.pipulse + 3 * i;
    if (pp[0] <= t && t <= pp[1]) { double a = pp[2] / sqrt(2.0); p = a; q = a; }
  }
  *p_out = p; *q_out = q;
}

__global__ void ctrl_table_kernel(DevSys S, CtrlArgs A, double* __restrict__ tab, int* __restrict__ gidx, double* __restrict__ gval,
                                  double* __restrict__ gcs) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= S.ntime + S.ee) return;
  double tstart = n * S.dt, tstop = (n + 1) * S.dt;
  if (S.nstage > 1) {
    /* sub-step i of time step nf: tcurr advanced stage by stage exactly as CompositionalImplMidpoint::evolveFWD does
     * (src/timestepper.cpp:783-801) */
    const int nf = n / S.nstage, i = n % S.nstage;
    const double dtf = (nf + 1) * S.dt - nf * S.dt;
    double tcurr = nf * S.dt;
    for (int s = 0; s < i; s++) tcurr = tcurr + S.gamma[s] * dtf;
    tstart = tcurr;
    tstop = tcurr + S.gamma[i] * dtf;
  }
  double tmid = (tstart + tstop) / 2.0;
  if (S.ee) tmid = tstart;      /* explicit Euler: row n holds the controls at t_n (forward step n uses A(tstart), the adjoint step n -> n-1 uses A(tstop = t_n)) */
  double* row = tab + (size_t)n * S.tabstride;
  row[0] = tstop - tstart;
  for (int k = 0; k < S.Q; k++) {
    double p, q;
    eval_control_dev(S, A, k, tmid, &p, &q, gidx + ((size_t)n * S.Q + k) * 4, gval + ((size_t)n * S.Q + k) * 3,
                     gcs + ((size_t)n * S.ncar_total + S.car_off[k]) * 2);
    row[1 + k] = p;
    row[1 + S.Q + k] = q;
  }
  for (int id = 0; id < S.npairs; id++) {
    double sn, cs;
    sincos(S.eta[id] * tmid, &sn, &cs);
    row[1 + 2 * S.Q + id] = cs;
    row[1 + 2 * S.Q + S.npairs + id] = sn;
  }
}

/* p_k(t), q_k(t) (+ cos/sin eta t) at arbitrary times: one thread per time.  Row layout as the control table. */
__global__ void ctrl_eval_kernel(DevSys S, CtrlArgs A, const double* __restrict__ times, int nt, double* __restrict__ tab) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nt) return;
  double t = times[n];
  double* row = tab + (size_t)n * S.tabstride;
  row[0] = 0.0;
  for (int k = 0; k < S.Q; k++) {
    double p, q;
    eval_control_dev(S, A, k, t, &p, &q, nullptr, nullptr, nullptr);
    row[1 + k] = p;
    row[1 + S.Q + k] = q;
  }
  for (int id = 0; id < S.npairs; id++) {
    double sn, cs;
    sincos(S.eta[id] * t, &sn, &cs);
    row[1 + 2 * S.Q + id] = cs;
    row[1 + 2 * S.Q + S.npairs + id] = sn;
  }
}

/* Per-step records for the qubit kernels: [tab row | gcs row | gval row | gidx row (ints)] with a fixed 16-byte-aligned
 * stride, so that one TMA bulk copy brings everything a step needs into shared memory. */
__global__ void pack_rec_kernel(DevSys S, const double* __restrict__ tab, const int* __restrict__ gidx, const double* __restrict__ gval,
                                const double* __restrict__ gcs, double* __restrict__ rec, int recstride, int o_gcs, int o_gval, int o_gidx) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= S.ntime) return;
  double* r = rec + (size_t)n * recstride;
  for (int i = 0; i < S.tabstride; i++) r[i] = tab[(size_t)n * S.tabstride + i];
  for (int i = 0; i < 2 * S.ncar_total; i++) r[o_gcs + i] = gcs[(size_t)n * S.ncar_total * 2 + i];
  for (int i = 0; i < 3 * S.Q; i++) r[o_gval + i] = gval[(size_t)n * S.Q * 3 + i];
  int* ri = reinterpret_cast<int*>(r + o_gidx);
  for (int i = 0; i < 4 * S.Q; i++) ri[i] = gidx[(size_t)n * S.Q * 4 + i];
}

/* penalty_ic = sum_{n=1..ntime} wdt[n]*(c0 + <a,x_n> + <d,x_n^2>) + inv_ntime*<g,x_n^2>: one CTA per initial condition,
 * fixed-order reduction (TimeStepper::penaltyIntegral accumulated over solveODE, src/timestepper.cpp:141,256-300) */
__global__ void penalty_reduce_kernel(const double2* __restrict__ traj, int ntime, int dim, int dimp, PenArgs P, const double* __restrict__ c0, double* __restrict__ out,
                                      const double* __restrict__ ovval) {
  __shared__ double sh[256];
  const int ic = blockIdx.x;
  const double2* tr = traj + (size_t)ic * (ntime + 1) * dimp;
  double acc = 0.0;
  for (int n = 1; n <= ntime; n++) {
    const double w = P.wdt[n];
    double s1 = 0.0, s2 = 0.0;
    for (int it = threadIdx.x; it < dim; it += blockDim.x) {
      const double2 x = tr[(size_t)n * dimp + it];
      if (!ovval) {
        const double2 a = P.a[(size_t)ic * dimp + it], d = P.d[it];
        s1 += a.x * x.x + a.y * x.y + d.x * x.x * x.x + d.y * x.y * x.y;
      }
      if (P.g) { const double2 g = P.g[it]; s2 += g.x * x.x * x.x + g.y * x.y * x.y; }
    }
    acc += w * s1 + (P.gw ? P.gw[n] : 0.0) * s2;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double wsum = 0.0;
    for (int n = 1; n <= ntime; n++) wsum += P.wdt[n];
    out[ic] = ovval ? sh[0] + ovval[ic] : sh[0] + c0[ic] * wsum;     /* overlap mode: the objective part comes from overlap_table_kernel */
  }
}

/* Schroedinger + Jtrace penalty: per initial condition and step the overlap of the stored state with the target,
 * ov[ic][n] = (-2 Re<a,x_n>/purity^2, -2 Im<a,x_n>) (the coefficients of dJ/dx, HilbertSchmidtOverlap_diff through finalizeJ_diff,
 * src/optimtarget.cpp:404-447, 883-897), and val[ic] = sum_n wdt[n] * (1 - (Re<a,x_n>/purity)^2 - (Im<a,x_n>)^2) (:756-760, 864-880) */
__global__ void __launch_bounds__(256) overlap_table_kernel(const double2* __restrict__ traj, int ntime, int dim, int dimp, const double2* __restrict__ a,
                                                            const double* __restrict__ purity, const double* __restrict__ wdt, double2* __restrict__ ov,
                                                            double* __restrict__ val) {
  __shared__ double sh[2 * 256];
  const int ic = blockIdx.x;
  const double2* tr = traj + (size_t)ic * (ntime + 1) * dimp;
  const double2* ai = a + (size_t)ic * dimp;
  const double p = purity ? purity[ic] : 1.0;
  double acc = 0.0;
  for (int n = 0; n <= ntime; n++) {
    double r = 0.0, im = 0.0;
    for (int it = threadIdx.x; it < dim; it += 256) {
      const double2 x = tr[(size_t)n * dimp + it], t = ai[it];
      r += t.x * x.x + t.y * x.y;
      im += -t.y * x.x + t.x * x.y;
    }
    sh[threadIdx.x] = r; sh[256 + threadIdx.x] = im;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if ((int)threadIdx.x < st) { sh[threadIdx.x] += sh[threadIdx.x + st]; sh[256 + threadIdx.x] += sh[256 + threadIdx.x + st]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double rr = sh[0], ii = sh[256];
      ov[(size_t)ic * (ntime + 1) + n] = make_double2(-2.0 * rr / (p * p), -2.0 * ii);
      if (n >= 1) acc += wdt[n] * (1.0 - (rr / p) * (rr / p) - ii * ii);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) val[ic] = acc;
}

/* ---- final-time objective (OptimTarget, src/optimtarget.cpp:343-447, 713-862) -------------------------------------------
 * One CTA per initial condition over the final states [ninit][dimp]; fixed-tree reductions.  ObjArgs is filled by
 * qb_engine_set_objective. */
struct ObjArgs {
  const double2* target;   /* [ninit][dimp] or null (pure target) */
  const double* purity;    /* [ninit] */
  const double* weight;    /* [ninit] */
  double fid_scale;
  int kind, pure_id, lind, N, dim, dimp;
};

template <int NV>
__device__ __forceinline__ void obj_block_sum(double (&v)[NV], double* sh) {
  for (int i = 0; i < NV; i++) sh[i * 256 + threadIdx.x] = v[i];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int i = 0; i < NV; i++) sh[i * 256 + threadIdx.x] += sh[i * 256 + threadIdx.x + s];
    __syncthreads();
  }
  for (int i = 0; i < NV; i++) v[i] = sh[i * 256];
}

/* out4[ic] = { w J_re, w J_im, fid_scale HS_re, fid_scale HS_im } (evalJ :713-797, HilbertSchmidtOverlap :343-402) */
__global__ void __launch_bounds__(256) objective_eval_kernel(const double2* __restrict__ xT, ObjArgs O, double* __restrict__ out4) {
  __shared__ double sh[3 * 256];
  const int ic = blockIdx.x;
  const double2* x = xT + (size_t)ic * O.dimp;
  const double2* tau = O.target ? O.target + (size_t)ic * O.dimp : nullptr;
  const int idm = O.pure_id < 0 ? 0 : (O.lind ? O.pure_id + O.pure_id * O.N : O.pure_id);
  /* v[0], v[1]: Hilbert-Schmidt overlap; v[2]: the objective's own sum (Frobenius distance / measurement) */
  double v[3] = {0.0, 0.0, 0.0};
  for (int it = threadIdx.x; it < O.dim; it += 256) {
    const double2 xi = x[it];
    if (tau) {
      const double2 t = tau[it];
      v[0] += t.x * xi.x + t.y * xi.y;
      if (!O.lind) v[1] += -t.y * xi.x + t.x * xi.y;
      if (O.kind == QB_OBJ_JFROBENIUS) { const double dx = t.x - xi.x, dy = t.y - xi.y; v[2] += dx * dx + dy * dy; }
    } else {
      if (it == idm) { v[0] = xi.x; v[1] = xi.y; }
      if (O.kind == QB_OBJ_JFROBENIUS) { const double dx = xi.x - (it == idm ? 1.0 : 0.0); v[2] += dx * dx + xi.y * xi.y; }
      if (O.kind == QB_OBJ_JMEASURE) {
        if (O.lind) { const int r = it % O.N, c = it / O.N; if (r == c) v[2] += fabs((double)(r - O.pure_id)) * xi.x; }
        else v[2] += fabs((double)(it - O.pure_id)) * (xi.x * xi.x + xi.y * xi.y);
      }
    }
  }
  obj_block_sum<3>(v, sh);
  if (threadIdx.x == 0) {
    double jre = 0.0, jim = 0.0;
    if (O.kind == QB_OBJ_JFROBENIUS) jre = v[2] / 2.0;
    else if (O.kind == QB_OBJ_JTRACE) { jre = v[0] / (O.purity ? O.purity[ic] : 1.0); jim = v[1]; }
    else jre = v[2];
    const double w = O.weight[ic];
    out4[4 * ic + 0] = w * jre; out4[4 * ic + 1] = w * jim;
    out4[4 * ic + 2] = O.fid_scale * v[0]; out4[4 * ic + 3] = O.fid_scale * v[1];
  }
}

/* sum[0..3] = sum over initial conditions, in index order within each of 256 strided partial sums, then a fixed tree */
__global__ void __launch_bounds__(256) objective_sum_kernel(const double* __restrict__ out4, int ninit, double* __restrict__ sum) {
  __shared__ double sh[4 * 256];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int ic = threadIdx.x; ic < ninit; ic += 256)
    for (int i = 0; i < 4; i++) v[i] += out4[4 * ic + i];
  obj_block_sum<4>(v, sh);
  if (threadIdx.x < 4) sum[threadIdx.x] = v[threadIdx.x];
}

/* terminal adjoint states: evalJ_diff (:800-862) / HilbertSchmidtOverlap_diff (:404-447) seeded with w_i * (re_bar, im_bar) */
__global__ void __launch_bounds__(256) objective_bar_kernel(const double2* __restrict__ xT, ObjArgs O, double re_bar, double im_bar, double2* __restrict__ xbar) {
  const int ic = blockIdx.x;
  const double2* x = xT + (size_t)ic * O.dimp;
  const double2* tau = O.target ? O.target + (size_t)ic * O.dimp : nullptr;
  double2* xb = xbar + (size_t)ic * O.dimp;
  const int idm = O.pure_id < 0 ? -1 : (O.lind ? O.pure_id + O.pure_id * O.N : O.pure_id);
  const double w = O.weight[ic];
  const double rb = w * re_bar, ib = w * im_bar;
  const double scale = 1.0 / (O.purity ? O.purity[ic] : 1.0);
  for (int it = threadIdx.x; it < O.dimp; it += 256) {
    double2 o = make_double2(0.0, 0.0);
    if (it < O.dim) {
      const double2 xi = x[it];
      if (O.kind == QB_OBJ_JFROBENIUS) {
        if (tau) {
          const double jb = rb / 2.0;
          const double2 t = tau[it];
          o.x = 2.0 * jb * xi.x; o.y = 2.0 * jb * xi.y;
          o.x += -2.0 * jb * t.x; o.y += -2.0 * jb * t.y;
        } else {
          o.x = rb * xi.x; o.y = rb * xi.y;
          if (it == idm) o.x += -1.0 * rb;
        }
      } else if (O.kind == QB_OBJ_JTRACE) {
        if (!tau) { if (it == idm) { o.x = rb * scale; o.y = ib; } }
        else {
          const double2 t = tau[it];
          if (O.lind) { o.x = rb * scale * t.x; o.y = rb * scale * t.y; }
          else { o.x = t.x * rb * scale - t.y * ib; o.y = t.y * rb * scale + t.x * ib; }
        }
      } else {   /* Jmeasure */
        if (O.lind) { const int r = it % O.N, c = it / O.N; if (r == c) o.x = fabs((double)(r - O.pure_id)) * rb; }
        else { const double l = fabs((double)(it - O.pure_id)); o.x = 2. * rb * l * xi.x; o.y = 2. * rb * l * xi.y; }
      }
    }
    xb[it] = o;
  }
}

/* out[ic] = sum_{n=1..ntime-1} sum_i (|x_{n+1,i}|^2 - 2|x_{n,i}|^2 + |x_{n-1,i}|^2)^2 / dt^4 (penaltyDpDm over solveODE, src/timestepper.cpp:144-151,342-368) */
__global__ void dpdm_reduce_kernel(const double2* __restrict__ traj, int ntime, int dim, int dimp, double dtinv, double* __restrict__ out) {
  __shared__ double sh[256];
  const int ic = blockIdx.x;
  const double2* tr = traj + (size_t)ic * (ntime + 1) * dimp;
  double acc = 0.0;
  for (int it = threadIdx.x; it < dim; it += blockDim.x) {
    double2 a = tr[it], b = tr[(size_t)dimp + it];
    double pa = a.x * a.x + a.y * a.y, pb = b.x * b.x + b.y * b.y;
    for (int n = 1; n < ntime; n++) {
      const double2 c = tr[(size_t)(n + 1) * dimp + it];
      const double pc = c.x * c.x + c.y * c.y;
      const double t = pc - 2.0 * pb + pa;
      acc += dtinv * t * t;
      pa = pb; pb = pc;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[ic] = sh[0];
}

/* G[j] = sum over initial conditions (in order) of grad_ic[ic][j]:  the VecAXPY(G, 1.0, redgrad) of
 * src/optimproblem.cpp:442,514.  Also zeroes the 8 scalar slots that follow the gradient. */
__global__ void grad_reduce_kernel(const double* __restrict__ grad_ic, int ninit, int nd, double* __restrict__ G) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nd) return;
  double s = 0.0;
  for (int ic = 0; ic < ninit; ic++) s += grad_ic[(size_t)ic * nd + j];
  G[j] = s;
}

/* host [n][2*dim] blocked re|im  <->  device [n][dimp] interleaved */
__global__ void pack_kernel(const double* __restrict__ host_layout, double2* __restrict__ dev_layout, int dim, int dimp, int n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * dimp) return;
  int v = (int)(i / dimp), e = (int)(i % dimp);
  double2 o = make_double2(0.0, 0.0);
  if (e < dim) { o.x = host_layout[(size_t)v * 2 * dim + e]; o.y = host_layout[(size_t)v * 2 * dim + dim + e]; }
  dev_layout[i] = o;
}
__global__ void unpack_kernel(const double2* __restrict__ dev_layout, double* __restrict__ host_layout, int dim, int dimp, int n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * dimp) return;
  int v = (int)(i / dimp), e = (int)(i % dimp);
  if (e < dim) {
    double2 o = dev_layout[i];
    host_layout[(size_t)v * 2 * dim + e] = o.x;
    host_layout[(size_t)v * 2 * dim + dim + e] = o.y;
  }
}

__global__ void dfma_peak_kernel(double* __restrict__ out, int iters, double a) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  const double b = 1e-3;
#pragma unroll 16
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

size_t smem_bytes(const KernelSet& ks, int mcap, int nd) {
  return (size_t)ks.smem_fixed + (size_t)(mcap > 7 ? mcap : 7) * ks.vec_bytes + sizeof(double) * ((size_t)kLdh * (mcap + 1) + nd);     /* = Lay::nvec */
}

void pick(int Q, bool all2, bool hasj, bool lind, int dim, KernelSet* ks) {
  switch (Q) {
    case 1: qb_pick<1>(all2, hasj, lind, dim, ks); break;
    case 2: qb_pick<2>(all2, hasj, lind, dim, ks); break;
    case 3: qb_pick<3>(all2, hasj, lind, dim, ks); break;
    case 4: qb_pick<4>(all2, hasj, lind, dim, ks); break;
    default: qb_pick<5>(all2, hasj, lind, dim, ks); break;
  }
}

}  // namespace

/* =========================================================================================================
 * Engine object
 * ========================================================================================================= */
struct qb_engine {
  DevSys S;
  int dimp = 0;      /* threads per CTA = padded dimension */
  int mcap = 0;      /* Krylov capacity */
  bool lind = false;
  cudaStream_t stream = nullptr;      /* the stream everything is enqueued on */
  cudaStream_t own_stream = nullptr;  /* engine-owned stream (default) */
  /* CUDA-event pairs around every sweep kernel, resolved lazily (qb_engine_get_stats) so that timing never adds a
   * synchronisation to the stream */
  struct Timed { cudaEvent_t a, b; int kind; };
  std::vector<Timed> pending;
  int device = 0;
  SweepFn fwd = nullptr, adj = nullptr;
  ApplyFn app = nullptr;
  /* large-state kernels (qb_big.cuh): 256 < dim <= 4096 */
  bool use_big = false;
  BKernelSet bk;
  int maxlev = 0;
  int ntime_user = 0;   /* time steps of the caller; S.ntime counts compositional sub-steps */
  double* d_pen_gw = nullptr;
  double2* d_scratch = nullptr;
  /* penalty functional (qb_engine_set_penalty) */
  PenArgs pen;
  double2 *d_pen_a = nullptr, *d_pen_d = nullptr, *d_pen_g = nullptr;
  double *d_pen_wdt = nullptr, *d_pen_jbar = nullptr, *d_pen_c0 = nullptr, *d_pen_out = nullptr;
  int pen_ninit = 0;
  /* overlap mode of the penalty (Schroedinger + Jtrace) */
  double2* d_pen_ov = nullptr;
  double *d_pen_purity = nullptr, *d_pen_ovval = nullptr;
  bool pen_overlap = false, ov_valid = false;
  /* final-time objective (qb_engine_set_objective) */
  ObjArgs obj;
  double2* d_obj_target = nullptr;
  double *d_obj_purity = nullptr, *d_obj_weight = nullptr, *d_obj_out = nullptr, *d_obj_sum = nullptr;
  int obj_ninit = 0;
  void* nccl_comm = nullptr;    /* ncclComm_t, set by qb_engine_comm_init */
  int comm_size = 1;
  double* d_dp_jbar = nullptr;
  double* d_dp_out = nullptr;
  int dp_ninit = 0;
  size_t bsmem_fwd = 0, bsmem_adj = 0, bsmem_app = 0;
  /* qubit-system kernels (qb_qubit.cuh): used when available for this system */
  bool use_q = false;
  /* variants by elements per thread (E = 4, 2, 1 for four qubits): the fewer initial conditions a sweep has, the more
   * threads each state is spread over (a sweep is a chain of dependent CTA-wide phases; with few CTAs per SM more warps per
   * state shorten the chain, with many CTAs per SM fewer instructions per state win) */
  int nqv = 0;
  QKernelSet qv[3];
  size_t qv_smem_fwd[3] = {0, 0, 0}, qv_smem_adj[3] = {0, 0, 0};
  int qv_force = -1;
  int nsm = 0;
  double* d_rec = nullptr;
  double* d_hh = nullptr;      /* qubit kernels, fallback solver: Hessenberg columns [ninit][8][mcap^2] */
  int recstride = 0, rec_gcs = 0, rec_gval = 0, rec_gidx = 0, nvec = 0;
  unsigned rec_fwd_bytes = 0, rec_bytes = 0;
  size_t smem_fwd = 0, smem_adj = 0, smem_app = 0;
  /* device copies of the control description */
  qb_segment* d_segs = nullptr;
  double* d_carrier = nullptr;
  double* d_pipulse = nullptr;
  double* d_alpha = nullptr;
  int* d_upd = nullptr;
  CtrlArgs cargs;
  /* tables */
  double* d_tab = nullptr;
  int* d_gidx = nullptr;
  double* d_gval = nullptr;
  double* d_gcs = nullptr;
  bool params_set = false;
  /* state buffers, sized for ninit_cap */
  int ninit_cap = 0;
  double2* d_rho0 = nullptr;    /* initial states */
  double2* d_rhoT = nullptr;    /* final states of the last forward sweep (also read by the Schroedinger adjoint) */
  double2* d_rbar = nullptr;    /* terminal adjoint states */
  double2* d_rho0bar = nullptr; /* adjoint states at t=0 */
  double2* d_traj = nullptr;
  double2* d_stage_traj = nullptr;   /* stored stages (QB_FLAG_STORE_STAGE), sized with d_traj */
  bool store_stage = false;
  bool last_fwd_stage = false;
  size_t traj_cap = 0;          /* in double2 */
  size_t stage_cap = 0;
  double* d_grad_ic = nullptr;
  double* d_G = nullptr;        /* [nd + 8] */
  double* d_stage = nullptr;    /* host-layout staging [ninit_cap][2*dim] */
  unsigned long long* d_counters = nullptr;
  int last_fwd_ninit = -1;
  bool last_fwd_stored = false;   /* the adjoint sweep's primal input (states and/or stages) is on the device */
  bool last_fwd_states = false;   /* ... and the state trajectory x_0..x_ntime itself (penalties, qb_engine_get_state) */
  qb_stats stats;
  std::vector<double> h_carrier;
};

/* large-state Schroedinger kernels (qb_bigs_q*.cu) */
void qbig_pick_s1(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_s2(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_s3(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_s4(bool hasj, int dim, BKernelSet* ks);
void qbig_pick_s5(bool hasj, int dim, BKernelSet* ks);

extern "C" { static void free_penalty(qb_engine* e); static void free_objective(qb_engine* e); }

namespace {

const QKernelSet& q_variant(const qb_engine* e, int ninit, int* idx) {
  int v = 0;
  if (e->qv_force >= 0) v = e->qv_force;
  else if (e->nqv == 3) v = ninit > 2 * e->nsm ? 0 : 1;     /* measured on C4: E=2 up to two CTAs per SM, E=4 beyond; E=1 never wins */
  if (v >= e->nqv) v = e->nqv - 1;
  *idx = v;
  return e->qv[v];
}

int ensure_capacity(qb_engine* e, int ninit, bool need_traj, bool need_stage = false) {
  if (ninit > e->ninit_cap) {
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_rho0); cudaFree(e->d_rhoT); cudaFree(e->d_rbar); cudaFree(e->d_rho0bar); cudaFree(e->d_grad_ic); cudaFree(e->d_stage);
    e->d_rho0 = e->d_rhoT = e->d_rbar = e->d_rho0bar = nullptr; e->d_grad_ic = nullptr; e->d_stage = nullptr;
    e->ninit_cap = 0;
    size_t vb = (size_t)ninit * e->dimp * sizeof(double2);
    CUDA_TRY(cudaMalloc(&e->d_rho0, vb));
    CUDA_TRY(cudaMalloc(&e->d_rhoT, vb));
    CUDA_TRY(cudaMalloc(&e->d_rbar, vb));
    CUDA_TRY(cudaMalloc(&e->d_rho0bar, vb));
    CUDA_TRY(cudaMalloc(&e->d_grad_ic, (size_t)ninit * (e->S.nd > 0 ? e->S.nd : 1) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&e->d_stage, (size_t)ninit * 2 * e->S.dim * sizeof(double)));
    if (e->use_big) {
      cudaFree(e->d_scratch); e->d_scratch = nullptr;
      CUDA_TRY(cudaMalloc(&e->d_scratch, (size_t)ninit * (e->nvec + 3) * e->dimp * sizeof(double2)));
    }
    if (e->use_q) {
      cudaFree(e->d_scratch); cudaFree(e->d_hh); e->d_scratch = nullptr; e->d_hh = nullptr;
      CUDA_TRY(cudaMalloc(&e->d_scratch, (size_t)ninit * (e->mcap + 1) * e->dimp * sizeof(double2)));
      CUDA_TRY(cudaMalloc(&e->d_hh, (size_t)ninit * 8 * e->mcap * e->mcap * sizeof(double)));
    }
    e->ninit_cap = ninit;
  }
  if (need_traj) {
    size_t need = (size_t)ninit * (e->S.ntime + 1) * e->dimp;
    if (need > e->traj_cap) {
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      cudaFree(e->d_traj);
      e->d_traj = nullptr;
      e->traj_cap = 0;
      cudaError_t err = cudaMalloc(&e->d_traj, need * sizeof(double2));
      if (err != cudaSuccess) {
        return fail(QB_ERR_NOMEM, "trajectory storage of %.2f GB for %d initial conditions does not fit on the device: %s",
                    need * sizeof(double2) / 1e9, ninit, cudaGetErrorString(err));
      }
      e->traj_cap = need;
    }
  }
  if (need_stage) {
    size_t need = (size_t)ninit * (e->S.ntime + 1) * e->dimp;
    if (need > e->stage_cap) {
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      cudaFree(e->d_stage_traj);
      e->d_stage_traj = nullptr;
      e->stage_cap = 0;
      cudaError_t err = cudaMalloc(&e->d_stage_traj, need * sizeof(double2));
      if (err != cudaSuccess)
        return fail(QB_ERR_NOMEM, "stage storage of %.2f GB (QB_FLAG_STORE_STAGE) does not fit on the device: %s", need * sizeof(double2) / 1e9, cudaGetErrorString(err));
      e->stage_cap = need;
    }
  }
  return QB_OK;
}

/* overlap mode of the penalty: table and value from the trajectory of the last forward sweep, once per sweep */
int ensure_overlap_table(qb_engine* e, int ninit) {
  if (e->ov_valid) return QB_OK;
  if (!e->last_fwd_states || !e->d_traj) return fail(QB_ERR_STATE, "the Jtrace penalty of the Schroedinger solver needs the stored trajectory: call forward with store_trajectory=1");
  overlap_table_kernel<<<ninit, 256, 0, e->stream>>>(e->d_traj, e->S.ntime, e->S.dim, e->dimp, e->d_pen_a, e->d_pen_purity, e->d_pen_wdt, e->d_pen_ov, e->d_pen_ovval);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  e->ov_valid = true;
  return QB_OK;
}

int launch_pack(qb_engine* e, const double* host, double2* dev, int n) {
  size_t hb = (size_t)n * 2 * e->S.dim * sizeof(double);
  CUDA_TRY(cudaMemcpyAsync(e->d_stage, host, hb, cudaMemcpyHostToDevice, e->stream));
  size_t tot = (size_t)n * e->dimp;
  pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(e->d_stage, dev, e->S.dim, e->dimp, n);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return QB_OK;
}
int launch_unpack(qb_engine* e, const double2* dev, double* host, int n) {
  size_t tot = (size_t)n * e->dimp;
  unpack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(dev, e->d_stage, e->S.dim, e->dimp, n);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  size_t hb = (size_t)n * 2 * e->S.dim * sizeof(double);
  CUDA_TRY(cudaMemcpyAsync(host, e->d_stage, hb, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

qq::QArgs make_qargs(const qb_engine* e, const SweepArgs& A) {
  qq::QArgs Q;
  memset(&Q, 0, sizeof(Q));
  Q.rec = e->d_rec; Q.recstride = e->recstride; Q.rec_fwd_bytes = e->rec_fwd_bytes; Q.rec_bytes = e->rec_bytes;
  Q.rec_gcs = e->rec_gcs; Q.rec_gval = e->rec_gval; Q.rec_gidx = e->rec_gidx;
  Q.upd = A.upd; Q.state = A.state; Q.state_out = A.state_out; Q.final_state = A.final_state; Q.traj = A.traj;
  Q.stage_traj = nullptr;
  Q.scratch = e->d_scratch; Q.hh = e->d_hh;
  Q.pen = A.pen;
  Q.grad_ic = A.grad_ic; Q.counters = A.counters; Q.ninit = A.ninit; Q.mcap = A.mcap;
  return Q;
}

qbig::BArgs make_bargs(const qb_engine* e, const SweepArgs& A) {
  qbig::BArgs B;
  memset(&B, 0, sizeof(B));
  B.tab = A.tab; B.gidx = A.gidx; B.gval = A.gval; B.gcs = A.gcs; B.upd = A.upd;
  B.state = A.state; B.state_out = A.state_out; B.final_state = A.final_state; B.traj = A.traj; B.stage_traj = nullptr;
  B.pen = A.pen;
  B.scratch = e->d_scratch; B.grad_ic = A.grad_ic; B.counters = A.counters;
  B.ninit = A.ninit; B.mcap = A.mcap; B.nvec = e->nvec; B.store = A.store; B.dimp = e->dimp; B.L = e->maxlev;
  return B;
}

int resolve_timings(qb_engine* e) {
  for (auto& t : e->pending) {
    float ms = 0.f;
    CUDA_TRY(cudaEventSynchronize(t.b));
    CUDA_TRY(cudaEventElapsedTime(&ms, t.a, t.b));
    if (t.kind == 0) { e->stats.last_fwd_ms = ms; e->stats.total_fwd_ms += ms; e->stats.n_fwd++; }
    else { e->stats.last_adj_ms = ms; e->stats.total_adj_ms += ms; e->stats.n_adj++; }
    cudaEventDestroy(t.a); cudaEventDestroy(t.b);
  }
  e->pending.clear();
  return QB_OK;
}
int timed_begin(qb_engine* e, int kind, qb_engine::Timed* t) {
  if (e->pending.size() >= 4096) { int rc = resolve_timings(e); if (rc) return rc; }
  t->kind = kind;
  CUDA_TRY(cudaEventCreate(&t->a));
  CUDA_TRY(cudaEventCreate(&t->b));
  CUDA_TRY(cudaEventRecord(t->a, e->stream));
  return QB_OK;
}
int timed_end(qb_engine* e, const qb_engine::Timed& t) {
  CUDA_TRY(cudaEventRecord(t.b, e->stream));
  e->pending.push_back(t);
  return QB_OK;
}

int read_counters(qb_engine* e, long long* iters, long long* evals) {
  unsigned long long h[2];
  CUDA_TRY(cudaMemcpyAsync(h, e->d_counters, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  *iters = (long long)h[0];
  *evals = (long long)h[1];
  return QB_OK;
}

}  // namespace

extern "C" {

const char* qb_last_error(void) { return g_err.c_str(); }

int qb_engine_create(const qb_system_desc* d, qb_engine** out) {
  if (!d || !out) return fail(QB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->nosc < 1 || d->nosc > kMaxOsc) return fail(QB_ERR_UNSUPPORTED, "matrix-free engine supports 1..%d oscillators, got %d", kMaxOsc, d->nosc);
  if (d->ntime < 1) return fail(QB_ERR_INVALID, "ntime must be >= 1");
  if (d->stepper != QB_STEPPER_IMR && d->stepper != QB_STEPPER_IMR4 && d->stepper != QB_STEPPER_IMR8 && d->stepper != QB_STEPPER_EE)
    return fail(QB_ERR_UNSUPPORTED, "timestepper must be IMR, IMR4, IMR8 or EE (got %d)", d->stepper);
  if (d->linsolve_type != QB_LINSOLVE_GMRES && d->linsolve_type != QB_LINSOLVE_NEUMANN) return fail(QB_ERR_INVALID, "unknown linear solver type %d", d->linsolve_type);
  if (d->linsolve_maxiter < 0) return fail(QB_ERR_INVALID, "linearsolver_maxiter must be >= 0");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(QB_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  if (d->device < 0 || d->device >= ndev) return fail(QB_ERR_INVALID, "device %d out of range (%d devices)", d->device, ndev);
  CUDA_TRY(cudaSetDevice(d->device));

  qb_engine* e = new qb_engine();
  memset(&e->S, 0, sizeof(e->S));
  memset(&e->pen, 0, sizeof(e->pen));
  memset(&e->stats, 0, sizeof(e->stats));
  DevSys& S = e->S;
  S.Q = d->nosc;
  S.N = 1;
  bool all2 = true;
  for (int k = 0; k < S.Q; k++) {
    if (d->nlevels[k] < 1) { delete e; return fail(QB_ERR_INVALID, "nlevels[%d] = %d", k, d->nlevels[k]); }
    S.nlev[k] = d->nlevels[k];
    S.N *= d->nlevels[k];
    if (d->nlevels[k] != 2) all2 = false;
  }
  S.lind = d->lindblad;
  e->lind = d->lindblad != QB_LINDBLAD_NONE;
  long long dim = e->lind ? (long long)S.N * S.N : S.N;
  if (dim > 4096) {
    delete e;
    return fail(QB_ERR_UNSUPPORTED, "state dimension %lld is not supported (the vectorised state must fit one SM: dim = N^2 <= 4096 for Lindblad, N <= 4096 for Schroedinger)", dim);
  }
  S.dim = (int)dim;
  S.npairs = S.Q * (S.Q - 1) / 2;
  int s = 1;
  for (int k = S.Q - 1; k >= 0; k--) { S.stride[k] = s; s *= S.nlev[k]; }
  const bool addT1 = d->lindblad == QB_LINDBLAD_DECAY || d->lindblad == QB_LINDBLAD_BOTH;
  const bool addT2 = d->lindblad == QB_LINDBLAD_DEPHASE || d->lindblad == QB_LINDBLAD_BOTH;
  for (int k = 0; k < S.Q; k++) {
    S.detuning[k] = d->detuning[k];
    S.selfkerr[k] = d->selfkerr[k];
    S.decay[k] = (d->decay_time[k] > 1e-14 && addT1) ? 1. / d->decay_time[k] : 0.0;     /* src/mastereq.cpp:2070-2077 */
    S.dephase[k] = (d->dephase_time[k] > 1e-14 && addT2) ? 1. / d->dephase_time[k] : 0.0;
  }
  bool hasj = false;
  for (int id = 0; id < S.npairs; id++) {
    S.crosskerr[id] = d->crosskerr[id];
    S.Jkl[id] = d->Jkl[id];
    S.eta[id] = d->eta[id];
    if (fabs(d->Jkl[id]) > 1e-10) hasj = true;
  }
  /* compositional steppers = implicit midpoint on a refined, non-uniform grid of ntime*nstage sub-steps */
  S.nstage = 1;
  S.gamma[0] = 1.0;
  if (d->stepper == QB_STEPPER_IMR8) {
    static const double g8[15] = {0.74167036435061295344822780, -0.40910082580003159399730010, 0.19075471029623837995387626, -0.57386247111608226665638773,
                                  0.29906418130365592384446354, 0.33462491824529818378495798, 0.31529309239676659663205666, -0.79688793935291635401978884,
                                  0.31529309239676659663205666, 0.33462491824529818378495798, 0.29906418130365592384446354, -0.57386247111608226665638773,
                                  0.19075471029623837995387626, -0.40910082580003159399730010, 0.74167036435061295344822780};
    S.nstage = 15;
    for (int i = 0; i < 15; i++) S.gamma[i] = g8[i];
  } else if (d->stepper == QB_STEPPER_IMR4) {
    S.nstage = 3;
    S.gamma[0] = 1. / (2. - pow(2., 1. / 3.));
    S.gamma[1] = -pow(2., 1. / 3.) * S.gamma[0];
    S.gamma[2] = 1. / (2. - pow(2., 1. / 3.));
  }
  e->ntime_user = d->ntime;
  S.ntime = d->ntime * S.nstage;
  S.ee = d->stepper == QB_STEPPER_EE ? 1 : 0;
  { const char* sb = getenv("QB_SPRED_BIAS"); S.spred_bias = sb ? atoi(sb) : 0; }
  S.dt = d->total_time / d->ntime;
  S.linsolve = d->linsolve_type;
  S.maxiter = d->linsolve_maxiter;
  S.abstol = d->linsolve_abstol;
  S.reltol = d->linsolve_reltol;
  S.flags = d->flags;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) S.sq8[i * 8 + j] = sqrt((double)(i * j));
  { const char* ar = getenv("QB_ARNOLDI"); if (ar && atoi(ar) != 0) S.flags |= QB_FLAG_ARNOLDI; }   /* verification switch, see QB_FLAG_ARNOLDI */
  /* parameters and carriers */
  int nseg_total = d->seg_off[S.Q], ncar_total = d->carrier_off[S.Q], npp_total = d->pipulse_off[S.Q];
  S.ncar_total = ncar_total;
  S.osc_off[0] = 0;
  std::vector<int> upd;
  for (int k = 0; k < S.Q; k++) {
    S.car_off[k] = d->carrier_off[k];
    int nc = d->carrier_off[k + 1] - d->carrier_off[k];
    int np = 0;
    for (int si = d->seg_off[k]; si < d->seg_off[k + 1]; si++) {
      const qb_segment& sg = d->segments[si];
      if (sg.type != QB_CTRL_BSPLINE2ND && sg.type != QB_CTRL_BSPLINE0) { delete e; return fail(QB_ERR_UNSUPPORTED, "control segment type %d has no device implementation", sg.type); }
      if (sg.type == QB_CTRL_BSPLINE2ND && sg.nsplines < 3) { delete e; return fail(QB_ERR_INVALID, "BSpline2nd needs at least 3 splines"); }
      np += 2 * sg.nsplines * nc;
    }
    S.osc_off[k + 1] = S.osc_off[k] + np;
    if (np > 0)
      for (int f = 0; f < nc; f++)
        for (int j = 0; j < 3; j++) { upd.push_back(k); upd.push_back(d->carrier_off[k] + f); upd.push_back(f); upd.push_back(j); }
  }
  S.car_off[S.Q] = ncar_total;
  S.nd = S.osc_off[S.Q];
  S.nupd = (int)upd.size() / 4;
  S.tabstride = 1 + 2 * S.Q + 2 * S.npairs;
  KernelSet ks;
  memset(&ks, 0, sizeof(ks));
  for (int k = 0; k < S.Q; k++) e->maxlev = S.nlev[k] > e->maxlev ? S.nlev[k] : e->maxlev;
  if (S.dim > kMaxThreads && S.ee) { delete e; return fail(QB_ERR_UNSUPPORTED, "timestepper = EE is implemented for state dimensions up to %d (got %d)", kMaxThreads, S.dim); }
  if (S.dim > kMaxThreads) {
    if (!e->lind) switch (S.Q) {     /* Schroedinger solver, 256 < N <= 4096 (qb_bigs_q*.cu) */
      case 1: qbig_pick_s1(hasj, S.dim, &e->bk); break;
      case 2: qbig_pick_s2(hasj, S.dim, &e->bk); break;
      case 3: qbig_pick_s3(hasj, S.dim, &e->bk); break;
      case 4: qbig_pick_s4(hasj, S.dim, &e->bk); break;
      default: qbig_pick_s5(hasj, S.dim, &e->bk); break;
    }
    else switch (S.Q) {
      case 1: qbig_pick_q1(hasj, S.dim, &e->bk); break;
      case 2: qbig_pick_q2(hasj, S.dim, &e->bk); break;
      case 3: qbig_pick_q3(hasj, S.dim, &e->bk); break;
      case 4: qbig_pick_q4(hasj, S.dim, &e->bk); break;
      default: qbig_pick_q5(hasj, S.dim, &e->bk); break;
    }
    if (!e->bk.fwd || e->maxlev > 31) { delete e; return fail(QB_ERR_UNSUPPORTED, "no large-state kernel for dim %d / %d levels", S.dim, e->maxlev); }
    e->use_big = true;
    e->dimp = e->bk.dimp;
  } else {
    pick(S.Q, all2, hasj, e->lind, S.dim, &ks);
    if (!ks.fwd) { delete e; return fail(QB_ERR_UNSUPPORTED, "no small-N kernel for this system size (dim %d)", S.dim); }
    e->fwd = ks.fwd; e->adj = ks.adj; e->app = ks.app;
    e->dimp = ks.dimp;
  }
  if (S.nupd > e->dimp || S.nupd > kMaxUpd) { delete e; return fail(QB_ERR_UNSUPPORTED, "too many carrier waves (%d updater threads) for a CTA of %d threads", S.nupd, e->dimp); }
  const bool gm = d->linsolve_type == QB_LINSOLVE_GMRES;
  e->mcap = gm ? (S.maxiter < kRestart ? (S.maxiter > 0 ? S.maxiter : 1) : kRestart) : 1;
  e->nvec = e->mcap > 2 ? e->mcap : 2;
  if (e->use_big && e->nvec < qbig::kBMax + 1) e->nvec = qbig::kBMax + 1;   /* the one-reduction solver keeps u_1..u_kBMax there */
  /* qubit-system kernels: all oscillators two-level, no Jkl coupling (QB_KERNELS=generic forces the general ones,
   * QB_QUBIT_E=4|8 picks the elements per thread for 4 qubits) */
  {
    const char* force = getenv("QB_KERNELS");
    const char* epe = getenv("QB_QUBIT_E");
    if (!e->use_big && all2 && !hasj && !S.ee && !(force && strcmp(force, "generic") == 0)) {
      qq_pick(S.Q, e->lind, 0, &e->qv[0]);
      e->use_q = e->qv[0].fwd[0] != nullptr && e->qv[0].dim == S.dim;
      e->nqv = e->use_q ? 1 : 0;
      if (e->use_q && S.Q == 4) {
        qq_pick(S.Q, e->lind, 2, &e->qv[1]);
        qq_pick(S.Q, e->lind, 1, &e->qv[2]);
        e->nqv = 3;
        if (epe) { const int E = atoi(epe); e->qv_force = E == 1 ? 2 : (E == 2 ? 1 : 0); }
      }
    }
  }
  if (e->use_q) {
    e->rec_gcs = S.tabstride;
    e->rec_gval = e->rec_gcs + 2 * S.ncar_total;
    e->rec_gidx = e->rec_gval + 3 * S.Q;
    e->recstride = (e->rec_gidx + 2 * S.Q + 1) & ~1;
    e->rec_bytes = (unsigned)e->recstride * 8u;
    e->rec_fwd_bytes = (unsigned)(((1 + 2 * S.Q) * 8 + 15) & ~15);
    e->store_stage = e->lind && (S.flags & QB_FLAG_STORE_STAGE) != 0;
    for (int v = 0; v < e->nqv; v++) {
      e->qv_smem_fwd[v] = qq_smem(e->qv[v], false, e->store_stage, e->rec_fwd_bytes, 0);
      e->qv_smem_adj[v] = qq_smem(e->qv[v], true, false, e->rec_bytes, S.nd);
    }
  }
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
  if (e->use_big) {
    e->bsmem_fwd = qbig_smem(e->bk, e->mcap, e->maxlev, false, 0);
    e->bsmem_adj = qbig_smem(e->bk, e->mcap, e->maxlev, true, S.nd);
    e->bsmem_app = qbig_smem(e->bk, 1, e->maxlev, false, 0);
    if (e->bsmem_adj > (size_t)prop.sharedMemPerBlockOptin) {
      size_t need = e->bsmem_adj;
      delete e;
      return fail(QB_ERR_UNSUPPORTED, "shared memory per initial condition (%zu B) exceeds the device limit (%zu B)", need, (size_t)prop.sharedMemPerBlockOptin);
    }
    CUDA_TRY(cudaFuncSetAttribute((const void*)e->bk.fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bsmem_fwd));
    CUDA_TRY(cudaFuncSetAttribute((const void*)e->bk.adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bsmem_adj));
    CUDA_TRY(cudaFuncSetAttribute((const void*)e->bk.app, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bsmem_app));
    e->store_stage = (S.flags & QB_FLAG_STORE_STAGE) != 0;
  } else {
  e->smem_fwd = smem_bytes(ks, e->mcap, 0);
  e->smem_adj = smem_bytes(ks, e->mcap, S.nd);
  e->smem_app = smem_bytes(ks, 1, 0);
  }
  if (!e->use_big && e->smem_adj > (size_t)prop.sharedMemPerBlockOptin) {
    size_t need = e->smem_adj;
    delete e;
    return fail(QB_ERR_UNSUPPORTED, "shared memory per initial condition (%zu B) exceeds the device limit (%zu B): large-N path not implemented yet", need,
                (size_t)prop.sharedMemPerBlockOptin);
  }
  if (!e->use_big && !e->use_q) e->store_stage = e->lind && !S.ee && (S.flags & QB_FLAG_STORE_STAGE) != 0;     /* general kernels: Lindblad sweeps */
  if ((S.flags & QB_FLAG_STORE_STAGE) && !e->store_stage) S.flags |= QB_FLAG_RECOMPUTE_STAGE;   /* same values, re-solved */
  if (!e->use_big) {
    CUDA_TRY(cudaFuncSetAttribute((const void*)e->fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_fwd));
    CUDA_TRY(cudaFuncSetAttribute((const void*)e->adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_adj));
    CUDA_TRY(cudaFuncSetAttribute((const void*)e->app, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_app));
  }
  if (e->use_q) {
    e->nsm = prop.multiProcessorCount;
    for (int v = 0; v < e->nqv; v++) {
      if (e->qv_smem_adj[v] > (size_t)prop.sharedMemPerBlockOptin || e->qv_smem_fwd[v] > (size_t)prop.sharedMemPerBlockOptin) {
        delete e;
        return fail(QB_ERR_UNSUPPORTED, "shared memory of the qubit-system kernels exceeds the device limit");
      }
      for (int g = 0; g < 2; g++) {
        CUDA_TRY(cudaFuncSetAttribute((const void*)e->qv[v].fwd[g], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->qv_smem_fwd[v]));
        CUDA_TRY(cudaFuncSetAttribute((const void*)e->qv[v].adj[g], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->qv_smem_adj[v]));
        CUDA_TRY(cudaFuncSetAttribute((const void*)e->qv[v].adj_stored[g], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->qv_smem_adj[v]));
      }
      CUDA_TRY(cudaFuncSetAttribute((const void*)e->qv[v].app, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(e->qv[v].vec_bytes + 256)));
    }
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  e->device = d->device;

  CUDA_TRY(cudaMalloc(&e->d_segs, sizeof(qb_segment) * (nseg_total > 0 ? nseg_total : 1)));
  if (nseg_total > 0) CUDA_TRY(cudaMemcpy(e->d_segs, d->segments, sizeof(qb_segment) * nseg_total, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&e->d_carrier, sizeof(double) * (ncar_total > 0 ? ncar_total : 1)));
  if (ncar_total > 0) CUDA_TRY(cudaMemcpy(e->d_carrier, d->carrier_freq, sizeof(double) * ncar_total, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&e->d_pipulse, sizeof(double) * 3 * (npp_total > 0 ? npp_total : 1)));
  if (npp_total > 0) CUDA_TRY(cudaMemcpy(e->d_pipulse, d->pipulse, sizeof(double) * 3 * npp_total, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&e->d_alpha, sizeof(double) * (S.nd > 0 ? S.nd : 1)));
  CUDA_TRY(cudaMemset(e->d_alpha, 0, sizeof(double) * (S.nd > 0 ? S.nd : 1)));
  CUDA_TRY(cudaMalloc(&e->d_upd, sizeof(int) * (upd.size() > 0 ? upd.size() : 4)));
  if (!upd.empty()) CUDA_TRY(cudaMemcpy(e->d_upd, upd.data(), sizeof(int) * upd.size(), cudaMemcpyHostToDevice));
  e->cargs.segs = e->d_segs;
  e->cargs.carrier = e->d_carrier;
  e->cargs.pipulse = e->d_pipulse;
  e->cargs.alpha = e->d_alpha;
  for (int k = 0; k <= S.Q; k++) { e->cargs.seg_off[k] = d->seg_off[k]; e->cargs.pp_off[k] = d->pipulse_off[k]; }

  CUDA_TRY(cudaMalloc(&e->d_tab, sizeof(double) * (size_t)(S.ntime + 1) * S.tabstride));
  CUDA_TRY(cudaMalloc(&e->d_gidx, sizeof(int) * (size_t)(S.ntime + 1) * S.Q * 4));
  CUDA_TRY(cudaMalloc(&e->d_gval, sizeof(double) * (size_t)(S.ntime + 1) * S.Q * 3));
  CUDA_TRY(cudaMalloc(&e->d_gcs, sizeof(double) * (size_t)(S.ntime + 1) * (ncar_total > 0 ? ncar_total : 1) * 2));
  if (e->use_q) CUDA_TRY(cudaMalloc(&e->d_rec, sizeof(double) * (size_t)S.ntime * e->recstride));
  CUDA_TRY(cudaMalloc(&e->d_G, sizeof(double) * (S.nd + 8)));
  CUDA_TRY(cudaMemset(e->d_G, 0, sizeof(double) * (S.nd + 8)));
  CUDA_TRY(cudaMalloc(&e->d_counters, 2 * sizeof(unsigned long long)));
  *out = e;
  return QB_OK;
}

int qb_engine_destroy(qb_engine* e) {
  if (!e) return QB_OK;
  cudaSetDevice(e->device);
  free_penalty(e);
  free_objective(e);
  if (e->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->nccl_comm);
  cudaFree(e->d_dp_jbar); cudaFree(e->d_dp_out);
  cudaFree(e->d_segs); cudaFree(e->d_carrier); cudaFree(e->d_pipulse); cudaFree(e->d_alpha); cudaFree(e->d_upd);
  cudaFree(e->d_tab); cudaFree(e->d_gidx); cudaFree(e->d_gval); cudaFree(e->d_gcs);
  cudaFree(e->d_rho0); cudaFree(e->d_rhoT); cudaFree(e->d_rbar); cudaFree(e->d_rho0bar); cudaFree(e->d_traj); cudaFree(e->d_stage_traj); cudaFree(e->d_grad_ic); cudaFree(e->d_G);
  cudaFree(e->d_stage); cudaFree(e->d_counters); cudaFree(e->d_rec); cudaFree(e->d_scratch); cudaFree(e->d_hh);
  for (auto& t : e->pending) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
  return QB_OK;
}

int qb_engine_ndesign(const qb_engine* e) { return e ? e->S.nd : -1; }
int qb_engine_dim(const qb_engine* e) { return e ? e->S.dim : -1; }

int qb_engine_set_params(qb_engine* e, const double* alpha, int ndesign) {
  if (!e || (!alpha && ndesign > 0)) return fail(QB_ERR_INVALID, "null argument");
  if (ndesign != e->S.nd) return fail(QB_ERR_INVALID, "design vector has %d entries, engine expects %d", ndesign, e->S.nd);
  if (ndesign > 0) CUDA_TRY(cudaMemcpyAsync(e->d_alpha, alpha, sizeof(double) * ndesign, cudaMemcpyHostToDevice, e->stream));
  ctrl_table_kernel<<<(e->S.ntime + 1 + 127) / 128, 128, 0, e->stream>>>(e->S, e->cargs, e->d_tab, e->d_gidx, e->d_gval, e->d_gcs);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  if (e->use_q) {
    pack_rec_kernel<<<(e->S.ntime + 127) / 128, 128, 0, e->stream>>>(e->S, e->d_tab, e->d_gidx, e->d_gval, e->d_gcs, e->d_rec, e->recstride, e->rec_gcs,
                                                                    e->rec_gval, e->rec_gidx);
    e->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  e->params_set = true;
  return QB_OK;
}

int qb_engine_eval_controls(qb_engine* e, const double* t, int nt, double* out_p, double* out_q) {
  if (!e || !t || !out_p || !out_q || nt < 0) return fail(QB_ERR_INVALID, "bad argument");
  if (!e->params_set) return fail(QB_ERR_STATE, "qb_engine_set_params has not been called");
  if (nt == 0) return QB_OK;
  double *d_t = nullptr, *d_rows = nullptr;
  CUDA_TRY(cudaMalloc(&d_t, sizeof(double) * nt));
  CUDA_TRY(cudaMalloc(&d_rows, sizeof(double) * (size_t)nt * e->S.tabstride));
  CUDA_TRY(cudaMemcpyAsync(d_t, t, sizeof(double) * nt, cudaMemcpyHostToDevice, e->stream));
  ctrl_eval_kernel<<<(nt + 127) / 128, 128, 0, e->stream>>>(e->S, e->cargs, d_t, nt, d_rows);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  std::vector<double> rows((size_t)nt * e->S.tabstride);
  CUDA_TRY(cudaMemcpyAsync(rows.data(), d_rows, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (int n = 0; n < nt; n++)
    for (int k = 0; k < e->S.Q; k++) {
      out_p[(size_t)n * e->S.Q + k] = rows[(size_t)n * e->S.tabstride + 1 + k];
      out_q[(size_t)n * e->S.Q + k] = rows[(size_t)n * e->S.tabstride + 1 + e->S.Q + k];
    }
  cudaFree(d_t); cudaFree(d_rows);
  return QB_OK;
}

int qb_engine_upload_states(qb_engine* e, const double* host_states, int ninit, int which) {
  if (!e || !host_states || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (which != 0 && which != 1) return fail(QB_ERR_INVALID, "which must be 0 (rho0) or 1 (rhoT_bar)");
  if (which == 1 && ninit > e->ninit_cap) return fail(QB_ERR_STATE, "terminal adjoint states for %d initial conditions before a forward sweep of that size", ninit);
  int rc = ensure_capacity(e, ninit, false);
  if (rc) return rc;
  return launch_pack(e, host_states, which == 0 ? e->d_rho0 : e->d_rbar, ninit);
}

int qb_engine_forward_dev(qb_engine* e, int ninit, int store_trajectory) {
  if (!e || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (!e->params_set) return fail(QB_ERR_STATE, "qb_engine_set_params has not been called");
  if (store_trajectory < 0 || store_trajectory > 2) return fail(QB_ERR_INVALID, "store_trajectory must be 0, 1 or 2");
  /* 2 = keep only what the adjoint sweep reads: the stages under QB_FLAG_STORE_STAGE (qubit-system kernels), else the states */
  const bool stages_only = store_trajectory == 2 && e->store_stage && e->use_q;
  const bool keep_states = store_trajectory != 0 && !stages_only;
  const bool keep_stages = store_trajectory != 0 && e->store_stage;
  int rc = ensure_capacity(e, ninit, keep_states, keep_stages);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, 2 * sizeof(unsigned long long), e->stream));
  SweepArgs A;
  memset(&A, 0, sizeof(A));
  A.tab = e->d_tab; A.gidx = e->d_gidx; A.gval = e->d_gval; A.gcs = e->d_gcs; A.upd = e->d_upd;
  A.state = e->d_rho0; A.state_out = e->d_rhoT; A.final_state = nullptr;
  A.traj = keep_states ? e->d_traj : nullptr;
  A.grad_ic = e->d_grad_ic; A.counters = e->d_counters;
  A.ninit = ninit; A.mcap = e->mcap; A.store = keep_states ? 1 : 0;
  qb_engine::Timed tm;
  int rt = timed_begin(e, 0, &tm);
  if (rt) return rt;
  if (e->use_big) {
    qbig::BArgs BA = make_bargs(e, A);
    if (e->store_stage && store_trajectory) BA.stage_traj = e->d_stage_traj;
    e->bk.fwd<<<ninit, e->bk.threads, e->bsmem_fwd, e->stream>>>(e->S, BA);
  } else if (e->use_q) {
    qq::QArgs QA = make_qargs(e, A);
    if (keep_stages) QA.stage_traj = e->d_stage_traj;
    int vi;
    const QKernelSet& qk = q_variant(e, ninit, &vi);
    qk.fwd[e->S.linsolve == QB_LINSOLVE_GMRES ? 0 : 1]<<<ninit, qk.threads, e->qv_smem_fwd[vi], e->stream>>>(e->S, QA);
  } else {
    if (keep_stages) A.stage_traj = e->d_stage_traj;
    e->fwd<<<ninit, e->dimp, e->smem_fwd, e->stream>>>(e->S, A);
  }
  CUDA_TRY(cudaGetLastError());
  rt = timed_end(e, tm);
  if (rt) return rt;
  e->stats.kernel_launches++;
  /* d_rhoT stays untouched until the next forward sweep: the Schroedinger adjoint recomputes the primal backwards
   * from it (src/optimproblem.cpp:402) */
  e->last_fwd_ninit = ninit;
  e->last_fwd_stored = store_trajectory != 0;
  e->last_fwd_states = keep_states;
  e->ov_valid = false;
  e->stats.steps_fwd = (long long)ninit * e->ntime_user;
  e->stats.steps_adj = 0;
  e->stats.rhs_evals = -1;  /* resolved lazily in get_stats */
  return QB_OK;
}

int qb_engine_adjoint_dev(qb_engine* e, int ninit) {
  if (!e || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (e->last_fwd_ninit != ninit) return fail(QB_ERR_STATE, "adjoint sweep for %d initial conditions must follow a forward sweep of the same batch (last: %d)", ninit, e->last_fwd_ninit);
  if (e->lind && !e->last_fwd_stored) return fail(QB_ERR_STATE, "Lindblad adjoint needs the forward trajectory: call forward with store_trajectory=1");
  SweepArgs A;
  memset(&A, 0, sizeof(A));
  A.tab = e->d_tab; A.gidx = e->d_gidx; A.gval = e->d_gval; A.gcs = e->d_gcs; A.upd = e->d_upd;
  A.state = e->d_rbar; A.state_out = e->d_rho0bar; A.final_state = e->d_rhoT;
  A.traj = (e->lind && e->last_fwd_states) ? e->d_traj : nullptr;
  A.grad_ic = e->d_grad_ic; A.counters = e->d_counters;
  A.ninit = ninit; A.mcap = e->mcap; A.store = 0;
  if (e->pen.on) {
    if (e->pen_ninit != ninit) return fail(QB_ERR_STATE, "penalty was installed for %d initial conditions, adjoint sweep has %d", e->pen_ninit, ninit);
    if (e->lind && !e->last_fwd_states) return fail(QB_ERR_STATE, "penalty terms need the stored state trajectory: call forward with store_trajectory=1");
    if (e->pen_overlap) {
      int ro = ensure_overlap_table(e, ninit);
      if (ro) return ro;
    }
    A.pen = e->pen;
  }
  if (e->dp_ninit > 0) {
    if (e->dp_ninit != ninit) return fail(QB_ERR_STATE, "dpdm penalty was installed for %d initial conditions, adjoint sweep has %d", e->dp_ninit, ninit);
    if (!e->last_fwd_states) return fail(QB_ERR_STATE, "the dpdm penalty needs the stored trajectory: call forward with store_trajectory=1");
    A.pen.dp_traj = e->d_traj; A.pen.dp_jbar = e->d_dp_jbar; A.pen.dp_dtinv = 1.0 / (e->S.dt * e->S.dt * e->S.dt * e->S.dt); A.pen.dp_on = 1;
  }
  qb_engine::Timed tm;
  int rt = timed_begin(e, 1, &tm);
  if (rt) return rt;
  if (e->use_big) {
    qbig::BArgs BA = make_bargs(e, A);
    if (e->store_stage && e->last_fwd_stored) BA.stage_traj = e->d_stage_traj;
    e->bk.adj<<<ninit, e->bk.threads, e->bsmem_adj, e->stream>>>(e->S, BA);
  } else if (e->use_q) {
    qq::QArgs QA = make_qargs(e, A);
    if (e->store_stage && e->last_fwd_stored) QA.stage_traj = e->d_stage_traj;
    const bool st = QA.stage_traj != nullptr;
    int vi;
    const QKernelSet& qk = q_variant(e, ninit, &vi);
    const int gi = e->S.linsolve == QB_LINSOLVE_GMRES ? 0 : 1;
    (st ? qk.adj_stored[gi] : qk.adj[gi])<<<ninit, qk.threads, e->qv_smem_adj[vi], e->stream>>>(e->S, QA);
  } else {
    if (e->store_stage && e->last_fwd_stored) A.stage_traj = e->d_stage_traj;
    e->adj<<<ninit, e->dimp, e->smem_adj, e->stream>>>(e->S, A);
  }
  CUDA_TRY(cudaGetLastError());
  rt = timed_end(e, tm);
  if (rt) return rt;
  e->stats.kernel_launches++;
  if (e->S.nd > 0) {
    grad_reduce_kernel<<<(e->S.nd + 127) / 128, 128, 0, e->stream>>>(e->d_grad_ic, ninit, e->S.nd, e->d_G);
    CUDA_TRY(cudaGetLastError());
    e->stats.kernel_launches++;
  }
  e->stats.steps_adj = (long long)ninit * e->ntime_user;
  e->stats.rhs_evals = -2;
  return QB_OK;
}

int qb_engine_download_states(qb_engine* e, double* host_states, int ninit, int which) {
  if (!e || !host_states || ninit < 1 || ninit > e->ninit_cap) return fail(QB_ERR_INVALID, "bad argument");
  if (which != 0 && which != 1) return fail(QB_ERR_INVALID, "which must be 0 (rhoT) or 1 (rho0_bar)");
  return launch_unpack(e, which == 0 ? e->d_rhoT : e->d_rho0bar, host_states, ninit);
}

int qb_engine_download_grad(qb_engine* e, double* redgrad) {
  if (!e || (!redgrad && e->S.nd > 0)) return fail(QB_ERR_INVALID, "bad argument");
  if (e->S.nd > 0) CUDA_TRY(cudaMemcpyAsync(redgrad, e->d_G, sizeof(double) * e->S.nd, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

void* qb_engine_grad_dev(qb_engine* e) { return e ? (void*)e->d_G : nullptr; }

int qb_engine_forward(qb_engine* e, const double* rho0, int ninit, int store_trajectory, double* rhoT) {
  if (!e || !rho0 || !rhoT) return fail(QB_ERR_INVALID, "null argument");
  int rc = qb_engine_upload_states(e, rho0, ninit, 0);
  if (rc) return rc;
  rc = qb_engine_forward_dev(e, ninit, store_trajectory);
  if (rc) return rc;
  return qb_engine_download_states(e, rhoT, ninit, 0);
}

int qb_engine_adjoint(qb_engine* e, const double* rhoT_bar, int ninit, double* redgrad, double* rho0_bar) {
  if (!e || !rhoT_bar || (!redgrad && e->S.nd > 0)) return fail(QB_ERR_INVALID, "null argument");
  if (e->last_fwd_ninit != ninit) return fail(QB_ERR_STATE, "adjoint sweep for %d initial conditions must follow a forward sweep of the same batch (last: %d)", ninit, e->last_fwd_ninit);
  int rc = qb_engine_upload_states(e, rhoT_bar, ninit, 1);
  if (rc) return rc;
  rc = qb_engine_adjoint_dev(e, ninit);
  if (rc) return rc;
  if (rho0_bar) { rc = qb_engine_download_states(e, rho0_bar, ninit, 1); if (rc) return rc; }
  return qb_engine_download_grad(e, redgrad);
}

int qb_engine_get_state(qb_engine* e, int iinit, int tindex, double* out) {
  if (!e || !out) return fail(QB_ERR_INVALID, "null argument");
  if (!e->last_fwd_states || !e->d_traj) return fail(QB_ERR_STATE, "time-stepper requested state at time index %d, but didn't store it", tindex);
  if (iinit < 0 || iinit >= e->last_fwd_ninit || tindex < 0 || tindex > e->ntime_user) return fail(QB_ERR_INVALID, "state (%d, %d) out of range", iinit, tindex);
  const double2* src = e->d_traj + ((size_t)iinit * (e->S.ntime + 1) + (size_t)tindex * e->S.nstage) * e->dimp;
  return launch_unpack(e, src, out, 1);
}

/* qb_dense.cu: tensor-core (DMMA) formulation of the operator for N = 64, and the fp64 tensor-pipe peak */
bool qb_dense_supported(const DevSys& S);
cudaError_t qb_dense_apply_launch(const DevSys& S, const double* d_row, const double2* x, double2* y, int dimp, int nbatch, int reps, cudaStream_t stream);
cudaError_t qb_dmma_peak(int device, double* tflops);

int qb_measure_fp64_mma_peak(int device, double* tflops) {
  if (!tflops) return fail(QB_ERR_INVALID, "null argument");
  CUDA_TRY(qb_dmma_peak(device, tflops));
  return QB_OK;
}

/* Times `reps` applications of M(t) inside ONE launch over `nbatch` device-resident states (the states of the last
 * qb_engine_upload_states(..., 1)), stencil kernel (dense = 0) or tensor-core kernel (dense = 1): the A/B measurement behind the
 * choice of the operator formulation for N = 64.  ms_out: device time of the launch (CUDA events). */
int qb_engine_apply_bench(qb_engine* e, double t, int nbatch, int reps, int dense, double* ms_out) {
  if (!e || !ms_out || nbatch < 1 || reps < 1 || reps > 4096) return fail(QB_ERR_INVALID, "bad argument");
  if (!e->params_set) return fail(QB_ERR_STATE, "qb_engine_set_params has not been called");
  if (nbatch > e->ninit_cap) return fail(QB_ERR_STATE, "upload %d states first (qb_engine_upload_states)", nbatch);
  if (dense && !qb_dense_supported(e->S)) return fail(QB_ERR_UNSUPPORTED, "the tensor-core operator exists for Lindblad systems with N = 64 only");
  if (!dense && !e->use_big) return fail(QB_ERR_UNSUPPORTED, "apply benchmark: large-state kernels only");
  double *d_t = nullptr, *d_row = nullptr;
  CUDA_TRY(cudaMalloc(&d_t, sizeof(double)));
  CUDA_TRY(cudaMalloc(&d_row, sizeof(double) * e->S.tabstride));
  CUDA_TRY(cudaMemcpyAsync(d_t, &t, sizeof(double), cudaMemcpyHostToDevice, e->stream));
  ctrl_eval_kernel<<<1, 32, 0, e->stream>>>(e->S, e->cargs, d_t, 1, d_row);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  CUDA_TRY(cudaEventRecord(e0, e->stream));
  if (dense) CUDA_TRY(qb_dense_apply_launch(e->S, d_row, e->d_rbar, e->d_rho0bar, e->dimp, nbatch, reps, e->stream));
  else e->bk.app<<<nbatch, e->bk.threads, e->bsmem_app, e->stream>>>(e->S, d_row, e->d_rbar, e->d_rho0bar, (reps - 1) << 8, e->dimp, e->maxlev);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(e1, e->stream));
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_t); cudaFree(d_row);
  e->stats.kernel_launches += 2;
  return QB_OK;
}

int qb_engine_rhs_apply(qb_engine* e, double t, int transpose, const double* x, double* y, int nbatch) {
  if (!e || !x || !y || nbatch < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (!e->params_set) return fail(QB_ERR_STATE, "qb_engine_set_params has not been called");
  int rc = ensure_capacity(e, nbatch, false);
  if (rc) return rc;
  double *d_t = nullptr, *d_row = nullptr;
  CUDA_TRY(cudaMalloc(&d_t, sizeof(double)));
  CUDA_TRY(cudaMalloc(&d_row, sizeof(double) * e->S.tabstride));
  CUDA_TRY(cudaMemcpyAsync(d_t, &t, sizeof(double), cudaMemcpyHostToDevice, e->stream));
  ctrl_eval_kernel<<<1, 32, 0, e->stream>>>(e->S, e->cargs, d_t, 1, d_row);
  e->stats.kernel_launches++;
  rc = launch_pack(e, x, e->d_rbar, nbatch);
  if (rc) return rc;
  const char* dn = getenv("QB_DENSE_APPLY");      /* verification switch: route the unit-test hook through the tensor-core operator */
  if (dn && atoi(dn) != 0 && !transpose && qb_dense_supported(e->S)) CUDA_TRY(qb_dense_apply_launch(e->S, d_row, e->d_rbar, e->d_rho0bar, e->dimp, nbatch, 1, e->stream));
  else if (e->use_big) e->bk.app<<<nbatch, e->bk.threads, e->bsmem_app, e->stream>>>(e->S, d_row, e->d_rbar, e->d_rho0bar, transpose, e->dimp, e->maxlev);
  else if (e->use_q) e->qv[0].app<<<nbatch, e->qv[0].threads, e->qv[0].vec_bytes + 256, e->stream>>>(e->S, d_row, e->d_rbar, e->d_rho0bar, transpose);
  else e->app<<<nbatch, e->dimp, e->smem_app, e->stream>>>(e->S, d_row, e->d_rbar, e->d_rho0bar, transpose);
  CUDA_TRY(cudaGetLastError());
  e->stats.kernel_launches++;
  rc = launch_unpack(e, e->d_rho0bar, y, nbatch);
  cudaFree(d_t); cudaFree(d_row);
  return rc;
}

static void free_penalty(qb_engine* e) {
  cudaFree(e->d_pen_a); cudaFree(e->d_pen_d); cudaFree(e->d_pen_g); cudaFree(e->d_pen_wdt); cudaFree(e->d_pen_jbar); cudaFree(e->d_pen_c0); cudaFree(e->d_pen_out); cudaFree(e->d_pen_gw);
  cudaFree(e->d_pen_ov); cudaFree(e->d_pen_purity); cudaFree(e->d_pen_ovval);
  e->d_pen_ov = nullptr; e->d_pen_purity = e->d_pen_ovval = nullptr;
  e->pen_overlap = false; e->ov_valid = false;
  e->d_pen_gw = nullptr;
  e->d_pen_a = e->d_pen_d = e->d_pen_g = nullptr;
  e->d_pen_wdt = e->d_pen_jbar = e->d_pen_c0 = e->d_pen_out = nullptr;
  memset(&e->pen, 0, sizeof(e->pen));
  e->pen_ninit = 0;
}

int qb_engine_set_penalty(qb_engine* e, const qb_penalty_desc* p) {
  if (!e) return fail(QB_ERR_INVALID, "null engine");
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  free_penalty(e);
  if (!p) return QB_OK;
  const bool ovl = p->overlap != 0;
  if (!p->a || !p->wdt || !p->jbar || p->ninit < 1 || (!ovl && (!p->c0 || !p->d))) return fail(QB_ERR_INVALID, "incomplete penalty descriptor");
  if (ovl && e->lind) return fail(QB_ERR_INVALID, "the overlap form of the penalty is the Schroedinger solver's Jtrace; Lindblad uses the linear form");
  const int n = p->ninit, dim = e->S.dim, dimp = e->dimp, nt = e->S.ntime;
  int rc = ensure_capacity(e, n, false);
  if (rc) return rc;
  CUDA_TRY(cudaMalloc(&e->d_pen_a, (size_t)n * dimp * sizeof(double2)));
  CUDA_TRY(cudaMalloc(&e->d_pen_d, (size_t)dimp * sizeof(double2)));
  CUDA_TRY(cudaMalloc(&e->d_pen_wdt, (size_t)(nt + 1) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&e->d_pen_jbar, (size_t)n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&e->d_pen_c0, (size_t)n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&e->d_pen_out, (size_t)n * sizeof(double)));
  rc = launch_pack(e, p->a, e->d_pen_a, n);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(e->stream));     /* d_stage is reused below */
  if (ovl) CUDA_TRY(cudaMemset(e->d_pen_d, 0, (size_t)dimp * sizeof(double2)));
  else {
    rc = launch_pack(e, p->d, e->d_pen_d, 1);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  }
  if (p->guard) {
    CUDA_TRY(cudaMalloc(&e->d_pen_g, (size_t)dimp * sizeof(double2)));
    rc = launch_pack(e, p->guard, e->d_pen_g, 1);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  }
  {
    std::vector<double> wsub((size_t)nt + 1, 0.0), gsub((size_t)nt + 1, 0.0);
    for (int j = 0; j <= nt; j++)
      if (j % e->S.nstage == 0) { wsub[j] = p->wdt[j / e->S.nstage]; gsub[j] = 1.0 / e->ntime_user; }
    CUDA_TRY(cudaMemcpy(e->d_pen_wdt, wsub.data(), (size_t)(nt + 1) * sizeof(double), cudaMemcpyHostToDevice));
    if (p->guard) {
      CUDA_TRY(cudaMalloc(&e->d_pen_gw, (size_t)(nt + 1) * sizeof(double)));
      CUDA_TRY(cudaMemcpy(e->d_pen_gw, gsub.data(), (size_t)(nt + 1) * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  CUDA_TRY(cudaMemcpy(e->d_pen_jbar, p->jbar, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  if (ovl) {
    CUDA_TRY(cudaMemset(e->d_pen_c0, 0, (size_t)n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&e->d_pen_ov, (size_t)n * (nt + 1) * sizeof(double2)));
    CUDA_TRY(cudaMalloc(&e->d_pen_ovval, (size_t)n * sizeof(double)));
    if (p->purity) {
      CUDA_TRY(cudaMalloc(&e->d_pen_purity, (size_t)n * sizeof(double)));
      CUDA_TRY(cudaMemcpy(e->d_pen_purity, p->purity, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    }
    e->pen_overlap = true;
    e->pen.ov = e->d_pen_ov;
  } else {
    CUDA_TRY(cudaMemcpy(e->d_pen_c0, p->c0, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  }
  e->pen.a = e->d_pen_a; e->pen.d = e->d_pen_d; e->pen.g = e->d_pen_g; e->pen.wdt = e->d_pen_wdt; e->pen.jbar = e->d_pen_jbar;
  e->pen.gw = p->guard ? e->d_pen_gw : nullptr;
  e->pen.on = 1;
  e->pen_ninit = n;
  (void)dim;
  return QB_OK;
}

static void free_objective(qb_engine* e) {
  cudaFree(e->d_obj_target); cudaFree(e->d_obj_purity); cudaFree(e->d_obj_weight); cudaFree(e->d_obj_out); cudaFree(e->d_obj_sum);
  e->d_obj_target = nullptr;
  e->d_obj_purity = e->d_obj_weight = e->d_obj_out = e->d_obj_sum = nullptr;
  memset(&e->obj, 0, sizeof(e->obj));
  e->obj_ninit = 0;
}

int qb_engine_set_objective(qb_engine* e, const qb_objective_desc* o) {
  if (!e) return fail(QB_ERR_INVALID, "null engine");
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  free_objective(e);
  if (!o) return QB_OK;
  if (!o->weight || o->ninit < 1) return fail(QB_ERR_INVALID, "incomplete objective descriptor");
  if (o->objective != QB_OBJ_JFROBENIUS && o->objective != QB_OBJ_JTRACE && o->objective != QB_OBJ_JMEASURE)
    return fail(QB_ERR_INVALID, "unknown objective type %d", o->objective);
  if (o->pure_id < 0 && !o->target) return fail(QB_ERR_INVALID, "objective needs either a pure target (pure_id >= 0) or target states");
  if (o->pure_id >= e->S.N) return fail(QB_ERR_INVALID, "pure target state %d exceeds the system dimension %d", o->pure_id, e->S.N);
  if (o->objective == QB_OBJ_JMEASURE && o->pure_id < 0)
    return fail(QB_ERR_INVALID, "Jmeasure can only be used for 'pure' targets");     /* src/optimtarget.cpp:762-766 */
  const int n = o->ninit;
  int rc = ensure_capacity(e, n, false);
  if (rc) return rc;
  if (o->pure_id < 0) {
    CUDA_TRY(cudaMalloc(&e->d_obj_target, (size_t)n * e->dimp * sizeof(double2)));
    rc = launch_pack(e, o->target, e->d_obj_target, n);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  }
  if (o->purity) {
    CUDA_TRY(cudaMalloc(&e->d_obj_purity, (size_t)n * sizeof(double)));
    CUDA_TRY(cudaMemcpy(e->d_obj_purity, o->purity, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMalloc(&e->d_obj_weight, (size_t)n * sizeof(double)));
  CUDA_TRY(cudaMemcpy(e->d_obj_weight, o->weight, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&e->d_obj_out, (size_t)n * 4 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&e->d_obj_sum, 4 * sizeof(double)));
  e->obj.target = e->d_obj_target; e->obj.purity = e->d_obj_purity; e->obj.weight = e->d_obj_weight;
  e->obj.fid_scale = o->fid_scale; e->obj.kind = o->objective; e->obj.pure_id = o->pure_id < 0 ? -1 : o->pure_id;
  e->obj.lind = e->lind ? 1 : 0; e->obj.N = e->S.N; e->obj.dim = e->S.dim; e->obj.dimp = e->dimp;
  e->obj_ninit = n;
  return QB_OK;
}

int qb_engine_objective(qb_engine* e, int ninit, double out[4]) {
  if (!e || !out || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (e->obj_ninit != ninit) return fail(QB_ERR_STATE, "no objective installed for %d initial conditions (installed: %d)", ninit, e->obj_ninit);
  if (e->last_fwd_ninit != ninit) return fail(QB_ERR_STATE, "the objective needs the final states of a forward sweep of the same batch (last: %d)", e->last_fwd_ninit);
  objective_eval_kernel<<<ninit, 256, 0, e->stream>>>(e->d_rhoT, e->obj, e->d_obj_out);
  objective_sum_kernel<<<1, 256, 0, e->stream>>>(e->d_obj_out, ninit, e->d_obj_sum);
  e->stats.kernel_launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, e->d_obj_sum, 4 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

int qb_engine_objective_bar(qb_engine* e, int ninit, double re_bar, double im_bar) {
  if (!e || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (e->obj_ninit != ninit) return fail(QB_ERR_STATE, "no objective installed for %d initial conditions (installed: %d)", ninit, e->obj_ninit);
  if (e->last_fwd_ninit != ninit) return fail(QB_ERR_STATE, "the terminal adjoints need the final states of a forward sweep of the same batch (last: %d)", e->last_fwd_ninit);
  objective_bar_kernel<<<ninit, 256, 0, e->stream>>>(e->d_rhoT, e->obj, re_bar, im_bar, e->d_rbar);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return QB_OK;
}

int qb_engine_max_batch(qb_engine* e, int store_trajectory) {
  if (!e) { fail(QB_ERR_INVALID, "null engine"); return -1; }
  size_t free_b = 0, total_b = 0;
  if (cudaSetDevice(e->device) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { fail(QB_ERR_CUDA, "cudaMemGetInfo failed"); return -1; }
  /* buffers this engine already holds are reused */
  size_t have = (e->traj_cap + e->stage_cap) * sizeof(double2);
  /* states (rho0, rhoT, rbar, rho0bar, host-layout staging), per-initial-condition gradient, penalty / objective vectors */
  size_t per_ic = (size_t)e->dimp * sizeof(double2) * 8 + (size_t)e->S.nd * sizeof(double) + (size_t)(e->S.ntime + 1) * sizeof(double2);
  if (e->use_big) per_ic += (size_t)(e->nvec + 3) * e->dimp * sizeof(double2);
  if (e->use_q) per_ic += (size_t)(e->mcap + 1) * e->dimp * sizeof(double2) + (size_t)8 * e->mcap * e->mcap * sizeof(double);
  if (store_trajectory) {
    const bool stages_only = store_trajectory == 2 && e->store_stage && e->use_q;
    per_ic += (size_t)(e->S.ntime + 1) * e->dimp * sizeof(double2) * ((e->store_stage && !stages_only) ? 2 : 1);
  }
  const size_t margin = (size_t)1 << 30;
  const size_t avail = free_b + have > margin ? free_b + have - margin : 0;
  const size_t n = avail / per_ic;
  return n > 0x7fffffff ? 0x7fffffff : (int)n;
}

int qb_engine_set_dpdm(qb_engine* e, const double* jbar, int ninit) {
  if (!e) return fail(QB_ERR_INVALID, "null engine");
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  cudaFree(e->d_dp_jbar); cudaFree(e->d_dp_out);
  e->d_dp_jbar = e->d_dp_out = nullptr;
  e->dp_ninit = 0;
  if (!jbar) return QB_OK;
  if (ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (e->lind) return fail(QB_ERR_UNSUPPORTED, "the dpdm penalty exists for the Schroedinger solver only (as in the reference, src/optimproblem.cpp:116-121)");
  if (e->S.nstage != 1) return fail(QB_ERR_UNSUPPORTED, "the dpdm penalty is implemented for timestepper = IMR only");
  CUDA_TRY(cudaMalloc(&e->d_dp_jbar, (size_t)ninit * sizeof(double)));
  CUDA_TRY(cudaMalloc(&e->d_dp_out, (size_t)ninit * sizeof(double)));
  CUDA_TRY(cudaMemcpy(e->d_dp_jbar, jbar, (size_t)ninit * sizeof(double), cudaMemcpyHostToDevice));
  e->dp_ninit = ninit;
  return QB_OK;
}

int qb_engine_dpdm_integrals(qb_engine* e, int ninit, double* out) {
  if (!e || !out || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (e->dp_ninit != ninit) return fail(QB_ERR_STATE, "no dpdm penalty installed for %d initial conditions", ninit);
  if (!e->last_fwd_states || e->last_fwd_ninit != ninit) return fail(QB_ERR_STATE, "dpdm integrals need the stored trajectory of a forward sweep of the same batch");
  const double dtinv = 1.0 / (e->S.dt * e->S.dt * e->S.dt * e->S.dt);
  dpdm_reduce_kernel<<<ninit, 256, 0, e->stream>>>(e->d_traj, e->S.ntime, e->S.dim, e->dimp, dtinv, e->d_dp_out);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, e->d_dp_out, (size_t)ninit * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

int qb_engine_penalty_integrals(qb_engine* e, int ninit, double* out) {
  if (!e || !out || ninit < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (!e->pen.on || e->pen_ninit != ninit) return fail(QB_ERR_STATE, "no penalty installed for %d initial conditions", ninit);
  if (!e->last_fwd_states || e->last_fwd_ninit != ninit) return fail(QB_ERR_STATE, "penalty integrals need the stored trajectory of a forward sweep of the same batch");
  if (e->pen_overlap) {
    int ro = ensure_overlap_table(e, ninit);
    if (ro) return ro;
  }
  penalty_reduce_kernel<<<ninit, 256, 0, e->stream>>>(e->d_traj, e->S.ntime, e->S.dim, e->dimp, e->pen, e->d_pen_c0, e->d_pen_out,
                                                      e->pen_overlap ? e->d_pen_ovval : nullptr);
  e->stats.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, e->d_pen_out, (size_t)ninit * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

int qb_comm_unique_id(char id[128]) {
  if (!id) return fail(QB_ERR_INVALID, "null argument");
  int rc = nccl_load();
  if (rc) return rc;
  NcclId u;
  int r = g_nccl.GetUniqueId(&u);
  if (r != 0) return fail(QB_ERR_CUDA, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  memcpy(id, u.internal, 128);
  return QB_OK;
}

int qb_engine_comm_init(qb_engine* e, const char id[128], int rank, int nranks) {
  if (!e || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(QB_ERR_INVALID, "bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  /* ncclCommInitRank(ncclComm_t*, int nranks, ncclUniqueId id (by value), int rank) */
  typedef int (*init_fn)(void**, int, NcclId, int);
  init_fn init = (init_fn)dlsym(g_nccl.lib, "ncclCommInitRank");
  NcclId u;
  memcpy(u.internal, id, 128);
  int r = init(&e->nccl_comm, nranks, u, rank);
  if (r != 0) return fail(QB_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  e->comm_size = nranks;
  return QB_OK;
}

int qb_engine_allreduce(qb_engine* e, double* buf, int n) {
  if (!e || !buf || n < 1) return fail(QB_ERR_INVALID, "bad argument");
  if (n > e->S.nd + 8) return fail(QB_ERR_INVALID, "all-reduce of %d values exceeds the engine's buffer (%d)", n, e->S.nd + 8);
  if (e->comm_size <= 1) return QB_OK;
  if (!e->nccl_comm) return fail(QB_ERR_STATE, "qb_engine_comm_init has not been called");
  CUDA_TRY(cudaMemcpyAsync(e->d_G, buf, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
  int r = g_nccl.AllReduce(e->d_G, e->d_G, (size_t)n, 8 /* ncclDouble (ncclFloat64) */, 0 /* ncclSum */, e->nccl_comm, e->stream);
  if (r != 0) return fail(QB_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  CUDA_TRY(cudaMemcpyAsync(buf, e->d_G, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

int qb_engine_allreduce_grad_dev(qb_engine* e) {
  if (!e) return fail(QB_ERR_INVALID, "null engine");
  if (e->comm_size <= 1) return QB_OK;
  if (!e->nccl_comm) return fail(QB_ERR_STATE, "qb_engine_comm_init has not been called");
  int r = g_nccl.AllReduce(e->d_G, e->d_G, (size_t)(e->S.nd + 8), 8 /* ncclDouble */, 0 /* ncclSum */, e->nccl_comm, e->stream);
  if (r != 0) return fail(QB_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return QB_OK;
}

int qb_engine_set_stream(qb_engine* e, void* cuda_stream, int external) {
  if (!e) return fail(QB_ERR_INVALID, "null engine");
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->stream = external ? (cudaStream_t)cuda_stream : e->own_stream;
  return QB_OK;
}

/* FP64 FMA peak of the device, measured: 8 independent DFMA chains per thread, 2 flops per DFMA. */
int qb_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(QB_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    CUDA_TRY(cudaEventRecord(e0, 0));
    dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0 + 1e-9 * rep);
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  *tflops = best;
  return QB_OK;
}

int qb_engine_sync(qb_engine* e) {
  if (!e) return fail(QB_ERR_INVALID, "null engine");
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return QB_OK;
}

int qb_engine_get_stats(qb_engine* e, qb_stats* out) {
  if (!e || !out) return fail(QB_ERR_INVALID, "null argument");
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  long long iters = 0, evals = 0;
  int rc = read_counters(e, &iters, &evals);
  if (rc) return rc;
  e->stats.linsolve_iters = iters;
  e->stats.rhs_evals = evals;
  rc = resolve_timings(e);
  if (rc) return rc;
  *out = e->stats;
  return QB_OK;
}

}  // extern "C"
