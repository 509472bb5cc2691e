/*
 * qb_kernels.cuh -- device code of the small-N (one thread per element of vec(rho), dim <= 256) sweep kernels.
 * Included by the per-oscillator-count translation units qb_q1.cu .. qb_q5.cu (compiled in parallel) which
 * instantiate qb_pick<Q>().  See engine.cu for the overview and DESIGN.md for the layout and the roofline.
 *
 * Reference code these kernels replace (paths relative to the reference root):
 *   applyRHS_matfree / _transpose      src/mastereq.cpp:1279-2893, include/mastereq.hpp:316-912
 *   compute_dRHS_dParams_matfree       src/mastereq.cpp:970-1276, include/mastereq.hpp:553-604
 *   ImplMidpoint::evolveFWD/evolveBWD  src/timestepper.cpp:584-694,  NeumannSolve :697-727, KSPGMRES set-up :541-550
 *   TimeStepper::solveODE/solveAdjointODE  src/timestepper.cpp:96-253
 *   Oscillator::evalControl_diff       src/oscillator.cpp:339-381
 */
#pragma once
#include "qb_common.h"

namespace qbk {

/* ---------------------------------------------------------------------------------------------------------
 * Per-thread stencil of one element of vec(rho).
 *
 * The real-vectorised operator is M = H + D + L1, with
 *   H   the Hamiltonian part: for every neighbour nb a complex weight, y += (b - i a) * x_nb, plus the diagonal
 *       y += -i dA x.  H is skew-symmetric, hence M^T = -H + D + L1^T.
 *   D   the real diagonal dB (dephasing L2 + diagonal of the decay L1)
 *   L1  the decay shift rho_{i+1,i'+1} -> rho_{i,i'} (transpose: the opposite shift).
 * Control neighbours of oscillator k carry a = ca * p_k(t), b = cb * q_k(t); coupling neighbours of the pair
 * (k,l) carry a = ca * cos(eta_kl t), b = cb * sin(eta_kl t).
 * ALL2 (every oscillator has two levels): each element has exactly one bra neighbour (and one ket neighbour if
 * Lindblad) per oscillator and |ca| = |cb| = 1, so only the two sign bits are kept.
 * --------------------------------------------------------------------------------------------------------- */
template <int Q, bool ALL2, bool HASJ, bool LIND>
struct Stencil {
  static constexpr int NPAIR = Q * (Q - 1) / 2;
  static constexpr int SIDES = LIND ? 2 : 1;
  static constexpr int CPO = (ALL2 ? 1 : 2) * SIDES;       /* control slots per oscillator */
  static constexpr int JPP = (ALL2 ? 1 : 2) * SIDES;       /* coupling slots per pair */
  static constexpr int NC = CPO * Q;
  static constexpr int NJ = HASJ ? JPP * NPAIR : 0;
  static constexpr int NS = NC + NJ;
  static constexpr int NCOEF = (ALL2 ? 0 : NC) + NJ;       /* slots with explicit double coefficients */
  static constexpr int COEF0 = ALL2 ? NC : 0;              /* first slot that has them */
  int off[NS];                         /* neighbour element index */
  unsigned ma[ALL2 ? NC : 1];          /* ALL2: sign masks (0 or 0x80000000) of a and b for the control slots */
  unsigned mb[ALL2 ? NC : 1];
  double ca[NCOEF > 0 ? NCOEF : 1];
  double cb[NCOEF > 0 ? NCOEF : 1];
  int l1off_f[LIND ? Q : 1], l1off_t[LIND ? Q : 1];
  double l1f[LIND ? Q : 1], l1t[LIND ? Q : 1];
  double dA, dB;      /* diagonal: Hamiltonian (hd - hd'), dissipative (l2 + l1diag) */
  bool active;        /* false for the padding threads (element index >= dim): they contribute exact zeros */
};

__device__ __forceinline__ double sflip(double x, unsigned mask) {
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

__device__ __forceinline__ double hdiag_dev(const DevSys& S, const int* i) {
  double detune = 0.0, kerr = 0.0, cross = 0.0;
  for (int k = 0; k < S.Q; k++) detune += S.detuning[k] * i[k];
  for (int k = 0; k < S.Q; k++) kerr += -S.selfkerr[k] / 2.0 * i[k] * (i[k] - 1);
  int id = 0;
  for (int k = 0; k < S.Q; k++)
    for (int l = k + 1; l < S.Q; l++) { cross += -S.crosskerr[id] * i[k] * i[l]; id++; }
  return detune + kerr + cross;
}

/* Fills the stencil of element `it` (it >= dim: an idle thread, all coefficients zero). */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ void stencil_init(Stencil<Q, ALL2, HASJ, LIND>& st, const DevSys& S, int it) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  const bool active = it < S.dim;
  const int self = active ? it : 0;
#pragma unroll
  for (int s = 0; s < ST::NS; s++) st.off[s] = self;
#pragma unroll
  for (int s = 0; s < (ALL2 ? ST::NC : 1); s++) { st.ma[s] = 0u; st.mb[s] = 0u; }
#pragma unroll
  for (int s = 0; s < (ST::NCOEF > 0 ? ST::NCOEF : 1); s++) { st.ca[s] = 0.0; st.cb[s] = 0.0; }
#pragma unroll
  for (int k = 0; k < (LIND ? Q : 1); k++) { st.l1off_f[k] = self; st.l1off_t[k] = self; st.l1f[k] = 0.0; st.l1t[k] = 0.0; }
  st.dA = 0.0; st.dB = 0.0;
  st.active = active;
  if (!active) return;
  int i[kMaxOsc], ip[kMaxOsc], np[kMaxOsc], sp[kMaxOsc];
  int row = it % S.N, col = it / S.N;
  for (int k = Q - 1; k >= 0; k--) { i[k] = row % S.nlev[k]; row /= S.nlev[k]; }
  for (int k = Q - 1; k >= 0; k--) {
    if (LIND) { ip[k] = col % S.nlev[k]; col /= S.nlev[k]; np[k] = S.nlev[k]; }
    else { ip[k] = 0; np[k] = 1; }
    sp[k] = S.stride[k] * S.N;
  }
  /* diagonal, e.g. src/mastereq.cpp:2133-2154 */
  double hd = hdiag_dev(S, i);
  double hdp = LIND ? hdiag_dev(S, ip) : 0.0;
  st.dA = hd - hdp;
  if (LIND) {
    double l1diag = 0.0, l2 = 0.0;
    for (int k = 0; k < Q; k++) l1diag += -S.decay[k] / 2.0 * (i[k] + ip[k]);
    for (int k = 0; k < Q; k++) l2 += S.dephase[k] * (i[k] * ip[k] - 1. / 2. * (i[k] * i[k] + ip[k] * ip[k]));
    st.dB = l2 + l1diag;
  }
  /* control(), include/mastereq.hpp:772-855.
   *   bra +1: a = +sq p, b = +sq q     bra -1: a = +sq p, b = -sq q
   *   ket +1: a = -sq p, b = +sq q     ket -1: a = -sq p, b = -sq q */
  constexpr unsigned NEG = 0x80000000u;
#pragma unroll
  for (int k = 0; k < Q; k++) {
    const int n = S.nlev[k], a = i[k], ap = ip[k], si = S.stride[k], sip = sp[k];
    const int base = ST::CPO * k;
    if (ALL2) {
      if (a == 0) { st.off[base] = it + si; st.ma[base] = 0u; st.mb[base] = 0u; }
      else        { st.off[base] = it - si; st.ma[base] = 0u; st.mb[base] = NEG; }
      if (LIND) {
        if (ap == 0) { st.off[base + 1] = it + sip; st.ma[base + 1] = NEG; st.mb[base + 1] = 0u; }
        else         { st.off[base + 1] = it - sip; st.ma[base + 1] = NEG; st.mb[base + 1] = NEG; }
      }
    } else {
      /* slot order: bra+1, bra-1, ket+1, ket-1 */
      if (a < n - 1) { st.off[base] = it + si; st.ca[base] = sqrt((double)(a + 1)); st.cb[base] = sqrt((double)(a + 1)); }
      if (a > 0) { st.off[base + 1] = it - si; st.ca[base + 1] = sqrt((double)a); st.cb[base + 1] = -sqrt((double)a); }
      if (LIND) {
        if (ap < np[k] - 1) { st.off[base + 2] = it + sip; st.ca[base + 2] = -sqrt((double)(ap + 1)); st.cb[base + 2] = sqrt((double)(ap + 1)); }
        if (ap > 0) { st.off[base + 3] = it - sip; st.ca[base + 3] = -sqrt((double)ap); st.cb[base + 3] = -sqrt((double)ap); }
      }
    }
  }
  /* Jkl_coupling(), include/mastereq.hpp:632-675:
   *   1) bra k-1,l+1: a = +c cos, b = +c sin     2) bra k+1,l-1: a = +c cos, b = -c sin
   *   3) ket k-1,l+1: a = -c cos, b = +c sin     4) ket k+1,l-1: a = -c cos, b = -c sin,   c = J sqrt(..) */
  if (HASJ) {
    int id = 0;
#pragma unroll
    for (int k = 0; k < Q; k++) {
#pragma unroll
      for (int l = k + 1; l < Q; l++, id++) {
        const double J = S.Jkl[id];
        if (!(fabs(J) > 1e-10)) continue;
        const int ni = S.nlev[k], nj = S.nlev[l], nip = np[k], njp = np[l];
        const int a = i[k], ap = ip[k], b = i[l], bp = ip[l];
        const int si = S.stride[k], sj = S.stride[l], sip = sp[k], sjp = sp[l];
        const int base = ST::NC + ST::JPP * id;       /* slot */
        const int cbase = base - ST::COEF0;           /* coefficient index */
        if (ALL2) {
          if (a > 0 && b < nj - 1) { double c = J * sqrt((double)(a * (b + 1))); st.off[base] = it - si + sj; st.ca[cbase] = c; st.cb[cbase] = c; }
          else if (a < ni - 1 && b > 0) { double c = J * sqrt((double)(b * (a + 1))); st.off[base] = it + si - sj; st.ca[cbase] = c; st.cb[cbase] = -c; }
          if (LIND) {
            if (ap > 0 && bp < njp - 1) { double c = J * sqrt((double)(ap * (bp + 1))); st.off[base + 1] = it - sip + sjp; st.ca[cbase + 1] = -c; st.cb[cbase + 1] = c; }
            else if (ap < nip - 1 && bp > 0) { double c = J * sqrt((double)(bp * (ap + 1))); st.off[base + 1] = it + sip - sjp; st.ca[cbase + 1] = -c; st.cb[cbase + 1] = -c; }
          }
        } else {
          if (a > 0 && b < nj - 1) { double c = J * sqrt((double)(a * (b + 1))); st.off[base] = it - si + sj; st.ca[cbase] = c; st.cb[cbase] = c; }
          if (a < ni - 1 && b > 0) { double c = J * sqrt((double)(b * (a + 1))); st.off[base + 1] = it + si - sj; st.ca[cbase + 1] = c; st.cb[cbase + 1] = -c; }
          if (LIND) {
            if (ap > 0 && bp < njp - 1) { double c = J * sqrt((double)(ap * (bp + 1))); st.off[base + 2] = it - sip + sjp; st.ca[cbase + 2] = -c; st.cb[cbase + 2] = c; }
            if (ap < nip - 1 && bp > 0) { double c = J * sqrt((double)(bp * (ap + 1))); st.off[base + 3] = it + sip - sjp; st.ca[cbase + 3] = -c; st.cb[cbase + 3] = -c; }
          }
        }
      }
    }
  }
  /* L1decay / L1decay_T, include/mastereq.hpp:744-769 */
  if (LIND) {
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const double dec = S.decay[k];
      if (!(fabs(dec) > 1e-12)) continue;
      const int n = S.nlev[k], a = i[k], ap = ip[k];
      if (a < n - 1 && ap < n - 1) { st.l1f[k] = dec * sqrt((double)((a + 1) * (ap + 1))); st.l1off_f[k] = it + S.stride[k] + sp[k]; }
      if (a > 0 && ap > 0) { st.l1t[k] = dec * sqrt((double)(a * ap)); st.l1off_t[k] = it - S.stride[k] - sp[k]; }
    }
  }
}

/* Time-dependent weights for one step. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
struct StepCoef {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  double p[ALL2 ? Q : 1], q[ALL2 ? Q : 1];     /* ALL2: controls themselves */
  double a[ST::NCOEF > 0 ? ST::NCOEF : 1];     /* explicit-coefficient slots */
  double b[ST::NCOEF > 0 ? ST::NCOEF : 1];
};

template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ void stepcoef_load(StepCoef<Q, ALL2, HASJ, LIND>& c, const Stencil<Q, ALL2, HASJ, LIND>& st, const double* __restrict__ row) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
#pragma unroll
  for (int k = 0; k < Q; k++) {
    const double p = row[1 + k], q = row[1 + Q + k];
    if (ALL2) { c.p[k] = p; c.q[k] = q; }
    else {
#pragma unroll
      for (int s = 0; s < ST::CPO; s++) { c.a[ST::CPO * k + s] = st.ca[ST::CPO * k + s] * p; c.b[ST::CPO * k + s] = st.cb[ST::CPO * k + s] * q; }
    }
  }
  if (HASJ) {
#pragma unroll
    for (int id = 0; id < ST::NPAIR; id++) {
      const double cs = row[1 + 2 * Q + id], sn = row[1 + 2 * Q + ST::NPAIR + id];
#pragma unroll
      for (int s = 0; s < ST::JPP; s++) {
        const int ci = ST::NC + ST::JPP * id + s - ST::COEF0;
        c.a[ci] = st.ca[ci] * cs;
        c.b[ci] = st.cb[ci] * sn;
      }
    }
  }
}

/* y_own = (M x)_own or (M^T x)_own; xs = the vector in shared memory (interleaved re/im), x = own element. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ double2 apply_rhs(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c,
                                             const double2* __restrict__ xs, double2 x, bool transpose) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  double hre = st.dA * x.y, him = -st.dA * x.x;
#pragma unroll
  for (int s = 0; s < ST::NS; s++) {
    const double2 nb = xs[st.off[s]];
    double a, b;
    if (ALL2 && s < ST::NC) { a = sflip(c.p[s / ST::CPO], st.ma[s]); b = sflip(c.q[s / ST::CPO], st.mb[s]); }
    else { a = c.a[s - ST::COEF0]; b = c.b[s - ST::COEF0]; }
    hre = fma(a, nb.y, hre);
    hre = fma(b, nb.x, hre);
    him = fma(-a, nb.x, him);
    him = fma(b, nb.y, him);
  }
  double dre = 0.0, dim = 0.0;
  if (LIND) {
    dre = st.dB * x.x; dim = st.dB * x.y;
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const double2 nb = xs[transpose ? st.l1off_t[k] : st.l1off_f[k]];
      const double l1 = transpose ? st.l1t[k] : st.l1f[k];
      dre = fma(l1, nb.x, dre);
      dim = fma(l1, nb.y, dim);
    }
  }
  if (!st.active) return make_double2(0.0, 0.0);
  return transpose ? make_double2(dre - hre, dim - him) : make_double2(dre + hre, dim + him);
}

/* ---------------------------------------------------------------------------------------------------------
 * Block-level machinery
 * --------------------------------------------------------------------------------------------------------- */
struct Smem {
  double2* X;         /* current vector whose neighbours are read by apply_rhs */
  double2* Y;         /* second vector (adjoint stage) */
  double2* V;         /* Krylov basis, (mcap+1) vectors of dimp (Neumann: ping-pong pair) */
  double* red;        /* [2][nwarps][kRedStride] reduction partials */
  double* hh;         /* rotated Hessenberg, column-major with ld = kLdh */
  double* cc;         /* Givens cos [kRestart+1] */
  double* ss;
  double* grs;        /* [kRestart+2] */
  double* nrs;        /* [kRestart+1] */
  double* bc;         /* broadcast scalars [8] */
  double* grad;       /* [nd] gradient accumulator of this initial condition */
  int dimp;
};

__device__ __forceinline__ Smem carve(unsigned char* base, int dimp, int mcap, int nd) {
  Smem sm;
  sm.dimp = dimp;
  const int nwarps = dimp >> 5;
  double2* v = reinterpret_cast<double2*>(base);
  sm.X = v; v += dimp;
  sm.Y = v; v += dimp;
  sm.V = v; v += (size_t)(mcap + 1) * dimp;
  double* d = reinterpret_cast<double*>(v);
  sm.red = d; d += 2 * nwarps * kRedStride;
  sm.hh = d; d += kLdh * (mcap + 1);
  sm.cc = d; d += kRestart + 1;
  sm.ss = d; d += kRestart + 1;
  sm.grs = d; d += kRestart + 2;
  sm.nrs = d; d += kRestart + 1;
  sm.bc = d; d += 8;
  sm.grad = d; d += nd;
  return sm;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Blk {
  int tid, lane, warp, nwarps;
  int parity;   /* which half of Smem::red the next reduction uses */
};

/* Sum of one value over the CTA; every thread gets the same bits.  One barrier. */
__device__ __forceinline__ double block_sum(double v, const Smem& sm, Blk& b) {
  v = warp_sum(v);
  double* r = sm.red + (size_t)b.parity * b.nwarps * kRedStride;
  if (b.lane == 0) r[b.warp * kRedStride] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < b.nwarps; w++) s += r[w * kRedStride];
  b.parity ^= 1;
  return s;
}

struct SolveStats { long long iters, evals; };

/* GMRES on (I - alpha*M) y = rhs  (or the transposed operator), zero initial guess.
 * Same algorithm and stopping rule as the reference's KSP set-up (src/timestepper.cpp:541-550: KSPGMRES, PCNONE,
 * rtol 1e-20, abstol 1e-10, maxit = linearsolver_maxiter; PETSc defaults restart 30, classical Gram-Schmidt, no
 * refinement, KSPConvergedDefault on the recurrence residual).  rhs: own element.  Returns own element of y. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ double2 gmres_solve(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const DevSys& S,
                                            const Smem& sm, Blk& b, double2 rhs, double alpha, bool transpose, int mcap, SolveStats& stats) {
  const int tid = b.tid, dimp = sm.dimp;
  const int ldh = kLdh;
  double2 y = make_double2(0.0, 0.0);
  int its = 0;
  bool converged = false, first = true;
  double ttol = 0.0;
  while (!converged && its < S.maxiter) {
    double2 r;
    if (first) r = rhs;
    else {
      /* r = b - A y */
      sm.X[tid] = y;
      __syncthreads();
      double2 my = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, sm.X, y, transpose);
      stats.evals++;
      r.x = rhs.x - (y.x - alpha * my.x);
      r.y = rhs.y - (y.y - alpha * my.y);
    }
    double res = sqrt(block_sum(r.x * r.x + r.y * r.y, sm, b));
    if (first) { ttol = fmax(S.reltol * res, S.abstol); first = false; }
    if (res == 0.0) { converged = true; break; }
    double2 v = make_double2(r.x / res, r.y / res);
    sm.V[tid] = v;
    if (tid == 0) sm.grs[0] = res;
    if (res <= ttol) { converged = true; break; }
    __syncthreads();
    int it = 0;
    while (!converged && it < mcap && its < S.maxiter) {
      /* w = A v_it */
      const double2* vin = sm.V + (size_t)it * dimp;
      const double2 vown = vin[tid];
      const double2 mv = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, vin, vown, transpose);
      stats.evals++;
      double2 w = make_double2(vown.x - alpha * mv.x, vown.y - alpha * mv.y);
      /* classical Gram-Schmidt: all it+1 inner products from the same w */
      double* red = sm.red + (size_t)b.parity * b.nwarps * kRedStride;
      for (int j = 0; j <= it; j++) {
        const double2 vj = sm.V[(size_t)j * dimp + tid];
        double d = warp_sum(w.x * vj.x + w.y * vj.y);
        if (b.lane == 0) red[b.warp * kRedStride + j] = d;
      }
      __syncthreads();
      for (int j = 0; j <= it; j++) {
        double h = 0.0;
        for (int wq = 0; wq < b.nwarps; wq++) h += red[wq * kRedStride + j];
        const double2 vj = sm.V[(size_t)j * dimp + tid];
        w.x -= h * vj.x;
        w.y -= h * vj.y;
        if (tid == 0) sm.hh[it * ldh + j] = h;
      }
      b.parity ^= 1;
      const double tt = sqrt(block_sum(w.x * w.x + w.y * w.y, sm, b));
      if (tt > 0.0) { w.x /= tt; w.y /= tt; }
      sm.V[(size_t)(it + 1) * dimp + tid] = w;
      if (tid == 0) {
        /* KSPGMRESUpdateHessenberg: previous rotations, new rotation, residual estimate */
        double* hcol = sm.hh + it * ldh;
        hcol[it + 1] = tt;
        for (int j = 0; j < it; j++) {
          double t = hcol[j];
          hcol[j] = sm.cc[j] * t + sm.ss[j] * hcol[j + 1];
          hcol[j + 1] = -sm.ss[j] * t + sm.cc[j] * hcol[j + 1];
        }
        double t = sqrt(hcol[it] * hcol[it] + hcol[it + 1] * hcol[it + 1]);
        double flag;
        if (t == 0.0) flag = 2.0;  /* breakdown: converged */
        else {
          sm.cc[it] = hcol[it] / t;
          sm.ss[it] = hcol[it + 1] / t;
          sm.grs[it + 1] = -sm.ss[it] * sm.grs[it];
          sm.grs[it] = sm.cc[it] * sm.grs[it];
          hcol[it] = sm.cc[it] * hcol[it] + sm.ss[it] * hcol[it + 1];
          flag = (fabs(sm.grs[it + 1]) <= ttol) ? 1.0 : 0.0;
        }
        sm.bc[0] = flag;
      }
      __syncthreads();
      const double flag = sm.bc[0];
      it++;
      its++;
      if (flag != 0.0) converged = true;
    }
    /* KSPGMRESBuildSoln: back substitution by thread 0, combination by everyone */
    if (it > 0) {
      if (tid == 0) {
        for (int kk = it - 1; kk >= 0; kk--) {
          double tsum = sm.grs[kk];
          for (int j = kk + 1; j < it; j++) tsum -= sm.hh[j * ldh + kk] * sm.nrs[j];
          sm.nrs[kk] = tsum / sm.hh[kk * ldh + kk];
        }
      }
      __syncthreads();
      for (int j = 0; j < it; j++) {
        const double2 vj = sm.V[(size_t)j * dimp + tid];
        const double cj = sm.nrs[j];
        y.x += cj * vj.x;
        y.y += cj * vj.y;
      }
    }
    __syncthreads(); /* basis and scalars are reused by the next cycle / next solve */
  }
  stats.iters += its;
  return y;
}

/* ImplMidpoint::NeumannSolve, src/timestepper.cpp:697-727: y = rhs + alpha*M*y, stop on ||dy|| < abstol or
 * ||dy||/||dy_0|| < reltol or maxiter. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ double2 neumann_solve(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const DevSys& S,
                                              const Smem& sm, Blk& b, double2 rhs, double alpha, bool transpose, SolveStats& stats) {
  const int tid = b.tid;
  double2 y = rhs;
  double2* cur = sm.V;
  double2* nxt = sm.V + sm.dimp;
  cur[tid] = y;
  __syncthreads();
  double errnorm0 = 0.0;
  int iter;
  for (iter = 0; iter < S.maxiter; iter++) {
    const double2 my = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, cur, y, transpose);
    stats.evals++;
    const double2 yn = make_double2(1.0 * rhs.x + alpha * my.x, 1.0 * rhs.y + alpha * my.y);
    const double ex = y.x - yn.x, ey = y.y - yn.y;
    y = yn;
    nxt[tid] = y;
    const double errnorm = sqrt(block_sum(ex * ex + ey * ey, sm, b));
    double2* t = cur; cur = nxt; nxt = t;
    if (iter == 0) errnorm0 = errnorm;
    if (errnorm < S.abstol) break;
    if (errnorm / errnorm0 < S.reltol) break;
  }
  stats.iters += iter;
  __syncthreads();
  return y;
}

template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ double2 linsolve(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const DevSys& S,
                                            const Smem& sm, Blk& b, double2 rhs, double alpha, bool transpose, int mcap, SolveStats& stats) {
  if (S.linsolve == QB_LINSOLVE_GMRES) return gmres_solve<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, rhs, alpha, transpose, mcap, stats);
  return neumann_solve<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, rhs, alpha, transpose, stats);
}

/* One implicit-midpoint step  x <- x + dtn * k,  (I - dtn/2 M) k = M x   (ImplMidpoint::evolveFWD,
 * src/timestepper.cpp:584-629).  x: own element; sm.X holds x on entry (visible to all threads). */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ double2 imr_step(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const DevSys& S,
                                            const Smem& sm, Blk& b, double2 x, double dtn, int mcap, SolveStats& stats) {
  const double2 rhs = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, sm.X, x, false);
  stats.evals++;
  const double2 k = linsolve<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, rhs, dtn / 2.0, false, mcap, stats);
  return make_double2(x.x + dtn * k.x, x.y + dtn * k.y);
}

/* Forward sweep of one initial condition per CTA: TimeStepper::solveODE, src/timestepper.cpp:96-181. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__global__ void __launch_bounds__(kMaxThreads, 2) fwd_sweep_kernel(const __grid_constant__ DevSys S, const __grid_constant__ SweepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dimp = blockDim.x;
  Smem sm = carve(smem_raw, dimp, A.mcap, 0);
  Blk b;
  b.tid = threadIdx.x; b.lane = threadIdx.x & 31; b.warp = threadIdx.x >> 5; b.nwarps = blockDim.x >> 5; b.parity = 0;
  const int ic = blockIdx.x;
  Stencil<Q, ALL2, HASJ, LIND> st;
  stencil_init(st, S, b.tid);
  StepCoef<Q, ALL2, HASJ, LIND> c;
  SolveStats stats = {0, 0};

  double2 x = A.state[(size_t)ic * dimp + b.tid];
  double2* traj = A.store ? A.traj + (size_t)ic * (S.ntime + 1) * dimp : nullptr;
  for (int n = 0; n < S.ntime; n++) {
    const double* row = A.tab + (size_t)n * S.tabstride;
    stepcoef_load(c, st, row);
    const double dtn = row[0];
    if (traj) traj[(size_t)n * dimp + b.tid] = x;
    sm.X[b.tid] = x;
    __syncthreads();
    x = imr_step<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, x, dtn, A.mcap, stats);
  }
  if (traj) traj[(size_t)S.ntime * dimp + b.tid] = x;
  A.state_out[(size_t)ic * dimp + b.tid] = x;
  if (b.tid == 0) {
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* Adjoint sweep of one initial condition per CTA: TimeStepper::solveAdjointODE + ImplMidpoint::evolveBWD,
 * src/timestepper.cpp:184-253,631-694, with compute_dRHS_dParams_matfree (src/mastereq.cpp:970-1276) and
 * Oscillator::evalControl_diff (src/oscillator.cpp:339-381) fused into the step. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__global__ void __launch_bounds__(kMaxThreads, 2) adj_sweep_kernel(const __grid_constant__ DevSys S, const __grid_constant__ SweepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dimp = blockDim.x;
  Smem sm = carve(smem_raw, dimp, A.mcap, S.nd);
  Blk b;
  b.tid = threadIdx.x; b.lane = threadIdx.x & 31; b.warp = threadIdx.x >> 5; b.nwarps = blockDim.x >> 5; b.parity = 0;
  const int ic = blockIdx.x;
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  ST st;
  stencil_init(st, S, b.tid);
  StepCoef<Q, ALL2, HASJ, LIND> c;
  SolveStats stats = {0, 0};
  const bool recompute = (S.flags & QB_FLAG_RECOMPUTE_STAGE) != 0;

  for (int i = b.tid; i < S.nd; i += blockDim.x) sm.grad[i] = 0.0;
  /* gradient updater role of this thread */
  int uk = -1, uc = 0, uf = 0, uj = 0;
  if (b.tid < S.nupd) { uk = A.upd[4 * b.tid]; uc = A.upd[4 * b.tid + 1]; uf = A.upd[4 * b.tid + 2]; uj = A.upd[4 * b.tid + 3]; }

  double2 xadj = A.state[(size_t)ic * dimp + b.tid];
  const double2* traj = A.traj ? A.traj + (size_t)ic * (S.ntime + 1) * dimp : nullptr;
  double2 xcur, xprev;
  if (traj) {
    xcur = traj[(size_t)S.ntime * dimp + b.tid];
    xprev = traj[(size_t)(S.ntime - 1) * dimp + b.tid];
  } else {
    xcur = A.final_state[(size_t)ic * dimp + b.tid];
    xprev = xcur;
  }
  __syncthreads();

  for (int n = S.ntime; n > 0; n--) {
    const double* row = A.tab + (size_t)(n - 1) * S.tabstride;
    stepcoef_load(c, st, row);
    const double dtn = row[0];
    /* primal state at n-1: from the stored trajectory (Lindblad) or one IMR step backwards (Schroedinger),
     * src/timestepper.cpp:230-231 */
    double2 xnext_prefetch = make_double2(0.0, 0.0);
    if (traj) {
      if (n >= 2) xnext_prefetch = traj[(size_t)(n - 2) * dimp + b.tid];  /* in flight during this step */
    } else {
      sm.X[b.tid] = xcur;
      __syncthreads();
      xprev = imr_step<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, xcur, -dtn, A.mcap, stats);
    }
    /* adjoint stage: (I - dtn/2 M)^T kbar = xadj ; kbar *= dtn */
    double2 kbar = linsolve<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, xadj, dtn / 2.0, true, A.mcap, stats);
    kbar.x *= dtn; kbar.y *= dtn;
    /* primal stage x_{n-1} + dtn/2 k */
    double2 stage;
    if (recompute) {
      sm.X[b.tid] = xprev;
      __syncthreads();
      const double2 rhs = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, sm.X, xprev, false);
      stats.evals++;
      const double2 k = linsolve<Q, ALL2, HASJ, LIND>(st, c, S, sm, b, rhs, dtn / 2.0, false, A.mcap, stats);
      stage = make_double2(xprev.x + dtn / 2.0 * k.x, xprev.y + dtn / 2.0 * k.y);
    } else {
      stage = make_double2(0.5 * (xprev.x + xcur.x), 0.5 * (xprev.y + xcur.y));
    }
    sm.X[b.tid] = stage;
    sm.Y[b.tid] = kbar;
    __syncthreads();
    /* dRHSdp coefficients: coeff_p[k] = kbar^T (dM/dp_k) stage, coeff_q[k] likewise */
    {
      double* red = sm.red + (size_t)b.parity * b.nwarps * kRedStride;
#pragma unroll
      for (int k = 0; k < Q; k++) {
        double cp = 0.0, cq = 0.0;
#pragma unroll
        for (int s = 0; s < ST::CPO; s++) {
          const int sl = ST::CPO * k + s;
          const double2 nb = sm.X[st.off[sl]];
          const double tp = nb.y * kbar.x - nb.x * kbar.y;
          const double tq = nb.x * kbar.x + nb.y * kbar.y;
          if (ALL2) { if (st.active) { cp += sflip(tp, st.ma[sl]); cq += sflip(tq, st.mb[sl]); } }
          else { cp += st.ca[sl] * tp; cq += st.cb[sl] * tq; }
        }
        cp = warp_sum(cp);
        cq = warp_sum(cq);
        if (b.lane == 0) { red[b.warp * kRedStride + 2 * k] = cp; red[b.warp * kRedStride + 2 * k + 1] = cq; }
      }
      stats.evals++;
      /* xadj += M^T kbar */
      const double2 mt = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, sm.Y, kbar, true);
      stats.evals++;
      xadj.x += mt.x; xadj.y += mt.y;
      __syncthreads();
      if (uk >= 0) {
        double cp = 0.0, cq = 0.0;
        for (int w = 0; w < b.nwarps; w++) { cp += red[w * kRedStride + 2 * uk]; cq += red[w * kRedStride + 2 * uk + 1]; }
        const int* gi = A.gidx + ((size_t)(n - 1) * Q + uk) * 4;
        const int idx = gi[uj];
        if (idx >= 0) {
          const int ns = gi[3];
          const double bval = A.gval[((size_t)(n - 1) * Q + uk) * 3 + uj];
          const double cs = A.gcs[((size_t)(n - 1) * S.ncar_total + uc) * 2], sn = A.gcs[((size_t)(n - 1) * S.ncar_total + uc) * 2 + 1];
          const double pbar = 1.0 * cp, qbar = 1.0 * cq;
          const double B1bar = sn * qbar + cs * pbar;
          const double B2bar = cs * qbar - sn * pbar;
          const int gid = S.osc_off[uk] + idx + uf * ns * 2;
          sm.grad[gid] += bval * B1bar;
          sm.grad[gid + ns] += bval * B2bar;
        }
      }
      b.parity ^= 1;
    }
    xcur = xprev;
    if (traj) xprev = xnext_prefetch;
  }
  __syncthreads();
  for (int i = b.tid; i < S.nd; i += blockDim.x) A.grad_ic[(size_t)ic * S.nd + i] = sm.grad[i];
  A.state_out[(size_t)ic * dimp + b.tid] = xadj;
  if (b.tid == 0) {
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* y = M(t) x or M(t)^T x for a batch of vectors, one CTA per vector (unit-test hook). */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__global__ void __launch_bounds__(kMaxThreads) rhs_apply_kernel(const __grid_constant__ DevSys S, const double* __restrict__ row,
                                                               const double2* __restrict__ x, double2* __restrict__ y, int transpose) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* xs = reinterpret_cast<double2*>(smem_raw);
  const int dimp = blockDim.x, tid = threadIdx.x;
  Stencil<Q, ALL2, HASJ, LIND> st;
  stencil_init(st, S, tid);
  StepCoef<Q, ALL2, HASJ, LIND> c;
  stepcoef_load(c, st, row);
  const double2 xo = x[(size_t)blockIdx.x * dimp + tid];
  xs[tid] = xo;
  __syncthreads();
  y[(size_t)blockIdx.x * dimp + tid] = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, xs, xo, transpose != 0);
}

template <int Q, bool ALL2, bool HASJ, bool LIND>
void pick3(SweepFn* fwd, SweepFn* adj, ApplyFn* app) {
  *app = rhs_apply_kernel<Q, ALL2, HASJ, LIND>;
  *fwd = fwd_sweep_kernel<Q, ALL2, HASJ, LIND>;
  *adj = adj_sweep_kernel<Q, ALL2, HASJ, LIND>;
}
template <int Q, bool ALL2>
void pick2(bool hasj, bool lind, SweepFn* fwd, SweepFn* adj, ApplyFn* app) {
  if constexpr (Q >= 2) {
    if (hasj) { if (lind) pick3<Q, ALL2, true, true>(fwd, adj, app); else pick3<Q, ALL2, true, false>(fwd, adj, app); return; }
  }
  if (lind) pick3<Q, ALL2, false, true>(fwd, adj, app); else pick3<Q, ALL2, false, false>(fwd, adj, app);
}

}  // namespace qbk

/* Selects the kernels for a system with Q oscillators. */
template <int Q>
void qb_pick(bool all2, bool hasj, bool lind, SweepFn* fwd, SweepFn* adj, ApplyFn* app) {
  if (all2) qbk::pick2<Q, true>(hasj, lind, fwd, adj, app);
  else qbk::pick2<Q, false>(hasj, lind, fwd, adj, app);
}
