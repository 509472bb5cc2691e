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
 *
 * Everything in a sweep is latency bound (a 4 KB state, ~10 dependent block-wide phases per linear-solver
 * iteration), so the code is organised around short dependency chains:
 *   - DIMP (threads per CTA = padded dimension) is a template parameter: every shared-memory address is
 *     `32-bit base + compile-time offset (+ per-thread byte offset kept in a register)`.
 *   - reductions are batched (all Gram-Schmidt inner products of an iteration share one butterfly and one
 *     barrier) and unrolled four-wide for instruction-level parallelism.
 *   - the Hessenberg/Givens recurrence of GMRES is carried redundantly by lane 0 of EVERY warp on a warp-private
 *     copy (identical arithmetic => identical decisions), so no barrier is spent on broadcasting the stopping flag.
 */
#pragma once
#include "qb_common.h"

namespace qbk {

/* ---------------------------------------------------------------------------------------------------------
 * shared-memory access through 32-bit shared-window addresses
 * --------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double2 lds2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts2(unsigned addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double lds1(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts1(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

/* 1/x and 1/sqrt(x) without the IEEE division / square-root slow paths: hardware seed + Newton steps in FMA
 * (relative error <= ~2 ulp; every thread of a CTA runs the same instructions on the same bits). */
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  /* Newton: r <- r + r*(1 - x r^2)/2 */
  double h = 0.5 * r;
  double e = fma(-x * r, r, 1.0);
  r = fma(h, e, r);
  h = 0.5 * r; e = fma(-x * r, r, 1.0);
  r = fma(h, e, r);
  h = 0.5 * r; e = fma(-x * r, r, 1.0);
  return fma(h, e, r);
}

}  // namespace qbk
#include "qb_mono.cuh"      /* scalar part of the one-reduction solver (uses qbk::fast_rcp above) */
namespace qbk {

/* Static shared-memory layout (byte offsets from the dynamic shared base). */
template <int DIMP>
struct Lay {
  static constexpr int NW = DIMP / 32;
  static constexpr unsigned RED = 0;                                   /* [2][kRedStride][NW] doubles */
  static constexpr unsigned RED_HALF = kRedStride * NW * 8;
  static constexpr unsigned GIV = RED + 2 * RED_HALF;                  /* per warp: cc[32] ss[32] grs[32] hcol[32] nrs[32] rdiag[32] */
  static constexpr unsigned GIV_WARP = 6 * 32 * 8;
  static constexpr unsigned X = GIV + NW * GIV_WARP;                   /* double2[DIMP] */
  static constexpr unsigned Y = X + DIMP * 16;
  static constexpr unsigned V = Y + DIMP * 16;                         /* double2[nvec][DIMP], nvec = max(mcap, 2) */
  static constexpr unsigned VEC = DIMP * 16;
  /* after V: hh (kLdh x (mcap+1) doubles, written by thread 0), then grad[nd] */
  /* Arnoldi basis (mcap vectors) or the monomial basis u_0..u_kSMax of the one-reduction solver, whichever is longer */
  static __host__ __device__ constexpr int nvec(int mcap) { return mcap > 7 ? mcap : 7; }
};

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
 * Neighbour offsets are stored as BYTE offsets inside a vector (element index * 16).
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
  static constexpr int ROW = 1 + 2 * Q + (HASJ ? 2 * NPAIR : 0);  /* doubles of a control-table row that are used */
  unsigned off[NS];                    /* neighbour byte offset */
  unsigned ma[ALL2 ? NC : 1];          /* ALL2: sign masks (0 or 0x80000000) of a and b for the control slots */
  unsigned mb[ALL2 ? NC : 1];
  double ca[NCOEF > 0 ? NCOEF : 1];
  double cb[NCOEF > 0 ? NCOEF : 1];
  unsigned l1off_f[LIND ? Q : 1], l1off_t[LIND ? Q : 1];
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
  int off[ST::NS], l1f_off[Q], l1t_off[Q];
#pragma unroll
  for (int s = 0; s < ST::NS; s++) off[s] = self;
#pragma unroll
  for (int s = 0; s < (ALL2 ? ST::NC : 1); s++) { st.ma[s] = 0u; st.mb[s] = 0u; }
#pragma unroll
  for (int s = 0; s < (ST::NCOEF > 0 ? ST::NCOEF : 1); s++) { st.ca[s] = 0.0; st.cb[s] = 0.0; }
#pragma unroll
  for (int k = 0; k < Q; k++) { l1f_off[k] = self; l1t_off[k] = self; }
#pragma unroll
  for (int k = 0; k < (LIND ? Q : 1); k++) { st.l1f[k] = 0.0; st.l1t[k] = 0.0; }
  st.dA = 0.0; st.dB = 0.0;
  st.active = active;
  if (active) {
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
        if (a == 0) { off[base] = it + si; st.ma[base] = 0u; st.mb[base] = 0u; }
        else        { off[base] = it - si; st.ma[base] = 0u; st.mb[base] = NEG; }
        if (LIND) {
          if (ap == 0) { off[base + 1] = it + sip; st.ma[base + 1] = NEG; st.mb[base + 1] = 0u; }
          else         { off[base + 1] = it - sip; st.ma[base + 1] = NEG; st.mb[base + 1] = NEG; }
        }
      } else {
        /* slot order: bra+1, bra-1, ket+1, ket-1 */
        if (a < n - 1) { off[base] = it + si; st.ca[base] = sqrt((double)(a + 1)); st.cb[base] = sqrt((double)(a + 1)); }
        if (a > 0) { off[base + 1] = it - si; st.ca[base + 1] = sqrt((double)a); st.cb[base + 1] = -sqrt((double)a); }
        if (LIND) {
          if (ap < np[k] - 1) { off[base + 2] = it + sip; st.ca[base + 2] = -sqrt((double)(ap + 1)); st.cb[base + 2] = sqrt((double)(ap + 1)); }
          if (ap > 0) { off[base + 3] = it - sip; st.ca[base + 3] = -sqrt((double)ap); st.cb[base + 3] = -sqrt((double)ap); }
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
            if (a > 0 && b < nj - 1) { double c = J * sqrt((double)(a * (b + 1))); off[base] = it - si + sj; st.ca[cbase] = c; st.cb[cbase] = c; }
            else if (a < ni - 1 && b > 0) { double c = J * sqrt((double)(b * (a + 1))); off[base] = it + si - sj; st.ca[cbase] = c; st.cb[cbase] = -c; }
            if (LIND) {
              if (ap > 0 && bp < njp - 1) { double c = J * sqrt((double)(ap * (bp + 1))); off[base + 1] = it - sip + sjp; st.ca[cbase + 1] = -c; st.cb[cbase + 1] = c; }
              else if (ap < nip - 1 && bp > 0) { double c = J * sqrt((double)(bp * (ap + 1))); off[base + 1] = it + sip - sjp; st.ca[cbase + 1] = -c; st.cb[cbase + 1] = -c; }
            }
          } else {
            if (a > 0 && b < nj - 1) { double c = J * sqrt((double)(a * (b + 1))); off[base] = it - si + sj; st.ca[cbase] = c; st.cb[cbase] = c; }
            if (a < ni - 1 && b > 0) { double c = J * sqrt((double)(b * (a + 1))); off[base + 1] = it + si - sj; st.ca[cbase + 1] = c; st.cb[cbase + 1] = -c; }
            if (LIND) {
              if (ap > 0 && bp < njp - 1) { double c = J * sqrt((double)(ap * (bp + 1))); off[base + 2] = it - sip + sjp; st.ca[cbase + 2] = -c; st.cb[cbase + 2] = c; }
              if (ap < nip - 1 && bp > 0) { double c = J * sqrt((double)(bp * (ap + 1))); off[base + 3] = it + sip - sjp; st.ca[cbase + 3] = -c; st.cb[cbase + 3] = -c; }
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
        if (a < n - 1 && ap < n - 1) { st.l1f[k] = dec * sqrt((double)((a + 1) * (ap + 1))); l1f_off[k] = it + S.stride[k] + sp[k]; }
        if (a > 0 && ap > 0) { st.l1t[k] = dec * sqrt((double)(a * ap)); l1t_off[k] = it - S.stride[k] - sp[k]; }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < ST::NS; s++) st.off[s] = (unsigned)off[s] * 16u;
#pragma unroll
  for (int k = 0; k < (LIND ? Q : 1); k++) { st.l1off_f[k] = (unsigned)l1f_off[k] * 16u; st.l1off_t[k] = (unsigned)l1t_off[k] * 16u; }
}

/* Time-dependent weights for one step. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
struct StepCoef {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  double p[ALL2 ? Q : 1], q[ALL2 ? Q : 1];     /* ALL2: controls themselves */
  double a[ST::NCOEF > 0 ? ST::NCOEF : 1];     /* explicit-coefficient slots */
  double b[ST::NCOEF > 0 ? ST::NCOEF : 1];
};

/* row: the first ST::ROW doubles of a control-table row, already in registers */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ void stepcoef_set(StepCoef<Q, ALL2, HASJ, LIND>& c, const Stencil<Q, ALL2, HASJ, LIND>& st, const double* row) {
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

template <int ROW>
__device__ __forceinline__ void row_load(double (&r)[ROW], const double* __restrict__ g) {
#pragma unroll
  for (int i = 0; i < ROW; i++) r[i] = __ldg(g + i);
}

/* y_own = (M x)_own or (M^T x)_own; xs = 32-bit shared address of the vector (interleaved re/im), x = own element. */
template <int Q, bool ALL2, bool HASJ, bool LIND>
__device__ __forceinline__ double2 apply_rhs(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, unsigned xs, double2 x,
                                             bool transpose) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  double2 nb[ST::NS];
#pragma unroll
  for (int s = 0; s < ST::NS; s++) nb[s] = lds2(xs + st.off[s]);
  double2 nl[LIND ? Q : 1];
  if (LIND) {
#pragma unroll
    for (int k = 0; k < Q; k++) nl[k] = lds2(xs + (transpose ? st.l1off_t[k] : st.l1off_f[k]));
  }
  /* two independent accumulation chains per component */
  double hre = st.dA * x.y, him = -st.dA * x.x, hre2 = 0.0, him2 = 0.0;
#pragma unroll
  for (int s = 0; s < ST::NS; s++) {
    double a, b;
    if (ALL2 && s < ST::NC) { a = sflip(c.p[s / ST::CPO], st.ma[s]); b = sflip(c.q[s / ST::CPO], st.mb[s]); }
    else { a = c.a[s - ST::COEF0]; b = c.b[s - ST::COEF0]; }
    hre = fma(a, nb[s].y, hre);
    hre2 = fma(b, nb[s].x, hre2);
    him = fma(-a, nb[s].x, him);
    him2 = fma(b, nb[s].y, him2);
  }
  hre += hre2; him += him2;
  double dre = 0.0, dim = 0.0;
  if (LIND) {
    dre = st.dB * x.x; dim = st.dB * x.y;
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const double l1 = transpose ? st.l1t[k] : st.l1f[k];
      dre = fma(l1, nl[k].x, dre);
      dim = fma(l1, nl[k].y, dim);
    }
  }
  if (!st.active) return make_double2(0.0, 0.0);
  return transpose ? make_double2(dre - hre, dim - him) : make_double2(dre + hre, dim + him);
}

/* ---------------------------------------------------------------------------------------------------------
 * Block-level machinery
 * --------------------------------------------------------------------------------------------------------- */
template <int DIMP>
struct Blk {
  using L = Lay<DIMP>;
  unsigned sbase;     /* 32-bit shared address of the dynamic shared memory */
  unsigned own;       /* tid * 16: byte offset of the own element in a vector */
  unsigned giv;       /* shared address of this warp's Givens scratch */
  unsigned hh;        /* shared address of the shared Hessenberg (thread 0 writes) */
  int tid, lane, warp;
  unsigned red;       /* shared address of the reduction half the NEXT reduction uses (toggles) */
  __device__ __forceinline__ void toggle() { red = (red == sbase + L::RED) ? sbase + L::RED + L::RED_HALF : sbase + L::RED; }
  __device__ __forceinline__ unsigned X() const { return sbase + L::X; }
  __device__ __forceinline__ unsigned Y() const { return sbase + L::Y; }
  __device__ __forceinline__ unsigned V(int j) const { return sbase + L::V + (unsigned)j * L::VEC; }
};

/* butterfly over the warp for NV values at once (instruction-level parallelism across the values) */
template <int NV>
__device__ __forceinline__ void warp_sum_n(double (&v)[NV]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}

/* sum of the NW warp partials stored contiguously at addr (fixed tree => every thread gets the same bits) */
template <int NW>
__device__ __forceinline__ double partial_sum(unsigned addr) {
  if constexpr (NW == 1) return lds1(addr);
  else {
    double2 p[NW / 2];
#pragma unroll
    for (int w = 0; w < NW / 2; w++) p[w] = lds2(addr + 16 * w);
    double s[NW / 2];
#pragma unroll
    for (int w = 0; w < NW / 2; w++) s[w] = p[w].x + p[w].y;
#pragma unroll
    for (int stride = 1; stride < NW / 2; stride *= 2) {
#pragma unroll
      for (int w = 0; w + stride < NW / 2; w += 2 * stride) s[w] += s[w + stride];
    }
    return s[0];
  }
}

/* Sum of NV values over the CTA; every thread gets the same bits.  One barrier. */
template <int DIMP, int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], Blk<DIMP>& b) {
  constexpr int NW = Lay<DIMP>::NW;
  static_assert(NV <= kRedStride, "reduction width");
  /* recursive-halving butterfly (qb_mono.cuh): about one shuffle per value instead of five; value `slot` completes on one lane */
  const double mine = qmono::Halver<NV, 16>::run(v, b.lane);
  const int slot = qmono::halver_slot<NV>(b.lane);
  const unsigned r = b.red;
  if (slot >= 0) sts1(r + (unsigned)(slot * NW + b.warp) * 8u, mine);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = partial_sum<NW>(r + (unsigned)(i * NW) * 8u);
  b.toggle();
}
template <int DIMP>
__device__ __forceinline__ double block_sum(double x, Blk<DIMP>& b) {
  double v[1] = {x};
  block_sum_n<DIMP, 1>(v, b);
  return v[0];
}

struct SolveStats { int iters, evals; };

struct SolvePar {
  int maxiter, mcap;
  double abstol, reltol;
};

/* GMRES on (I - alpha*M) y = rhs  (or the transposed operator), zero initial guess.
 * Same algorithm and stopping rule as the reference's KSP set-up (src/timestepper.cpp:541-550: KSPGMRES, PCNONE,
 * rtol 1e-20, abstol 1e-10, maxit = linearsolver_maxiter; PETSc defaults restart 30, classical Gram-Schmidt, no
 * refinement, KSPConvergedDefault on the recurrence residual).  rhs: own element.  Returns own element of y. */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__device__ __forceinline__ double2 gmres_solve(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const SolvePar& P,
                                               Blk<DIMP>& b, double2 rhs, double alpha, bool transpose, SolveStats& stats) {
  constexpr int NW = Lay<DIMP>::NW;
  const unsigned gcc = b.giv, gss = b.giv + 256, ggrs = b.giv + 512, ghc = b.giv + 768, gnrs = b.giv + 1024, grd = b.giv + 1280;
  constexpr unsigned ldh = kLdh * 8;
  double2 y = make_double2(0.0, 0.0);
  int its = 0;
  bool converged = false, first = true;
  double ttol = 0.0;
  while (!converged && its < P.maxiter) {
    double2 r;
    if (first) r = rhs;
    else {
      /* restart: r = b - A y */
      sts2(b.X() + b.own, y);
      __syncthreads();
      const double2 my = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.X(), y, transpose);
      stats.evals++;
      r.x = rhs.x - (y.x - alpha * my.x);
      r.y = rhs.y - (y.y - alpha * my.y);
    }
    const double res2 = block_sum<DIMP>(r.x * r.x + r.y * r.y, b);
    if (res2 == 0.0) { converged = true; break; }
    const double rinv = fast_rsqrt(res2);   /* PETSc: VecScale(VEC_VV(0), 1.0/res_norm) */
    const double res = res2 * rinv;
    if (first) { ttol = fmax(P.reltol * res, P.abstol); first = false; }
    sts2(b.V(0) + b.own, make_double2(r.x * rinv, r.y * rinv));
    if (b.lane == 0) sts1(ggrs, res);
    if (res <= ttol) { converged = true; break; }
    __syncthreads();
    int it = 0;
    while (!converged && it < P.mcap && its < P.maxiter) {
      /* w = A v_it */
      const unsigned vin = b.V(it);
      const double2 vown = lds2(vin + b.own);
      const double2 mv = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, vin, vown, transpose);
      stats.evals++;
      double2 w = make_double2(vown.x - alpha * mv.x, vown.y - alpha * mv.y);
      /* classical Gram-Schmidt: all it+1 inner products from the same w, four at a time */
      const unsigned red = b.red;
      for (int j0 = 0; j0 <= it; j0 += 4) {
        double d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = (j0 + u <= it) ? j0 + u : it;
          const double2 vj = lds2(b.V(j) + b.own);
          d[u] = w.x * vj.x + w.y * vj.y;
        }
        warp_sum_n<4>(d);
        if (b.lane == 0) {
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (j0 + u <= it) sts1(red + (unsigned)((j0 + u) * NW + b.warp) * 8u, d[u]);
        }
      }
      __syncthreads();
      for (int j0 = 0; j0 <= it; j0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (j0 + u <= it) {
            const int j = j0 + u;
            const double h = partial_sum<NW>(red + (unsigned)(j * NW) * 8u);
            const double2 vj = lds2(b.V(j) + b.own);
            w.x -= h * vj.x;
            w.y -= h * vj.y;
            if (b.lane == 0) sts1(ghc + 8u * j, h);
          }
        }
      }
      b.toggle();
      const double tt2 = block_sum<DIMP>(w.x * w.x + w.y * w.y, b);
      double tt = 0.0;
      if (tt2 > 0.0) { const double tinv = fast_rsqrt(tt2); tt = tt2 * tinv; w.x *= tinv; w.y *= tinv; }   /* VecScale(VEC_VV(it+1), 1/tt) */
      if (it + 1 < P.mcap) sts2(b.V(it + 1) + b.own, w);   /* the last basis vector of a cycle is never read */
      /* KSPGMRESUpdateHessenberg on this warp's private copy: previous rotations, new rotation, residual estimate */
      int flag = 0;
      if (b.lane == 0) {
        double hprev = lds1(ghc);
        for (int j = 0; j < it; j++) {
          const double cj = lds1(gcc + 8u * j), sj = lds1(gss + 8u * j);
          const double hnext = lds1(ghc + 8u * (j + 1));
          const double t = hprev;
          const double hj = cj * t + sj * hnext;
          hprev = -sj * t + cj * hnext;
          if (b.warp == 0) sts1(b.hh + it * ldh + 8u * j, hj);
        }
        const double hit = hprev, hit1 = tt;
        const double t2 = hit * hit + hit1 * hit1;
        if (t2 == 0.0) flag = 2;  /* breakdown: converged */
        else {
          const double rt = fast_rsqrt(t2);
          const double ci = hit * rt, si = hit1 * rt;
          const double g = lds1(ggrs + 8u * it);
          const double gnext = -si * g;
          sts1(gcc + 8u * it, ci);
          sts1(gss + 8u * it, si);
          sts1(ggrs + 8u * (it + 1), gnext);
          sts1(ggrs + 8u * it, ci * g);
          sts1(grd + 8u * it, rt);   /* 1 / R(it,it):  R(it,it) = ci*hit + si*hit1 = sqrt(t2) */
          flag = (fabs(gnext) <= ttol) ? 1 : 0;
        }
      }
      flag = __shfl_sync(0xffffffffu, flag, 0);
      __syncthreads();   /* V[it+1] (and thread 0's Hessenberg column) visible */
      it++;
      its++;
      if (flag != 0) converged = true;
    }
    /* KSPGMRESBuildSoln: back substitution (lane 0 of every warp, redundantly), combination by everyone */
    if (it > 0) {
      if (b.lane == 0) {
        for (int kk = it - 1; kk >= 0; kk--) {
          double tsum = lds1(ggrs + 8u * kk);
          for (int j = kk + 1; j < it; j++) tsum -= lds1(b.hh + j * ldh + 8u * kk) * lds1(gnrs + 8u * j);
          sts1(gnrs + 8u * kk, tsum * lds1(grd + 8u * kk));
        }
      }
      __syncwarp();
      for (int j = 0; j < it; j++) {
        const double2 vj = lds2(b.V(j) + b.own);
        const double cj = lds1(gnrs + 8u * j);
        y.x += cj * vj.x;
        y.y += cj * vj.y;
      }
    }
  }
  stats.iters += its;
  return y;
}

/* ImplMidpoint::NeumannSolve, src/timestepper.cpp:697-727: y = rhs + alpha*M*y, stop on ||dy|| < abstol or
 * ||dy||/||dy_0|| < reltol or maxiter. */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__device__ __forceinline__ double2 neumann_solve(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const SolvePar& P,
                                                 Blk<DIMP>& b, double2 rhs, double alpha, bool transpose, SolveStats& stats) {
  double2 y = rhs;
  unsigned cur = b.V(0), nxt = b.V(1);
  sts2(cur + b.own, y);
  __syncthreads();
  double errnorm0 = 0.0;
  int iter;
  for (iter = 0; iter < P.maxiter; iter++) {
    const double2 my = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, cur, y, transpose);
    stats.evals++;
    const double2 yn = make_double2(1.0 * rhs.x + alpha * my.x, 1.0 * rhs.y + alpha * my.y);
    const double ex = y.x - yn.x, ey = y.y - yn.y;
    y = yn;
    sts2(nxt + b.own, y);
    const double err2 = block_sum<DIMP>(ex * ex + ey * ey, b);
    const double errnorm = err2 > 0.0 ? err2 * fast_rsqrt(err2) : 0.0;
    const unsigned t = cur; cur = nxt; nxt = t;
    if (iter == 0) errnorm0 = errnorm;
    if (errnorm < P.abstol) break;
    if (errnorm < P.reltol * errnorm0) break;
  }
  stats.iters += iter;
  return y;
}

/* One-reduction GMRES (qb_mono.cuh) for one element per thread: basis u_j = (alpha M)^j rhs in the V slots, Gram partial sums from
 * the own element, ONE block reduction, scalar L D L^T recurrence in registers, combination.  Returns false if the solve has to be
 * run by gmres_solve (more than kSMax iterations, ill-conditioned basis, breakdown). */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__device__ __forceinline__ bool gmres_mono(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const SolvePar& P,
                                           Blk<DIMP>& b, double2 rhs, double alpha, bool transpose, SolveStats& stats, int& spred, double2& y) {
  using namespace qmono;
  const int mmax = P.maxiter < kSMax ? P.maxiter : kSMax;
  constexpr int KB = kSBlk;          /* measured on C1 / C2: a block of kSMax vectors (28 Gram entries) is 4 % slower than 4 + extension rounds */
  constexpr int NG = (KB + 1) * (KB + 2) / 2;
  int s = spred < 1 ? 1 : spred;
  if (s > KB) s = KB;
  if (s > mmax) s = mmax;
  y = make_double2(0.0, 0.0);
  if (mmax < 1) return true;
  sts2(b.V(0) + b.own, rhs);
  double2 uc = rhs;
#pragma unroll 1
  for (int nb = 1; nb <= s; nb++) {
    __syncthreads();
    const double2 un = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.V(nb - 1), uc, transpose);
    uc = make_double2(alpha * un.x, alpha * un.y);
    sts2(b.V(nb) + b.own, uc);
  }
  double g[NG];
  {
    double2 uu[KB + 1];
#pragma unroll
    for (int j = 0; j <= KB; j++) uu[j] = j <= s ? lds2(b.V(j) + b.own) : make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j <= KB; j++) {
#pragma unroll
      for (int i = 0; i <= j; i++) g[j * (j + 1) / 2 + i] = fma(uu[j].x, uu[i].x, uu[j].y * uu[i].y);
    }
  }
  if (s <= 3) {
    double g10[10];
#pragma unroll
    for (int i = 0; i < 10; i++) g10[i] = g[i];
    block_sum_n<DIMP, 10>(g10, b);
#pragma unroll
    for (int i = 0; i < 10; i++) g[i] = g10[i];
  } else block_sum_n<DIMP, NG>(g, b);
  const double g00 = g[0];
  const double tol2 = fmax(P.reltol * P.reltol * g00, P.abstol * P.abstol);
  if (!(g00 > tol2)) return true;
  Mono M;
  ldl_row<0, 0>(M, g);
  int m = -1;
  bool bail = false;
  static_for<1, KB + 1>([&](auto rc) {
    constexpr int r = decltype(rc)::value;
    if (m < 0 && !bail && r <= s) {
      const bool ok = ldl_row<r, r * (r + 1) / 2>(M, g);
      if (!ok) bail = true;
      else if (M.sigma * tol2 >= 1.0 || r == P.maxiter) m = r;
      else if (r == kSMax) bail = true;
    }
  });
#pragma unroll 1
  for (int r = s + 1; r <= mmax && m < 0 && !bail; r++) {
    const double2 un = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.V(r - 1), uc, transpose);     /* V(r-1): stored before the previous reduction's barrier */
    uc = make_double2(alpha * un.x, alpha * un.y);
    sts2(b.V(r) + b.own, uc);
    double gr[kSMax + 2];
#pragma unroll
    for (int i = 0; i <= kSMax + 1; i++) gr[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= kSMax; i++) {
      if (i < r) { const double2 ui = lds2(b.V(i) + b.own); gr[i] = fma(uc.x, ui.x, uc.y * ui.y); }
      else if (i == r) gr[i] = fma(uc.x, uc.x, uc.y * uc.y);
    }
    block_sum_n<DIMP, kSMax + 2>(gr, b);
    bool ok = false;
    static_for<2, kSMax + 1>([&](auto rc) {
      constexpr int rr = decltype(rc)::value;
      if (r == rr) ok = ldl_row<rr, 0>(M, gr);
    });
    if (!ok) bail = true;
    else if (M.sigma * tol2 >= 1.0 || r == P.maxiter) m = r;
    else if (r == kSMax) bail = true;
  }
  if (bail || m < 0) return false;
  double cf[kSMax];
  mono_coeffs(M, m, cf);
#pragma unroll
  for (int j = 0; j < kSMax; j++) {
    if (j < m) {
      const double2 uj = lds2(b.V(j) + b.own);
      y.x = fma(cf[j], uj.x, y.x); y.y = fma(cf[j], uj.y, y.y);
    }
  }
  stats.iters += m;
  stats.evals += m;
  spred = m;
  return true;
}

/* spred: length prediction of the one-reduction solver carried from solve to solve (>= 1), or -(solves that still skip it) */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__device__ __forceinline__ double2 linsolve(const Stencil<Q, ALL2, HASJ, LIND>& st, const StepCoef<Q, ALL2, HASJ, LIND>& c, const SolvePar& P,
                                            bool gm, Blk<DIMP>& b, double2 rhs, double alpha, bool transpose, SolveStats& stats, int& spred) {
  if (!gm) return neumann_solve<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, b, rhs, alpha, transpose, stats);
  if (spred >= 1) {
    double2 y;
    if (gmres_mono<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, b, rhs, alpha, transpose, stats, spred, y)) return y;
    spred = -7;
    __syncthreads();
  } else {
    spred = spred == 0 ? qmono::kSBlk : spred + 1;
  }
  return gmres_solve<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, b, rhs, alpha, transpose, stats);
}

template <int DIMP>
__device__ __forceinline__ void blk_init(Blk<DIMP>& b, int mcap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  b.sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
  b.tid = threadIdx.x; b.lane = threadIdx.x & 31; b.warp = threadIdx.x >> 5;
  b.own = threadIdx.x * 16u;
  b.giv = b.sbase + Lay<DIMP>::GIV + (unsigned)b.warp * Lay<DIMP>::GIV_WARP;
  b.hh = b.sbase + Lay<DIMP>::V + (unsigned)Lay<DIMP>::nvec(mcap) * Lay<DIMP>::VEC;
  b.red = b.sbase + Lay<DIMP>::RED;
}

/* Forward sweep of one initial condition per CTA: TimeStepper::solveODE, src/timestepper.cpp:96-181, with
 * ImplMidpoint::evolveFWD (src/timestepper.cpp:584-629): x <- x + dtn*k, (I - dtn/2 M) k = M x. */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__global__ void __launch_bounds__(DIMP, (DIMP >= 256 ? 2 : 4)) fwd_sweep_kernel(const __grid_constant__ DevSys S, const __grid_constant__ SweepArgs A) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  Blk<DIMP> b;
  blk_init<DIMP>(b, A.mcap);
  const int ic = blockIdx.x;
  ST st;
  stencil_init(st, S, b.tid);
  StepCoef<Q, ALL2, HASJ, LIND> c;
  SolveStats stats = {0, 0};
  const SolvePar P = {S.maxiter, A.mcap, S.abstol, S.reltol};
  const bool gm = S.linsolve == QB_LINSOLVE_GMRES;
  const int ntime = S.ntime, tabstride = S.tabstride;

  int spred = (S.flags & QB_FLAG_ARNOLDI) ? -(1 << 30) : 3;
  double2 x = A.state[(size_t)ic * DIMP + b.tid];
  double2* traj = A.store ? A.traj + (size_t)ic * (ntime + 1) * DIMP + b.tid : nullptr;
  double2* stg = A.stage_traj ? A.stage_traj + (size_t)ic * ntime * DIMP + b.tid : nullptr;
  double row[ST::ROW];
  row_load<ST::ROW>(row, A.tab);
  for (int n = 0; n < ntime; n++) {
    stepcoef_set(c, st, row);
    const double dtn = row[0];
    if (n + 1 < ntime) row_load<ST::ROW>(row, A.tab + (size_t)(n + 1) * tabstride);   /* prefetch, consumed next step */
    if (traj) traj[(size_t)n * DIMP] = x;
    sts2(b.X() + b.own, x);
    __syncthreads();
    const double2 rhs = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.X(), x, false);
    stats.evals++;
    if (S.ee) {
      /* ExplEuler::evolveFWD, src/timestepper.cpp:493-504: x <- x + dt A(tstart) x (the row holds the controls at t_n) */
      x = make_double2(x.x + dtn * rhs.x, x.y + dtn * rhs.y);
      __syncthreads();     /* X is rewritten at the top of the next step */
      continue;
    }
    const double2 k = linsolve<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, gm, b, rhs, dtn / 2.0, false, stats, spred);
    /* the stage ImplMidpoint::evolveBWD re-solves for (src/timestepper.cpp:674-680), kept for the adjoint sweep */
    if (stg) stg[(size_t)n * DIMP] = make_double2(x.x + dtn / 2.0 * k.x, x.y + dtn / 2.0 * k.y);
    x = make_double2(x.x + dtn * k.x, x.y + dtn * k.y);
  }
  if (traj) traj[(size_t)ntime * DIMP] = x;
  A.state_out[(size_t)ic * DIMP + b.tid] = x;
  if (b.tid == 0) {
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* Adjoint sweep of one initial condition per CTA: TimeStepper::solveAdjointODE + ImplMidpoint::evolveBWD,
 * src/timestepper.cpp:184-253,631-694, with compute_dRHS_dParams_matfree (src/mastereq.cpp:970-1276) and
 * Oscillator::evalControl_diff (src/oscillator.cpp:339-381) fused into the step. */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__global__ void __launch_bounds__(DIMP, (DIMP >= 256 ? 2 : 4)) adj_sweep_kernel(const __grid_constant__ DevSys S, const __grid_constant__ SweepArgs A) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  constexpr int NW = Lay<DIMP>::NW;
  Blk<DIMP> b;
  blk_init<DIMP>(b, A.mcap);
  const int ic = blockIdx.x;
  ST st;
  stencil_init(st, S, b.tid);
  StepCoef<Q, ALL2, HASJ, LIND> c;
  SolveStats stats = {0, 0};
  const SolvePar P = {S.maxiter, A.mcap, S.abstol, S.reltol};
  const bool gm = S.linsolve == QB_LINSOLVE_GMRES;
  const bool recompute = (S.flags & QB_FLAG_RECOMPUTE_STAGE) != 0;
  const int ntime = S.ntime, tabstride = S.tabstride, nd = S.nd;
  const unsigned sgrad = b.hh + (unsigned)(kLdh * (A.mcap + 1)) * 8u;

  for (int i = b.tid; i < nd; i += DIMP) sts1(sgrad + 8u * i, 0.0);
  /* gradient updater role of this thread: (oscillator uk, carrier uf of it, j-th non-zero basis function) */
  int uk = -1, uc = 0, uf = 0, uj = 0, uoff = 0;
  if (b.tid < S.nupd) {
    uk = A.upd[4 * b.tid]; uc = A.upd[4 * b.tid + 1]; uf = A.upd[4 * b.tid + 2]; uj = A.upd[4 * b.tid + 3];
    uoff = S.osc_off[uk];
  }
  const int ncar_total = S.ncar_total;

  int spred = (S.flags & QB_FLAG_ARNOLDI) ? -(1 << 30) : 3, spred_f = spred, spred_b = spred;
  double2 xadj = A.state[(size_t)ic * DIMP + b.tid];
  const double2* traj = A.traj ? A.traj + (size_t)ic * (ntime + 1) * DIMP + b.tid : nullptr;
  const double2* stg = A.stage_traj ? A.stage_traj + (size_t)ic * ntime * DIMP + b.tid : nullptr;
  double2 xcur, xprev;
  if (traj) {
    xcur = traj[(size_t)ntime * DIMP];
    xprev = traj[(size_t)(ntime - 1) * DIMP];
  } else {
    xcur = A.final_state[(size_t)ic * DIMP + b.tid];
    xprev = xcur;
  }
  /* row of the step n-1 -> n: its midpoint (implicit midpoint rule) or, for explicit Euler, t_n = tstop (row n: ExplEuler::evolveBWD
   * assembles A(tstop), src/timestepper.cpp:506-519) */
  const int roff = S.ee ? 1 : 0;
  double row[ST::ROW];
  row_load<ST::ROW>(row, A.tab + (size_t)(ntime - 1 + roff) * tabstride);
  double2 stage_pref = stg ? stg[(size_t)(ntime - 1) * DIMP] : make_double2(0.0, 0.0);
  __syncthreads();

  for (int n = ntime; n > 0; n--) {
    stepcoef_set(c, st, row);
    const double dtn = S.ee ? S.dt : row[0];
    if (n >= 2) row_load<ST::ROW>(row, A.tab + (size_t)(n - 2 + roff) * tabstride);   /* prefetch */
    if (A.pen.dp_on) {   /* derivative of the dpdm penalty at x_n, src/timestepper.cpp:224,371-442 */
      const double2* tr = A.pen.dp_traj + (size_t)ic * (ntime + 1) * DIMP + b.tid;
      auto pop = [&](int m) { const double2 v = tr[(size_t)m * DIMP]; return v.x * v.x + v.y * v.y; };
      const double2 xn = tr[(size_t)n * DIMP];
      const double p0 = xn.x * xn.x + xn.y * xn.y;
      double acc = 0.0;
      if (n > 1) acc += 2.0 * (pop(n - 2) - 2.0 * pop(n - 1) + p0);
      if (n < ntime) acc += -4.0 * (pop(n - 1) - 2.0 * p0 + pop(n + 1));
      if (n < ntime - 1) acc += 2.0 * (p0 - 2.0 * pop(n + 1) + pop(n + 2));
      const double sc = acc * A.pen.dp_dtinv * A.pen.dp_jbar[ic];
      xadj.x += sc * 2.0 * xn.x;
      xadj.y += sc * 2.0 * xn.y;
    }
    if (A.pen.on) {   /* derivative of the penalty integral at x_n, src/timestepper.cpp:227,302-339 */
      const double jb = A.pen.jbar[ic], cw = A.pen.wdt[n];
      double2 pa = A.pen.a[(size_t)ic * DIMP + b.tid];
      const double2 pd = A.pen.d[b.tid];
      if (A.pen.ov) { const double2 o = A.pen.ov[(size_t)ic * (ntime + 1) + n]; pa = make_double2(o.x * pa.x - o.y * pa.y, o.x * pa.y + o.y * pa.x); }
      double gx = 0.0, gy = 0.0;
      if (A.pen.g) { const double2 pg = A.pen.g[b.tid]; const double gwn = A.pen.gw[n]; gx = 2.0 * gwn * pg.x; gy = 2.0 * gwn * pg.y; }
      xadj.x += jb * (cw * (pa.x + 2.0 * pd.x * xcur.x) + gx * xcur.x);
      xadj.y += jb * (cw * (pa.y + 2.0 * pd.y * xcur.y) + gy * xcur.y);
    }
    /* gradient-table entries of this step for the updater threads (in flight during the solve) */
    int gi_idx = -1, gi_ns = 0;
    double g_bval = 0.0, g_cs = 0.0, g_sn = 0.0;
    if (uk >= 0) {
      const size_t gr = (size_t)(n - 1 + roff);
      const int* gi = A.gidx + (gr * Q + uk) * 4;
      gi_idx = __ldg(gi + uj); gi_ns = __ldg(gi + 3);
      g_bval = __ldg(A.gval + (gr * Q + uk) * 3 + uj);
      g_cs = __ldg(A.gcs + (gr * ncar_total + uc) * 2);
      g_sn = __ldg(A.gcs + (gr * ncar_total + uc) * 2 + 1);
    }
    /* primal state at n-1: from the stored trajectory (Lindblad) or one IMR step backwards (Schroedinger),
     * src/timestepper.cpp:230-231 */
    double2 xnext_prefetch = make_double2(0.0, 0.0);
    const double2 stage_stored = stage_pref;                       /* requested one full step ago */
    if (stg && n >= 2) stage_pref = stg[(size_t)(n - 2) * DIMP];   /* in flight during this step */
    if (traj) {
      if (n >= 2) xnext_prefetch = traj[(size_t)(n - 2) * DIMP];  /* in flight during this step */
    } else {
      sts2(b.X() + b.own, xcur);
      __syncthreads();
      const double2 rhs = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.X(), xcur, false);
      stats.evals++;
      if (S.ee) {
        /* ExplEuler::evolveFWD(tstop, tstart, x): one explicit step backwards with A(t_n), src/timestepper.cpp:231,493-504 */
        xprev = make_double2(xcur.x - dtn * rhs.x, xcur.y - dtn * rhs.y);
        __syncthreads();
      } else {
        const double2 k = linsolve<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, gm, b, rhs, -dtn / 2.0, false, stats, spred_b);
        xprev = make_double2(xcur.x - dtn * k.x, xcur.y - dtn * k.y);
      }
    }
    /* adjoint stage: (I - dtn/2 M)^T kbar = xadj ; kbar *= dtn.  Explicit Euler: kbar = dtn * xadj, and the "stage" of the gradient
     * term is the primal state x_{n-1} (compute_dRHS_dParams(tstop, x, x_adj, dt, grad), src/timestepper.cpp:510-512) */
    double2 kbar = S.ee ? xadj : linsolve<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, gm, b, xadj, dtn / 2.0, true, stats, spred);
    kbar.x *= dtn; kbar.y *= dtn;
    /* primal stage x_{n-1} + dtn/2 k */
    double2 stage;
    if (S.ee) {
      stage = xprev;
    } else if (stg) {
      stage = stage_stored;
    } else if (recompute) {
      sts2(b.X() + b.own, xprev);
      __syncthreads();
      const double2 rhs = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.X(), xprev, false);
      stats.evals++;
      const double2 k = linsolve<Q, ALL2, HASJ, LIND, DIMP>(st, c, P, gm, b, rhs, dtn / 2.0, false, stats, spred_f);
      stage = make_double2(xprev.x + dtn / 2.0 * k.x, xprev.y + dtn / 2.0 * k.y);
    } else {
      stage = make_double2(0.5 * (xprev.x + xcur.x), 0.5 * (xprev.y + xcur.y));
    }
    /* X and Y were last read before the final barrier of the previous step / of the solves above */
    sts2(b.X() + b.own, stage);
    sts2(b.Y() + b.own, kbar);
    __syncthreads();
    /* dRHSdp coefficients: coeff_p[k] = kbar^T (dM/dp_k) stage, coeff_q[k] likewise */
    {
      double cpq[2 * Q];
#pragma unroll
      for (int k = 0; k < Q; k++) {
        double cp = 0.0, cq = 0.0;
#pragma unroll
        for (int s = 0; s < ST::CPO; s++) {
          const int sl = ST::CPO * k + s;
          const double2 nb = lds2(b.X() + st.off[sl]);
          const double tp = nb.y * kbar.x - nb.x * kbar.y;
          const double tq = nb.x * kbar.x + nb.y * kbar.y;
          if (ALL2) { cp += sflip(tp, st.ma[sl]); cq += sflip(tq, st.mb[sl]); }
          else { cp += st.ca[sl] * tp; cq += st.cb[sl] * tq; }
        }
        cpq[2 * k] = st.active ? cp : 0.0;
        cpq[2 * k + 1] = st.active ? cq : 0.0;
      }
      stats.evals++;
      warp_sum_n<2 * Q>(cpq);
      const unsigned red = b.red;
      if (b.lane == 0) {
#pragma unroll
        for (int i = 0; i < 2 * Q; i++) sts1(red + (unsigned)(i * NW + b.warp) * 8u, cpq[i]);
      }
      /* xadj += M^T kbar */
      const double2 mt = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.Y(), kbar, true);
      stats.evals++;
      xadj.x += mt.x; xadj.y += mt.y;
      __syncthreads();
      if (uk >= 0 && gi_idx >= 0) {
        const double cp = partial_sum<NW>(red + (unsigned)((2 * uk) * NW) * 8u);
        const double cq = partial_sum<NW>(red + (unsigned)((2 * uk + 1) * NW) * 8u);
        const double pbar = 1.0 * cp, qbar = 1.0 * cq;
        const double B1bar = g_sn * qbar + g_cs * pbar;
        const double B2bar = g_cs * qbar - g_sn * pbar;
        const unsigned gaddr = sgrad + 8u * (unsigned)(uoff + gi_idx + uf * gi_ns * 2);
        sts1(gaddr, lds1(gaddr) + g_bval * B1bar);
        sts1(gaddr + 8u * gi_ns, lds1(gaddr + 8u * gi_ns) + g_bval * B2bar);
      }
      b.toggle();
    }
    xcur = xprev;
    if (traj) xprev = xnext_prefetch;
  }
  __syncthreads();
  for (int i = b.tid; i < nd; i += DIMP) A.grad_ic[(size_t)ic * nd + i] = lds1(sgrad + 8u * i);
  A.state_out[(size_t)ic * DIMP + b.tid] = xadj;
  if (b.tid == 0) {
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* y = M(t) x or M(t)^T x for a batch of vectors, one CTA per vector (unit-test hook). */
template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
__global__ void __launch_bounds__(DIMP) rhs_apply_kernel(const __grid_constant__ DevSys S, const double* __restrict__ rowg,
                                                        const double2* __restrict__ x, double2* __restrict__ y, int transpose) {
  using ST = Stencil<Q, ALL2, HASJ, LIND>;
  Blk<DIMP> b;
  blk_init<DIMP>(b, 1);
  ST st;
  stencil_init(st, S, b.tid);
  StepCoef<Q, ALL2, HASJ, LIND> c;
  double row[ST::ROW];
  row_load<ST::ROW>(row, rowg);
  stepcoef_set(c, st, row);
  const double2 xo = x[(size_t)blockIdx.x * DIMP + b.tid];
  sts2(b.X() + b.own, xo);
  __syncthreads();
  y[(size_t)blockIdx.x * DIMP + b.tid] = apply_rhs<Q, ALL2, HASJ, LIND>(st, c, b.X(), xo, transpose != 0);
}

template <int Q, bool ALL2, bool HASJ, bool LIND, int DIMP>
void pick4(KernelSet* ks) {
  ks->app = rhs_apply_kernel<Q, ALL2, HASJ, LIND, DIMP>;
  ks->fwd = fwd_sweep_kernel<Q, ALL2, HASJ, LIND, DIMP>;
  ks->adj = adj_sweep_kernel<Q, ALL2, HASJ, LIND, DIMP>;
  ks->dimp = DIMP;
  ks->smem_fixed = Lay<DIMP>::V;
  ks->vec_bytes = Lay<DIMP>::VEC;
}
/* threads per CTA: the padded dimension.  Qubit systems have a single possible size per (Q, LIND). */
template <int Q, bool ALL2, bool HASJ, bool LIND>
void pick3(int dim, KernelSet* ks) {
  if constexpr (ALL2) {
    constexpr int D = LIND ? (1 << (2 * Q)) : (1 << Q);
    if constexpr (D <= kMaxThreads) pick4<Q, ALL2, HASJ, LIND, (D < 32 ? 32 : D)>(ks);
  } else {
    if (dim <= 64) pick4<Q, ALL2, HASJ, LIND, 64>(ks);
    else pick4<Q, ALL2, HASJ, LIND, 256>(ks);
  }
}
template <int Q, bool ALL2>
void pick2(bool hasj, bool lind, int dim, KernelSet* ks) {
  if constexpr (Q >= 2) {
    if (hasj) { if (lind) pick3<Q, ALL2, true, true>(dim, ks); else pick3<Q, ALL2, true, false>(dim, ks); return; }
  }
  if (lind) pick3<Q, ALL2, false, true>(dim, ks); else pick3<Q, ALL2, false, false>(dim, ks);
}

}  // namespace qbk

/* Selects the kernels for a system with Q oscillators (fwd == nullptr: no small-N kernel for this size). */
template <int Q>
void qb_pick(bool all2, bool hasj, bool lind, int dim, KernelSet* ks) {
  ks->fwd = nullptr; ks->adj = nullptr; ks->app = nullptr;
  if (all2) qbk::pick2<Q, true>(hasj, lind, dim, ks);
  else qbk::pick2<Q, false>(hasj, lind, dim, ks);
}
