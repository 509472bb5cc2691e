/*
 * qb_qubit.cuh -- sweep kernels specialised for systems of two-level oscillators (qubits) without Jkl coupling:
 * BASELINE configs C3 (N=8) and C4 (N=16).
 *
 * Same mathematics and the same reference code as qb_kernels.cuh (applyRHS_matfree / _transpose src/mastereq.cpp:1279-2893,
 * ImplMidpoint::evolveFWD/evolveBWD src/timestepper.cpp:584-694, NeumannSolve :697-727, KSPGMRES set-up :541-550,
 * compute_dRHS_dParams_matfree src/mastereq.cpp:970-1276, Oscillator::evalControl_diff src/oscillator.cpp:339-381),
 * laid out for latency instead of generality:
 *
 *   - For qubits every stencil neighbour of element `it` of vec(rho) is `it XOR mask` (flip the bra or the ket bit of
 *     one oscillator; the decay term flips both).  A thread owns E elements that differ in the TOP log2(E) index bits,
 *     so neighbours across those bits are register moves; the others are one LDS.128 at `own ^ const`.
 *     Signs and decay coefficients become per-thread constants or compile-time facts; no per-element tables.
 *   - T = dim/E threads per initial condition.  A sweep is a chain of dependent CTA-wide phases, so the linear solver is
 *     organised to need ONE reduction per solve (see "the one-reduction linear solver" below) instead of two per Krylov
 *     iteration, and the adjoint step's M^T kbar falls out of the solve without another operator application.
 *   - The trajectory leaves / enters the SM through the TMA bulk-copy engine (cp.async.bulk + mbarrier): the forward
 *     sweep's exchange buffers ARE the sources of the bulk stores, the adjoint sweep prefetches its primal input (the stored
 *     stage, or x_{n-1}) one step ahead into a two-slot shared-memory ring and reads stencil neighbours straight from the
 *     slot.  The per-step control record (dt, p_k, q_k, carrier phases, B-spline values/indices) is streamed the same way,
 *     so a step issues no global load at all.
 *   - Warp reductions use recursive halving (about one shuffle per reduced value instead of five).
 */
#pragma once
#include "qb_kernels.cuh"
#include "qb_mono.cuh"
#include "qb_qubit_args.h"

namespace qq {

using namespace qmono;

/* -DQB_PHASE_CLOCKS: thread 0 of CTA 0 accumulates the cycles between the marks below and prints them at the end of the sweep
 * (a development aid for the latency chain of one step; off in the shipped library) */
#ifdef QB_PHASE_CLOCKS
}  // namespace qq
#include <cstdio>
namespace qq {
struct PhaseClk { long long acc[12]; long long last; };
__device__ PhaseClk g_pclk;
#define QB_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_pclk.acc[i] += t_ - g_pclk.last; g_pclk.last = t_; } } while (0)
#define QB_CLK_INIT() do { if (threadIdx.x == 0 && blockIdx.x == 0) { for (int i_ = 0; i_ < 12; i_++) g_pclk.acc[i_] = 0; g_pclk.last = clock64(); } } while (0)
#define QB_CLK_PRINT(tag, n) do { if (threadIdx.x == 0 && blockIdx.x == 0) { printf("%s cycles/step:", tag); for (int i_ = 0; i_ < 12; i_++) printf(" [%d] %.0f", i_, (double)g_pclk.acc[i_] / (n)); printf("\n"); } } while (0)
#else
#define QB_CLK(i) do { } while (0)
#define QB_CLK_INIT() do { } while (0)
#define QB_CLK_PRINT(tag, n) do { } while (0)
#endif

using qbk::fast_rsqrt;
using qbk::lds1;
using qbk::lds2;
using qbk::SolvePar;
using qbk::SolveStats;
using qbk::sts1;
using qbk::sts2;

constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

template <int Q, bool LIND, int E>
struct K {
  static constexpr int NB = LIND ? 2 * Q : Q;   /* index bits */
  static constexpr int DIM = 1 << NB;
  static constexpr int T = DIM / E;             /* threads per initial condition */
  static constexpr int LT = ilog2(T);           /* bits [0,LT): thread bits, [LT,NB): element bits */
  static constexpr int NW = T / 32;
  static constexpr unsigned VEC = DIM * 16;
  static constexpr unsigned ESTR = T * 16;      /* byte distance of consecutive elements of a thread */
  static_assert(T >= 32 && T % 32 == 0 && E * T == DIM, "bad elements-per-thread");
};

/* ---- TMA bulk copy / mbarrier primitives ------------------------------------------------------------------ */
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "QQ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra QQ_DONE;\n"
      "bra QQ_WAIT;\n"
      "QQ_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int NW>
__device__ __forceinline__ void csync() {
  if constexpr (NW == 1) __syncwarp();
  else __syncthreads();
}

/* ---- per-thread constants ----------------------------------------------------------------------------------- */
template <int Q, bool LIND, int E>
struct Th {
  using C = K<Q, LIND, E>;
  unsigned own;                         /* t*16 */
  unsigned nboff[C::LT];                /* (t ^ (1<<m))*16 for the thread bits m */
  unsigned l1off[LIND ? Q : 1];         /* (t ^ thread-bit part of the decay pair of oscillator k)*16 */
  double dA[E], dB[E];                  /* diagonal: Hamiltonian (hd - hd'), dissipative (l2 + l1diag) */
  double gf[LIND ? Q : 1], gt[LIND ? Q : 1];   /* decay_k if the thread bits of pair k are all 0 (gf) / all 1 (gt), else 0 */
  int tid, lane, warp;
};

template <int Q, bool LIND, int E>
__device__ void th_init(Th<Q, LIND, E>& th, const DevSys& S) {
  using C = K<Q, LIND, E>;
  const int t = threadIdx.x;
  th.tid = t; th.lane = t & 31; th.warp = t >> 5;
  th.own = (unsigned)t * 16u;
#pragma unroll
  for (int m = 0; m < C::LT; m++) th.nboff[m] = (unsigned)(t ^ (1 << m)) * 16u;
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int it = e * C::T + t;
    int i[kMaxOsc], ip[kMaxOsc];
    const int row = LIND ? (it & ((1 << Q) - 1)) : it, col = LIND ? (it >> Q) : 0;
#pragma unroll
    for (int k = 0; k < Q; k++) { i[k] = (row >> (Q - 1 - k)) & 1; ip[k] = (col >> (Q - 1 - k)) & 1; }
    const double hd = qbk::hdiag_dev(S, i);
    const double hdp = LIND ? qbk::hdiag_dev(S, ip) : 0.0;
    th.dA[e] = hd - hdp;
    th.dB[e] = 0.0;
    if (LIND) {
      double l1diag = 0.0, l2 = 0.0;
      for (int k = 0; k < Q; k++) l1diag += -S.decay[k] / 2.0 * (i[k] + ip[k]);
      for (int k = 0; k < Q; k++) l2 += S.dephase[k] * (i[k] * ip[k] - 1. / 2. * (i[k] * i[k] + ip[k] * ip[k]));
      th.dB[e] = l2 + l1diag;
    }
  }
  if (LIND) {
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const int rb = Q - 1 - k, cb = 2 * Q - 1 - k;
      int mask = 0;
      if (rb < C::LT) mask |= 1 << rb;
      if (cb < C::LT) mask |= 1 << cb;
      th.l1off[k] = (unsigned)(t ^ mask) * 16u;
      const double dec = (fabs(S.decay[k]) > 1e-12) ? S.decay[k] : 0.0;   /* L1decay guard, include/mastereq.hpp:744-769 */
      th.gf[k] = ((t & mask) == 0) ? dec : 0.0;
      th.gt[k] = ((t & mask) == mask) ? dec : 0.0;
    }
  }
}

/* per-step weights */
template <int Q, bool LIND, int E>
struct Co {
  using C = K<Q, LIND, E>;
  double p[Q], q[Q];
  double qr[C::LT];    /* q of the oscillator of thread bit m, with the sign of that bit: bit 0 -> +q, bit 1 -> -q */
};

template <int Q, bool LIND>
__device__ __forceinline__ constexpr int osc_of_bit(int m) { return (LIND && m >= Q) ? 2 * Q - 1 - m : Q - 1 - m; }

/* reads dt, p_k, q_k of a step from its record in shared memory (broadcast loads) */
template <int Q, bool LIND, int E>
__device__ __forceinline__ double co_load(Co<Q, LIND, E>& co, const Th<Q, LIND, E>& th, unsigned rec) {
  using C = K<Q, LIND, E>;
  const double dtn = lds1(rec);
#pragma unroll
  for (int k = 0; k < Q; k++) { co.p[k] = lds1(rec + 8u * (1 + k)); co.q[k] = lds1(rec + 8u * (1 + Q + k)); }
#pragma unroll
  for (int m = 0; m < C::LT; m++) {
    const double q = co.q[osc_of_bit<Q, LIND>(m)];
    co.qr[m] = ((th.tid >> m) & 1) ? -q : q;
  }
  return dtn;
}

/* y = M x (TR=false) or M^T x (TR=true) for the E own elements; vs: shared address of the full vector.
 * Same accumulation order per element as qbk::apply_rhs (oscillators in order, bra then ket; zero-coefficient terms dropped).
 * The loops run neighbour-major, element-minor: the 4E accumulator chains advance together, so consecutive dependent FMAs of
 * a chain are 4E instructions apart (a sweep runs with about one warp per scheduler: nothing else hides the fp64 latency). */
template <int Q, bool LIND, int E, bool TR>
__device__ __forceinline__ void apply(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, unsigned vs, const double2 (&x)[E], double2 (&y)[E]) {
  using C = K<Q, LIND, E>;
  /* S accumulator sets per element (S = 2 would run the bra-side and ket-side terms in separate chains, four FMAs deep) */
  constexpr int S = 1;     /* measured with S = 2 at E = 2: no gain (the application phase is bound by the shared-memory pipe and the barriers, not by the FMA chains) */
  double hre[S][E], him[S][E], hre2[S][E], him2[S][E];
#pragma unroll
  for (int e = 0; e < E; e++) {
#pragma unroll
    for (int sd = 0; sd < S; sd++) { hre[sd][e] = 0.0; him[sd][e] = 0.0; hre2[sd][e] = 0.0; him2[sd][e] = 0.0; }
    hre[0][e] = th.dA[e] * x[e].y; him[0][e] = -th.dA[e] * x[e].x;
  }
#pragma unroll
  for (int k = 0; k < Q; k++) {
#pragma unroll
    for (int side = 0; side < (LIND ? 2 : 1); side++) {
      const int m = side == 0 ? Q - 1 - k : 2 * Q - 1 - k;
      const int sd = S == 2 ? side : 0;
      const double a = side == 0 ? co.p[k] : -co.p[k];
      double2 nb[E];
#pragma unroll
      for (int e = 0; e < E; e++) {
        if (m < C::LT) nb[e] = lds2(vs + th.nboff[m < C::LT ? m : 0] + (unsigned)e * C::ESTR);
        else { const int lb = m >= C::LT ? m - C::LT : 0; nb[e] = x[e ^ (1 << lb)]; }
      }
#pragma unroll
      for (int e = 0; e < E; e++) {
        double b;
        if (m < C::LT) b = co.qr[m < C::LT ? m : 0];
        else { const int lb = m >= C::LT ? m - C::LT : 0; b = ((e >> lb) & 1) ? -co.q[k] : co.q[k]; }
        hre[sd][e] = fma(a, nb[e].y, hre[sd][e]);
        hre2[sd][e] = fma(b, nb[e].x, hre2[sd][e]);
        him[sd][e] = fma(-a, nb[e].x, him[sd][e]);
        him2[sd][e] = fma(b, nb[e].y, him2[sd][e]);
      }
    }
  }
  double dre[E], dim[E], hr[E], hi[E];
#pragma unroll
  for (int e = 0; e < E; e++) {
    if constexpr (S == 2) { hr[e] = (hre[0][e] + hre[1][e]) + (hre2[0][e] + hre2[1][e]); hi[e] = (him[0][e] + him[1][e]) + (him2[0][e] + him2[1][e]); }
    else { hr[e] = hre[0][e] + hre2[0][e]; hi[e] = him[0][e] + him2[0][e]; }
    dre[e] = 0.0; dim[e] = 0.0;
  }
  if (LIND) {
#pragma unroll
    for (int e = 0; e < E; e++) { dre[e] = th.dB[e] * x[e].x; dim[e] = th.dB[e] * x[e].y; }
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const int rb = Q - 1 - k, cb = 2 * Q - 1 - k;
      const double l1 = TR ? th.gt[k] : th.gf[k];
#pragma unroll
      for (int e = 0; e < E; e++) {
        /* forward: rho_{i+1,i'+1} feeds rho_{i,i'} (both bits 0 here); transpose: the opposite shift (both bits 1) */
        bool ok = true;
        int emask = 0;
        if (rb >= C::LT) { const int lb = rb >= C::LT ? rb - C::LT : 0; if (((e >> lb) & 1) != (TR ? 1 : 0)) ok = false; emask |= 1 << lb; }
        if (cb >= C::LT) { const int lb = cb >= C::LT ? cb - C::LT : 0; if (((e >> lb) & 1) != (TR ? 1 : 0)) ok = false; emask |= 1 << lb; }
        if (ok) {
          double2 nl;
          if (rb >= C::LT && cb >= C::LT) nl = x[e ^ emask];
          else nl = lds2(vs + th.l1off[k] + (unsigned)(e ^ emask) * C::ESTR);
          dre[e] = fma(l1, nl.x, dre[e]);
          dim[e] = fma(l1, nl.y, dim[e]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; e++) y[e] = TR ? make_double2(dre[e] - hr[e], dim[e] - hi[e]) : make_double2(dre[e] + hr[e], dim[e] + hi[e]);
}

/* ---- CTA-wide sums ------------------------------------------------------------------------------------------ */
template <int NW>
struct Sm {
  unsigned base;      /* shared address of the dynamic shared memory */
  unsigned red;       /* reduction half the NEXT reduction uses */
  unsigned red_other; /* the other half (the two are swapped after every use) */
  unsigned giv;       /* this warp's Givens scratch (fallback solver) */
  unsigned xb, u, vec;
  double2* vg;        /* fallback solver: this initial condition's Arnoldi basis in global memory (thread-private elements) */
  double* hhg;        /* fallback solver: this warp's Hessenberg columns */
};

/* Sum of the first NV entries of v over the CTA; every thread gets the same bits.  One barrier. */
template <int NV, int NW, int NTOT>
__device__ __forceinline__ void allsum(double (&v)[NTOT], Sm<NW>& sm, int lane, int warp) {
  static_assert(NV <= NTOT && NV <= 32 && NV <= kNRed, "reduction width");
  double t[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) t[i] = v[i];
  const double mine = Halver<NV, 16>::run(t, lane);
  const int slot = halver_slot<NV>(lane);
  const unsigned r = sm.red;
  if (slot >= 0) sts1(r + (unsigned)(slot * NW + warp) * 8u, mine);
  csync<NW>();
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = qbk::partial_sum<NW>(r + (unsigned)(i * NW) * 8u);
  sm.red = sm.red_other;
  sm.red_other = r;
}

template <int E>
__device__ __forceinline__ double dot_own(const double2 (&a)[E], const double2 (&b)[E]) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int e = 0; e < E; e++) { s0 = fma(a[e].x, b[e].x, s0); s1 = fma(a[e].y, b[e].y, s1); }
  return s0 + s1;
}

/* Builds the basis and solves.  The first operator application reads the vector v0 (registers) / slot `src0` (shared, visible
 * to the CTA) and writes scale0 * M v0 to U[j0]: the forward step passes x_n, scale0 = 1, j0 = 0, so that b = u_0 = M x_n is
 * the first vector of the same loop (one call site of the operator per kernel keeps the step's code inside the instruction
 * cache); the adjoint step passes b = u_0 (already stored in U[0]), scale0 = alpha, j0 = 1.
 * Returns false if the solve must be redone by the fallback (U[0] = b is intact in either case). */
template <int Q, bool LIND, int E, bool TR, bool WANT_MY>
__device__ __forceinline__ bool gmres_mono(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, Sm<K<Q, LIND, E>::NW>& sm,
                                           const double2 (&v0)[E], unsigned src0, double scale0, int j0, bool sync0, double alpha,
                                           double2 (&y)[E], double2 (&my)[E], SolveStats& stats, int& spred) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  /* longest speculative block: the whole basis for three qubits (dt = 0.05: five to six iterations per solve), four vectors for four
   * qubits (three iterations; the Gram phase of the long block costs registers and instructions there) */
  constexpr int KB = Q <= 3 ? kSMax : kSBlk;     /* measured: the long block gains 9 % on C3 (5.2 iterations per solve) and loses 6 % on C4 (3.0) */
  constexpr int NG = (KB + 1) * (KB + 2) / 2;
  const int mmax = P.maxiter < kSMax ? P.maxiter : kSMax;
  int s = spred < 1 ? 1 : spred;
  if (s > KB) s = KB;
  if (s > mmax) s = mmax;
  if (mmax < 1) s = 0;
  /* u_{j0} .. u_s without any reduction in between */
  double2 uc[E];
#pragma unroll
  for (int e = 0; e < E; e++) uc[e] = v0[e];
  unsigned src = src0;
  double scale = scale0;
  int nb = j0;              /* basis vectors stored so far: u_0 .. u_{nb-1} */
  bool sync = sync0;
#pragma unroll 1
  for (; nb <= s; nb++) {
    if (sync) csync<NW>();      /* the source vector is visible */
    double2 un[E];
    apply<Q, LIND, E, TR>(th, co, src, uc, un);
    const unsigned dst = sm.u + (unsigned)nb * sm.vec;
#pragma unroll
    for (int e = 0; e < E; e++) { uc[e].x = un[e].x * scale; uc[e].y = un[e].y * scale; sts2(dst + th.own + (unsigned)e * C::ESTR, uc[e]); }
    src = dst; scale = alpha; sync = true;
  }
#pragma unroll
  for (int e = 0; e < E; e++) { y[e] = make_double2(0.0, 0.0); my[e] = make_double2(0.0, 0.0); }
  /* Gram rows 0..s from the thread's own elements (re-read from shared memory: no visibility needed) */
  QB_CLK(1);      /* basis built */
  /* rows 0..3 unconditionally (slots beyond s hold finite leftovers of earlier solves -- the slots are zeroed at kernel start --
   * and rows beyond s are never looked at): ten independent accumulation chains instead of branches around each of them */
  double g[NG];
  {
    double2 uu[KB + 1][E];
#pragma unroll
    for (int j = 0; j <= KB; j++) {
#pragma unroll
      for (int e = 0; e < E; e++)
        uu[j][e] = (j < kSBlk || j <= s) ? lds2(sm.u + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j <= KB; j++) {
#pragma unroll
      for (int i = 0; i <= j; i++) g[j * (j + 1) / 2 + i] = (j < kSBlk || j <= s) ? dot_own<E>(uu[j], uu[i]) : 0.0;
    }
  }
  QB_CLK(2);      /* Gram partials */
  if (s <= 3) allsum<10, NW>(g, sm, th.lane, th.warp);
  else if (s == 4 || KB == 4) allsum<15, NW>(g, sm, th.lane, th.warp);
  else if constexpr (KB >= 6) {
    if (s == 5) allsum<21, NW>(g, sm, th.lane, th.warp);
    else allsum<28, NW>(g, sm, th.lane, th.warp);
  }
  QB_CLK(3);      /* reduction */
  if (mmax < 1) return true;
  const double g00 = g[0];
  const double tol2 = fmax(P.reltol * P.reltol * g00, P.abstol * P.abstol);
  if (!(g00 > tol2)) return true;        /* |b| <= tolerance (or b = 0): converged before the first iteration, y = 0 */
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
  QB_CLK(4);      /* L D L^T rows */
  /* not converged with the speculative block: one vector and one reduction at a time */
#pragma unroll 1
  for (int r = s + 1; r <= mmax && m < 0 && !bail; r++) {
    double2 un[E];
    apply<Q, LIND, E, TR>(th, co, src, uc, un);      /* u_{r-1}: stored before the previous reduction's barrier */
    const unsigned dst = sm.u + (unsigned)r * sm.vec;
#pragma unroll
    for (int e = 0; e < E; e++) { uc[e].x = un[e].x * alpha; uc[e].y = un[e].y * alpha; sts2(dst + th.own + (unsigned)e * C::ESTR, uc[e]); }
    src = dst;
    double gr[kSMax + 2];
#pragma unroll
    for (int i = 0; i <= kSMax + 1; i++) gr[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= kSMax; i++) {
      if (i < r) {
        double2 ui[E];
#pragma unroll
        for (int e = 0; e < E; e++) ui[e] = lds2(sm.u + (unsigned)i * sm.vec + th.own + (unsigned)e * C::ESTR);
        gr[i] = dot_own<E>(uc, ui);
      } else if (i == r) gr[i] = dot_own<E>(uc, uc);
    }
    allsum<kSMax + 2, NW>(gr, sm, th.lane, th.warp);
    bool ok = false;
    static_for<2, kSMax + 1>([&](auto rc) {
      constexpr int rr = decltype(rc)::value;
      if (r == rr) ok = ldl_row<rr, 0>(M, gr);
    });
    if (!ok) bail = true;
    else if (M.sigma * tol2 >= 1.0 || r == P.maxiter) m = r;
    else if (r == kSMax) bail = true;
  }
  QB_CLK(5);      /* extension rounds */
  if (bail || m < 0) return false;
  double c[kSMax];
  mono_coeffs(M, m, c);
  QB_CLK(6);      /* coefficients */
#pragma unroll
  for (int j = 0; j <= kSMax; j++) {
    if (j < kSBlk || j <= m) {      /* the first slots unconditionally: their coefficients are zero beyond m */
      const double cy = j < kSMax ? c[j < kSMax ? j : 0] : 0.0;      /* c[j] = 0 for j >= m */
      const double cm = (j >= 1 && j <= m) ? c[j >= 1 ? j - 1 : 0] : 0.0;
#pragma unroll
      for (int e = 0; e < E; e++) {
        const double2 uj = lds2(sm.u + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR);
        y[e].x = fma(cy, uj.x, y[e].x); y[e].y = fma(cy, uj.y, y[e].y);
        if (WANT_MY) { my[e].x = fma(cm, uj.x, my[e].x); my[e].y = fma(cm, uj.y, my[e].y); }
      }
    }
  }
  QB_CLK(7);      /* combination */
  stats.iters += m;
  stats.evals += m;
  spred = m;
  return true;
}

/* ImplMidpoint::NeumannSolve (src/timestepper.cpp:697-727) in the same basis: iterate k+1 is sum_{j<=k+1} u_j and the
 * difference of consecutive iterates is u_{k+1}, so the stopping test needs the norms |u_j| only -- reduced once for the
 * speculative block, then one at a time.  Arguments as gmres_mono; ring slots 0..kSMax-1 hold u_j at j % kSMax. */
template <int Q, bool LIND, int E, bool TR, bool WANT_MY>
__device__ __forceinline__ void neumann_mono(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, Sm<K<Q, LIND, E>::NW>& sm,
                                             const double2 (&v0)[E], unsigned src0, double scale0, int j0, bool sync0, double alpha,
                                             double2 (&y)[E], double2 (&my)[E], SolveStats& stats, int& spred) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  constexpr int NS = kSMax;
  int s = spred < 1 ? 1 : spred;
  if (s > kSBlk) s = kSBlk;
  if (s > P.maxiter) s = P.maxiter;
  if (s < 0) s = 0;
  double2 uc[E];
#pragma unroll
  for (int e = 0; e < E; e++) uc[e] = v0[e];
  unsigned src = src0;
  double scale = scale0;
  int nb = j0;
  bool sync = sync0;
#pragma unroll 1
  for (; nb <= s; nb++) {
    if (sync) csync<NW>();
    double2 un[E];
    apply<Q, LIND, E, TR>(th, co, src, uc, un);
    const unsigned dst = sm.u + (unsigned)nb * sm.vec;
#pragma unroll
    for (int e = 0; e < E; e++) { uc[e].x = un[e].x * scale; uc[e].y = un[e].y * scale; sts2(dst + th.own + (unsigned)e * C::ESTR, uc[e]); }
    src = dst; scale = alpha; sync = true;
  }
  /* y = u_0; norms of u_1..u_s */
  double g[kSBlk];
#pragma unroll
  for (int e = 0; e < E; e++) y[e] = lds2(sm.u + th.own + (unsigned)e * C::ESTR);
  double2 uu[kSBlk][E];
#pragma unroll
  for (int j = 1; j <= kSBlk; j++) {
#pragma unroll
    for (int e = 0; e < E; e++) uu[j - 1][e] = (j <= s) ? lds2(sm.u + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR) : make_double2(0.0, 0.0);
    g[j - 1] = dot_own<E>(uu[j - 1], uu[j - 1]);
  }
  allsum<kSBlk, NW>(g, sm, th.lane, th.warp);
  int m = 0;               /* u_1..u_m are part of the result */
  bool done = P.maxiter < 1;
  int iters = 0;
  double err0 = 0.0;
#pragma unroll
  for (int j = 1; j <= kSBlk; j++) {
    if (!done && j <= s) {
      const double err = g[j - 1] > 0.0 ? g[j - 1] * fast_rsqrt(g[j - 1]) : 0.0;
      if (j == 1) err0 = err;
      m = j;
      iters = j - 1;
#pragma unroll
      for (int e = 0; e < E; e++) { y[e].x += uu[j - 1][e].x; y[e].y += uu[j - 1][e].y; }
      if (err < P.abstol || err < P.reltol * err0) done = true;
      else if (j == P.maxiter) { done = true; iters = P.maxiter; }
    }
  }
  /* beyond the block (done is false only if m == s, so uc holds u_m and src its slot) */
#pragma unroll 1
  while (!done) {
    const int j = m + 1;
    double2 un[E];
    apply<Q, LIND, E, TR>(th, co, src, uc, un);
    const unsigned dst = sm.u + (unsigned)(j % NS) * sm.vec;
#pragma unroll
    for (int e = 0; e < E; e++) { uc[e].x = un[e].x * alpha; uc[e].y = un[e].y * alpha; sts2(dst + th.own + (unsigned)e * C::ESTR, uc[e]); }
    src = dst;
    double g1[1] = {dot_own<E>(uc, uc)};
    allsum<1, NW>(g1, sm, th.lane, th.warp);
    const double err = g1[0] > 0.0 ? g1[0] * fast_rsqrt(g1[0]) : 0.0;
#pragma unroll
    for (int e = 0; e < E; e++) { y[e].x += uc[e].x; y[e].y += uc[e].y; }
    m = j;
    iters = j - 1;
    if (err < P.abstol || err < P.reltol * err0) done = true;
    else if (j == P.maxiter) { done = true; iters = P.maxiter; }
  }
  stats.iters += iters;
  stats.evals += m;
  spred = m;
  if (WANT_MY) {
    const unsigned xs = sm.u + (unsigned)kSMax * sm.vec;
    csync<NW>();
#pragma unroll
    for (int e = 0; e < E; e++) sts2(xs + th.own + (unsigned)e * C::ESTR, y[e]);
    csync<NW>();
    apply<Q, LIND, E, TR>(th, co, xs, y, my);
#pragma unroll
    for (int e = 0; e < E; e++) { my[e].x *= alpha; my[e].y *= alpha; }
  }
}

/* ---- Arnoldi fallback: restarted GMRES with classical Gram-Schmidt and the Givens residual recurrence, the reference's
 * KSPGMRES defaults (src/timestepper.cpp:541-550), for solves the one-reduction path hands back (more than kSMax
 * iterations, ill-conditioned monomial basis, happy breakdown).  The basis lives in global memory: a thread only ever
 * touches the elements it owns, so no visibility question arises; the current vector is mirrored in two alternating
 * shared slots for the neighbour reads of the operator. */
template <int Q, bool LIND, int E, bool TR>
__device__ __forceinline__ void gmres_g(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, Sm<K<Q, LIND, E>::NW>& sm,
                                        const double2 (&rhs)[E], double alpha, double2 (&y)[E], SolveStats& stats) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  const unsigned gcc = sm.giv, gss = sm.giv + 256, ggrs = sm.giv + 512, grd = sm.giv + 768, gnrs = sm.giv + 1024, ghc = sm.giv + 1280;
  const int ldh = P.mcap;
  const bool l0 = th.lane == 0;
  double2* const vg = sm.vg + th.tid;
  double* const hh = sm.hhg;
  const unsigned xrs = sm.u + 3u * sm.vec;
#pragma unroll
  for (int e = 0; e < E; e++) y[e] = make_double2(0.0, 0.0);
  int its = 0;
  bool converged = false, first = true;
  double ttol = 0.0;
  csync<NW>();       /* whatever the caller still reads from the exchange slots is done */
  while (!converged && its < P.maxiter) {
    double2 v[E];
    if (first) {
#pragma unroll
      for (int e = 0; e < E; e++) v[e] = rhs[e];
    } else {
      /* restart: r = b - A y */
#pragma unroll
      for (int e = 0; e < E; e++) sts2(xrs + th.own + (unsigned)e * C::ESTR, y[e]);
      csync<NW>();
      double2 my[E];
      apply<Q, LIND, E, TR>(th, co, xrs, y, my);
      stats.evals++;
#pragma unroll
      for (int e = 0; e < E; e++) { v[e].x = rhs[e].x - (y[e].x - alpha * my[e].x); v[e].y = rhs[e].y - (y[e].y - alpha * my[e].y); }
    }
    double r2[1] = {dot_own<E>(v, v)};
    allsum<1, NW>(r2, sm, th.lane, th.warp);
    const double res2 = r2[0];
    if (res2 == 0.0) { converged = true; break; }
    const double rinv = fast_rsqrt(res2);
    const double res = res2 * rinv;
    if (first) { ttol = fmax(P.reltol * res, P.abstol); first = false; }
#pragma unroll
    for (int e = 0; e < E; e++) {
      v[e].x *= rinv; v[e].y *= rinv;
      vg[e * C::T] = v[e];
      sts2(sm.u + sm.vec + th.own + (unsigned)e * C::ESTR, v[e]);
    }
    if (l0) sts1(ggrs, res);
    if (res <= ttol) { converged = true; break; }
    csync<NW>();
    int it = 0;
    while (!converged && it < P.mcap && its < P.maxiter) {
      /* w = A v_it */
      double2 w[E];
      {
        double2 mv[E];
        apply<Q, LIND, E, TR>(th, co, sm.u + (unsigned)(1 + (it & 1)) * sm.vec, v, mv);
        stats.evals++;
#pragma unroll
        for (int e = 0; e < E; e++) { w[e].x = v[e].x - alpha * mv[e].x; w[e].y = v[e].y - alpha * mv[e].y; }
      }
      /* classical Gram-Schmidt: inner products in batches of four */
      for (int j0 = 0; j0 <= it; j0 += 4) {
        double d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = (j0 + u <= it) ? j0 + u : it;
          double2 vj[E];
#pragma unroll
          for (int e = 0; e < E; e++) vj[e] = vg[(size_t)j * C::DIM + e * C::T];
          d[u] = dot_own<E>(w, vj);
        }
        allsum<4, NW>(d, sm, th.lane, th.warp);
        if (l0) {
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (j0 + u <= it) sts1(ghc + 8u * (unsigned)(j0 + u), d[u]);
        }
      }
      __syncwarp();
      double hprev = 0.0;
      for (int j = 0; j <= it; j++) {
        const double h = lds1(ghc + 8u * (unsigned)j);
#pragma unroll
        for (int e = 0; e < E; e++) {
          const double2 vj = vg[(size_t)j * C::DIM + e * C::T];
          w[e].x -= h * vj.x;
          w[e].y -= h * vj.y;
        }
        if (j == 0) hprev = h;
        else {
          const double cj = lds1(gcc + 8u * (unsigned)(j - 1)), sj = lds1(gss + 8u * (unsigned)(j - 1));
          const double t = hprev;
          const double hj = cj * t + sj * h;
          hprev = -sj * t + cj * h;
          if (l0) hh[it * ldh + (j - 1)] = hj;
        }
      }
      double n2[1] = {dot_own<E>(w, w)};
      allsum<1, NW>(n2, sm, th.lane, th.warp);
      const double tt2 = n2[0];
      double tt = 0.0;
      if (tt2 > 0.0) {
        const double tinv = fast_rsqrt(tt2);
        tt = tt2 * tinv;
#pragma unroll
        for (int e = 0; e < E; e++) { w[e].x *= tinv; w[e].y *= tinv; }
      }
#pragma unroll
      for (int e = 0; e < E; e++) {
        vg[(size_t)(it + 1) * C::DIM + e * C::T] = w[e];
        sts2(sm.u + (unsigned)(1 + ((it + 1) & 1)) * sm.vec + th.own + (unsigned)e * C::ESTR, w[e]);
        v[e] = w[e];
      }
      /* new rotation and residual estimate (KSPGMRESUpdateHessenberg), all lanes redundantly */
      int flag = 0;
      {
        const double hit = hprev, hit1 = tt;
        const double t2 = hit * hit + hit1 * hit1;
        if (t2 == 0.0) flag = 2;
        else {
          const double rt = fast_rsqrt(t2);
          const double ci = hit * rt, si = hit1 * rt;
          const double g = lds1(ggrs + 8u * (unsigned)it);
          const double gnext = -si * g;
          __syncwarp();     /* every lane has read grs[it] before lane 0 overwrites it */
          if (l0) {
            sts1(gcc + 8u * (unsigned)it, ci);
            sts1(gss + 8u * (unsigned)it, si);
            sts1(ggrs + 8u * (unsigned)(it + 1), gnext);
            sts1(ggrs + 8u * (unsigned)it, ci * g);
            sts1(grd + 8u * (unsigned)it, rt);
          }
          flag = (fabs(gnext) <= ttol) ? 1 : 0;
        }
      }
      csync<NW>();   /* the mirrored v_{it+1} and this warp's recurrence arrays visible */
      it++;
      its++;
      if (flag != 0) converged = true;
    }
    /* KSPGMRESBuildSoln: back substitution (all lanes redundantly), then the combination */
    for (int kk = it - 1; kk >= 0; kk--) {
      double tsum = lds1(ggrs + 8u * (unsigned)kk);
      for (int j = kk + 1; j < it; j++) tsum -= hh[j * ldh + kk] * lds1(gnrs + 8u * (unsigned)j);
      __syncwarp();
      if (l0) sts1(gnrs + 8u * (unsigned)kk, tsum * lds1(grd + 8u * (unsigned)kk));
      __syncwarp();
    }
    for (int j = 0; j < it; j++) {
      const double cj = lds1(gnrs + 8u * (unsigned)j);
#pragma unroll
      for (int e = 0; e < E; e++) {
        const double2 vj = vg[(size_t)j * C::DIM + e * C::T];
        y[e].x += cj * vj.x;
        y[e].y += cj * vj.y;
      }
    }
  }
  stats.iters += its;
}

/* out-of-line wrapper of the fallback: called on copies, so that the hot path's per-thread constants stay in registers */
template <int Q, bool LIND, int E, bool TR>
__device__ __noinline__ void gmres_g_call(Th<Q, LIND, E> th, Co<Q, LIND, E> co, SolvePar P, Sm<K<Q, LIND, E>::NW>* sm, double2* y_io, double alpha, SolveStats* stats) {
  using C = K<Q, LIND, E>;
  double2 rhs[E], y[E];
#pragma unroll
  for (int e = 0; e < E; e++) rhs[e] = lds2(sm->u + th.own + (unsigned)e * C::ESTR);      /* b = U[0], own elements */
  gmres_g<Q, LIND, E, TR>(th, co, P, *sm, rhs, alpha, y, *stats);
#pragma unroll
  for (int e = 0; e < E; e++) y_io[e] = y[e];
}

/* (I - alpha M) y = b (TR: M^T) with the solver GM (GMRES) or Neumann; optionally my = alpha M y.
 * PRE: b = M v0 is formed inside (forward step; v0 = x_n is visible in slot src0); otherwise b = v0 and the caller has stored
 * it in U[0] (not yet synchronised).  spred: length prediction carried from solve to solve (>= 1), or -(number of solves that
 * still skip the fast path). */
template <int Q, bool LIND, int E, bool TR, bool WANT_MY, bool GM, bool PRE>
__device__ __forceinline__ void linsolve(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, Sm<K<Q, LIND, E>::NW>& sm,
                                         const double2 (&v0)[E], unsigned src0, double alpha, double2 (&y)[E], double2 (&my)[E], SolveStats& stats, int& spred) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  const unsigned src = PRE ? src0 : sm.u;
  const double scale0 = PRE ? 1.0 : alpha;
  constexpr int j0 = PRE ? 0 : 1;
  if constexpr (!GM) { neumann_mono<Q, LIND, E, TR, WANT_MY>(th, co, P, sm, v0, src, scale0, j0, !PRE, alpha, y, my, stats, spred); return; }
  bool ok = false;
  if (spred >= 1) {
    ok = gmres_mono<Q, LIND, E, TR, WANT_MY>(th, co, P, sm, v0, src, scale0, j0, !PRE, alpha, y, my, stats, spred);
    if (!ok) spred = -7;
  } else {
    spred = spred == 0 ? kSBlk : spred + 1;
    if (PRE) {     /* b = M v0 into U[0] */
      double2 un[E];
      apply<Q, LIND, E, TR>(th, co, src0, v0, un);
#pragma unroll
      for (int e = 0; e < E; e++) sts2(sm.u + th.own + (unsigned)e * C::ESTR, un[e]);
    }
  }
  if (!ok) {
    {
      Sm<NW> smc = sm;
      SolveStats stc = {0, 0};
      double2 yy[E];
      gmres_g_call<Q, LIND, E, TR>(th, co, P, &smc, yy, alpha, &stc);
      sm.red = smc.red; sm.red_other = smc.red_other;
      stats.iters += stc.iters; stats.evals += stc.evals;
#pragma unroll
      for (int e = 0; e < E; e++) y[e] = yy[e];
    }
    if (WANT_MY) {
      const unsigned xs = sm.u + 4u * sm.vec;
      csync<NW>();
#pragma unroll
      for (int e = 0; e < E; e++) sts2(xs + th.own + (unsigned)e * C::ESTR, y[e]);
      csync<NW>();
      apply<Q, LIND, E, TR>(th, co, xs, y, my);
#pragma unroll
      for (int e = 0; e < E; e++) { my[e].x *= alpha; my[e].y *= alpha; }
    }
  }
}

template <int Q, bool LIND, int E>
__device__ __forceinline__ void sm_init(Sm<K<Q, LIND, E>::NW>& sm, const QLayout& L, const QArgs& A, int ic, int warp) {
  using C = K<Q, LIND, E>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  sm.base = (unsigned)__cvta_generic_to_shared(smem_raw);
  sm.red = sm.base + L.red;
  sm.red_other = sm.red + (unsigned)(kNRed * C::NW * 8);
  sm.giv = sm.base + L.giv + (unsigned)warp * kGivW;
  sm.xb = sm.base + L.xb;
  sm.u = sm.base + L.u;
  sm.vec = C::VEC;
  for (unsigned o = threadIdx.x * 16u; o < (unsigned)(kSMax + 1) * C::VEC; o += C::T * 16u) sts2(sm.u + o, make_double2(0.0, 0.0));   /* finite leftovers only */
  sm.vg = A.scratch + (size_t)ic * (A.mcap + 1) * C::DIM;
  sm.hhg = A.hh + ((size_t)ic * C::NW + warp) * (size_t)(A.mcap * A.mcap);
}

/* ---- forward sweep: TimeStepper::solveODE + ImplMidpoint::evolveFWD ---------------------------------------- */
template <int Q, bool LIND, int E, bool GM>
__global__ void __launch_bounds__(K<Q, LIND, E>::T) qfwd_kernel(const __grid_constant__ DevSys S, const __grid_constant__ QArgs A) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  const bool store_stage = A.stage_traj != nullptr;
  const QLayout L = qlayout(NW, C::VEC, false, false, store_stage, A.rec_fwd_bytes, 0);
  Th<Q, LIND, E> th;
  th_init(th, S);
  const int ic = blockIdx.x;
  Sm<NW> sm;
  sm_init<Q, LIND, E>(sm, L, A, ic, th.warp);
  const int ntime = S.ntime;
  const SolvePar P = {S.maxiter, A.mcap, S.abstol, S.reltol};
  SolveStats stats = {0, 0};
  const unsigned bars = sm.base + L.bars, recring = sm.base + L.rec;
  const char* recg = reinterpret_cast<const char*>(A.rec);
  const size_t recstride_b = (size_t)A.recstride * 8;

  if (th.tid == 0) {
    for (int s = 0; s < kNR; s++) mbar_init(bars + 8u * s, 1);
    fence_proxy_async();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < kNR && s < ntime; s++) {
      mbar_expect_tx(bars + 8u * s, A.rec_fwd_bytes);
      bulk_g2s(recring + (unsigned)s * A.rec_fwd_bytes, recg + (size_t)s * recstride_b, A.rec_fwd_bytes, bars + 8u * s);
    }
  }
  double2 x[E];
#pragma unroll
  for (int e = 0; e < E; e++) x[e] = A.state[(size_t)ic * C::DIM + e * C::T + th.tid];
  double2* trajg = A.traj ? A.traj + (size_t)ic * (ntime + 1) * C::DIM : nullptr;
  double2* stageg = store_stage ? A.stage_traj + (size_t)ic * ntime * C::DIM : nullptr;
  const bool bulk = trajg != nullptr || stageg != nullptr;
  const unsigned sb0 = sm.base + L.sb, xb0 = sm.xb;
#pragma unroll
  for (int e = 0; e < E; e++) sts2(xb0 + th.own + (unsigned)e * C::ESTR, x[e]);
  if (bulk) fence_proxy_async();
  __syncthreads();   /* x_0 in the exchange buffer; mbarrier initialisation visible to every waiting thread */
  if (th.tid == 0 && trajg) bulk_s2g(trajg, xb0, C::VEC);

  Co<Q, LIND, E> co;
  int spred = (S.flags & QB_FLAG_ARNOLDI) ? -(1 << 30) : 3;
  QB_CLK_INIT();
  for (int n = 0; n < ntime; n++) {
    const int slot = n % kNR;
    mbar_wait(bars + 8u * slot, (unsigned)((n / kNR) & 1));
    const double dtn = co_load(co, th, recring + (unsigned)slot * A.rec_fwd_bytes);
    QB_CLK(0);      /* record wait + load */
    const unsigned xb = xb0 + (unsigned)(n & 1) * C::VEC, xbn = xb0 + (unsigned)((n + 1) & 1) * C::VEC;
    double2 k[E], unused[E];
    /* k = (I - dtn/2 M)^{-1} M x_n; the right-hand side M x_n is the first vector the solver builds */
    linsolve<Q, LIND, E, false, false, GM, true>(th, co, P, sm, x, xb, dtn / 2.0, k, unused, stats, spred);
    stats.evals++;
    /* the barriers inside linsolve also tell thread 0 that every thread has consumed this step's record */
    if (th.tid == 0 && n + kNR < ntime) {
      mbar_expect_tx(bars + 8u * slot, A.rec_fwd_bytes);
      bulk_g2s(recring + (unsigned)slot * A.rec_fwd_bytes, recg + (size_t)(n + kNR) * recstride_b, A.rec_fwd_bytes, bars + 8u * slot);
    }
    const unsigned sb = sb0 + (unsigned)(n & 1) * C::VEC;
    if (stageg) {
      /* the stage ImplMidpoint::evolveBWD re-solves for (src/timestepper.cpp:674-680), kept for the adjoint sweep */
#pragma unroll
      for (int e = 0; e < E; e++) sts2(sb + th.own + (unsigned)e * C::ESTR, make_double2(x[e].x + dtn / 2.0 * k[e].x, x[e].y + dtn / 2.0 * k[e].y));
    }
#pragma unroll
    for (int e = 0; e < E; e++) { x[e].x += dtn * k[e].x; x[e].y += dtn * k[e].y; sts2(xbn + th.own + (unsigned)e * C::ESTR, x[e]); }
    if (bulk) fence_proxy_async();
    QB_CLK(8);      /* record refill, stage / state update and stores to shared memory */
    csync<NW>();     /* x_{n+1} (and the stage) complete in shared memory: operand of the next step and source of the bulk stores */
    QB_CLK(9);      /* end-of-step barrier */
    if (th.tid == 0 && bulk) {
      if (stageg) bulk_s2g(stageg + (size_t)n * C::DIM, sb, C::VEC);
      if (trajg) bulk_s2g(trajg + (size_t)(n + 1) * C::DIM, xbn, C::VEC);
      /* the buffers written at the end of step n+1 were the sources of the groups issued one step ago */
      if (stageg && trajg) bulk_wait_read2(); else bulk_wait_read1();
    }
    QB_CLK(10);     /* bulk store issue + wait for the previous group's reads */
  }
  QB_CLK_PRINT("qfwd", ntime);
#pragma unroll
  for (int e = 0; e < E; e++) A.state_out[(size_t)ic * C::DIM + e * C::T + th.tid] = x[e];
  if (th.tid == 0) {
    bulk_wait_all();
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* ---- adjoint sweep: TimeStepper::solveAdjointODE + ImplMidpoint::evolveBWD + dRHSdp + evalControl_diff ------ */
/* STORED: the primal ring carries the stages kept by the forward sweep (QB_FLAG_STORE_STAGE); otherwise it carries the
 * states and the stage is re-solved (QB_FLAG_RECOMPUTE_STAGE) or averaged.  Lindblad only (the state trajectory or the
 * stages are always stored, src/timestepper.cpp:230). */
template <int Q, bool LIND, int E, bool STORED, bool GM>
__global__ void __launch_bounds__(K<Q, LIND, E>::T) qadj_kernel(const __grid_constant__ DevSys S, const __grid_constant__ QArgs A) {
  using C = K<Q, LIND, E>;
  static_assert(LIND, "the qubit-system kernels are instantiated for the Lindblad solver");
  constexpr int NW = C::NW;
  constexpr bool stored = STORED;
  const QLayout L = qlayout(NW, C::VEC, true, true, false, A.rec_bytes, S.nd);
  Th<Q, LIND, E> th;
  th_init(th, S);
  const int ic = blockIdx.x;
  Sm<NW> sm;
  sm_init<Q, LIND, E>(sm, L, A, ic, th.warp);
  const int ntime = S.ntime, nd = S.nd, nupd = S.nupd;
  const SolvePar P = {S.maxiter, A.mcap, S.abstol, S.reltol};
  const bool recompute = !stored && (S.flags & (QB_FLAG_RECOMPUTE_STAGE | QB_FLAG_STORE_STAGE)) != 0;
  SolveStats stats = {0, 0};
  const unsigned rbars = sm.base + L.bars, tbars = sm.base + L.bars + 8u * kNR;
  const unsigned recring = sm.base + L.rec, trajring = sm.base + L.traj, sgrad = sm.base + L.grad, supd = sm.base + L.upd;
  const char* recg = reinterpret_cast<const char*>(A.rec);
  const size_t recstride_b = (size_t)A.recstride * 8;
  /* primal ring: load q reads prim[-q]; states: q = 0 is x_ntime and step n = ntime-r uses q = r+1 (x_{n-1});
   * stored stages: step n = ntime-r uses q = r (stage of step n-1 -> n) */
  const double2* prim = stored ? A.stage_traj + ((size_t)ic * ntime + (ntime - 1)) * C::DIM
                               : A.traj + ((size_t)ic * (ntime + 1) + ntime) * C::DIM;
  const int qlast = stored ? ntime - 1 : ntime, qoff = stored ? 0 : 1;

  /* r-th record load = record ntime-1-r (used by step n = ntime-r) */
  if (th.tid == 0) {
    for (int s = 0; s < kNR + kNT; s++) mbar_init(rbars + 8u * s, 1);
    fence_proxy_async();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int r = 0; r < kNR && r < ntime; r++) {
      mbar_expect_tx(rbars + 8u * r, A.rec_bytes);
      bulk_g2s(recring + (unsigned)r * A.rec_bytes, recg + (size_t)(ntime - 1 - r) * recstride_b, A.rec_bytes, rbars + 8u * r);
    }
    for (int q = 0; q < kNT && q <= qlast; q++) {
      mbar_expect_tx(tbars + 8u * q, C::VEC);
      bulk_g2s(trajring + (unsigned)q * C::VEC, prim - (size_t)q * C::DIM, C::VEC, tbars + 8u * q);
    }
  }
  for (int i = th.tid; i < nd; i += C::T) sts1(sgrad + 8u * i, 0.0);
  for (int i = th.tid; i < nupd * 4; i += C::T) asm volatile("st.shared.s32 [%0], %1;" ::"r"(supd + 4u * i), "r"(A.upd[i]) : "memory");
  double2 xadj[E], xcur[E];
#pragma unroll
  for (int e = 0; e < E; e++) { xadj[e] = A.state[(size_t)ic * C::DIM + e * C::T + th.tid]; xcur[e] = make_double2(0.0, 0.0); }
  __syncthreads();
  if constexpr (!stored) {
    mbar_wait(tbars, 0u);     /* q = 0: x_ntime */
#pragma unroll
    for (int e = 0; e < E; e++) xcur[e] = lds2(trajring + th.own + (unsigned)e * C::ESTR);
    __syncthreads();
    if (th.tid == 0 && kNT <= qlast) {   /* slot 0 is free again: load q = kNT */
      mbar_expect_tx(tbars, C::VEC);
      bulk_g2s(trajring, prim - (size_t)kNT * C::DIM, C::VEC, tbars);
    }
  }

  Co<Q, LIND, E> co;
  int spred = (S.flags & QB_FLAG_ARNOLDI) ? -(1 << 30) : 3, spred_f = spred;
  for (int n = ntime; n > 0; n--) {
    const int r = ntime - n, rslot = r % kNR;
    mbar_wait(rbars + 8u * rslot, (unsigned)((r / kNR) & 1));
    const unsigned rec = recring + (unsigned)rslot * A.rec_bytes;
    const double dtn = co_load(co, th, rec);
    const int q = r + qoff, tslot = q % kNT;
    if (A.pen.on) {   /* derivative of the penalty integral at x_n, src/timestepper.cpp:227,302-339 */
      const double jb = A.pen.jbar[ic], cw = A.pen.wdt[n];
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int it = e * C::T + th.tid;
        const double2 xn = stored ? A.traj[((size_t)ic * (ntime + 1) + n) * C::DIM + it] : xcur[e];
        const double2 pa = A.pen.a[(size_t)ic * C::DIM + it], pd = A.pen.d[it];
        double gx = 0.0, gy = 0.0;
        if (A.pen.g) { const double2 pg = A.pen.g[it]; const double gwn = A.pen.gw[n]; gx = 2.0 * gwn * pg.x; gy = 2.0 * gwn * pg.y; }
        xadj[e].x += jb * (cw * (pa.x + 2.0 * pd.x * xn.x) + gx * xn.x);
        xadj[e].y += jb * (cw * (pa.y + 2.0 * pd.y * xn.y) + gy * xn.y);
      }
    }
    /* adjoint stage: (I - dtn/2 M)^T ksol = xadj, kbar = dtn ksol; and M^T kbar = 2 (dtn/2) M^T ksol comes with the solve */
    double2 kbar[E], mk[E];
#pragma unroll
    for (int e = 0; e < E; e++) sts2(sm.u + th.own + (unsigned)e * C::ESTR, xadj[e]);
    linsolve<Q, LIND, E, true, true, GM, false>(th, co, P, sm, xadj, 0u, dtn / 2.0, kbar, mk, stats, spred);
    stats.evals++;      /* the M^T kbar application of evolveBWD (src/timestepper.cpp:691), a by-product here */
    if (th.tid == 0 && r >= 1 && r - 1 + kNR < ntime) {
      /* the record of the previous step: its last readers (the gradient updaters) have passed the barrier inside linsolve */
      const int ps = (r - 1) % kNR;
      mbar_expect_tx(rbars + 8u * ps, A.rec_bytes);
      bulk_g2s(recring + (unsigned)ps * A.rec_bytes, recg + (size_t)(ntime - 1 - (r - 1 + kNR)) * recstride_b, A.rec_bytes, rbars + 8u * ps);
    }
#pragma unroll
    for (int e = 0; e < E; e++) {
      kbar[e].x *= dtn; kbar[e].y *= dtn;
      xadj[e].x = fma(2.0, mk[e].x, xadj[e].x); xadj[e].y = fma(2.0, mk[e].y, xadj[e].y);
    }
    /* primal stage x_{n-1} + dtn/2 k */
    mbar_wait(tbars + 8u * tslot, (unsigned)((q / kNT) & 1));
    double2 stage[E], xprev[E];
#pragma unroll
    for (int e = 0; e < E; e++) xprev[e] = lds2(trajring + (unsigned)tslot * C::VEC + th.own + (unsigned)e * C::ESTR);
    unsigned xs = trajring + (unsigned)tslot * C::VEC;
    if constexpr (stored) {
#pragma unroll
      for (int e = 0; e < E; e++) stage[e] = xprev[e];
    } else {
      if (recompute) {
        double2 k[E], unused[E];
        csync<NW>();      /* the adjoint solve's last reads of the Krylov slots are done before this solve overwrites them */
        linsolve<Q, LIND, E, false, false, GM, true>(th, co, P, sm, xprev, xs, dtn / 2.0, k, unused, stats, spred_f);
        stats.evals++;
#pragma unroll
        for (int e = 0; e < E; e++) stage[e] = make_double2(xprev[e].x + dtn / 2.0 * k[e].x, xprev[e].y + dtn / 2.0 * k[e].y);
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) stage[e] = make_double2(0.5 * (xprev[e].x + xcur[e].x), 0.5 * (xprev[e].y + xcur[e].y));
      }
      xs = sm.xb;
#pragma unroll
      for (int e = 0; e < E; e++) sts2(xs + th.own + (unsigned)e * C::ESTR, stage[e]);
      csync<NW>();
    }
    /* dRHSdp coefficients: coeff_p[k] = kbar^T (dM/dp_k) stage, coeff_q[k] likewise (src/mastereq.cpp:970-1276) */
    {
      double cpq[2 * Q];
#pragma unroll
      for (int k = 0; k < Q; k++) {
        double cp = 0.0, cq = 0.0;
#pragma unroll
        for (int e = 0; e < E; e++) {
#pragma unroll
          for (int side = 0; side < (LIND ? 2 : 1); side++) {
            const int m = side == 0 ? Q - 1 - k : 2 * Q - 1 - k;
            double2 nb;
            bool neg_b;
            if (m < C::LT) { nb = lds2(xs + th.nboff[m] + (unsigned)e * C::ESTR); neg_b = ((th.tid >> m) & 1) != 0; }
            else { const int lb = m >= C::LT ? m - C::LT : 0; nb = stage[e ^ (1 << lb)]; neg_b = ((e >> lb) & 1) != 0; }
            const double tp = nb.y * kbar[e].x - nb.x * kbar[e].y;
            const double tq = nb.x * kbar[e].x + nb.y * kbar[e].y;
            cp += side == 0 ? tp : -tp;
            cq += neg_b ? -tq : tq;
          }
        }
        cpq[2 * k] = cp;
        cpq[2 * k + 1] = cq;
      }
      stats.evals++;
      allsum<2 * Q, NW>(cpq, sm, th.lane, th.warp);
      /* every thread has read its part of the primal ring slot (and of the exchange buffer): refill the slot */
      if (th.tid == 0 && q + kNT <= qlast) {
        mbar_expect_tx(tbars + 8u * tslot, C::VEC);
        bulk_g2s(trajring + (unsigned)tslot * C::VEC, prim - (size_t)(q + kNT) * C::DIM, C::VEC, tbars + 8u * tslot);
      }
      /* Oscillator::evalControl_diff: one updater per (oscillator, carrier, non-zero basis function) */
      for (int u = th.tid; u < nupd; u += C::T) {
        int uk, uc, uf, uj;
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(uk), "=r"(uc), "=r"(uf), "=r"(uj) : "r"(supd + 16u * u) : "memory");
        int gi_idx, gi_ns;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(gi_idx) : "r"(rec + 8u * A.rec_gidx + 4u * (unsigned)(uk * 4 + uj)) : "memory");
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(gi_ns) : "r"(rec + 8u * A.rec_gidx + 4u * (unsigned)(uk * 4 + 3)) : "memory");
        if (gi_idx >= 0) {
          const double g_bval = lds1(rec + 8u * (unsigned)(A.rec_gval + uk * 3 + uj));
          const double g_cs = lds1(rec + 8u * (unsigned)(A.rec_gcs + 2 * uc)), g_sn = lds1(rec + 8u * (unsigned)(A.rec_gcs + 2 * uc + 1));
          double cp = 0.0, cq = 0.0;
#pragma unroll
          for (int k = 0; k < Q; k++)
            if (k == uk) { cp = cpq[2 * k]; cq = cpq[2 * k + 1]; }
          const double pbar = 1.0 * cp, qbar = 1.0 * cq;
          const double B1bar = g_sn * qbar + g_cs * pbar;
          const double B2bar = g_cs * qbar - g_sn * pbar;
          const unsigned gaddr = sgrad + 8u * (unsigned)(S.osc_off[uk] + gi_idx + uf * gi_ns * 2);
          sts1(gaddr, lds1(gaddr) + g_bval * B1bar);
          sts1(gaddr + 8u * gi_ns, lds1(gaddr + 8u * gi_ns) + g_bval * B2bar);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; e++) xcur[e] = xprev[e];
  }
  __syncthreads();
  for (int i = th.tid; i < nd; i += C::T) A.grad_ic[(size_t)ic * nd + i] = lds1(sgrad + 8u * i);
#pragma unroll
  for (int e = 0; e < E; e++) A.state_out[(size_t)ic * C::DIM + e * C::T + th.tid] = xadj[e];
  if (th.tid == 0) {
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* y = M(t) x or M(t)^T x for a batch of vectors (unit-test hook); rec: one record in global memory */
template <int Q, bool LIND, int E>
__global__ void __launch_bounds__(K<Q, LIND, E>::T) qapply_kernel(const __grid_constant__ DevSys S, const double* __restrict__ rowg,
                                                                  const double2* __restrict__ xg, double2* __restrict__ yg, int transpose) {
  using C = K<Q, LIND, E>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem_raw);
  Th<Q, LIND, E> th;
  th_init(th, S);
  if (th.tid < 1 + 2 * Q) sts1(base + C::VEC + 8u * th.tid, rowg[th.tid]);
  double2 x[E], y[E];
#pragma unroll
  for (int e = 0; e < E; e++) { x[e] = xg[(size_t)blockIdx.x * C::DIM + e * C::T + th.tid]; sts2(base + th.own + (unsigned)e * C::ESTR, x[e]); }
  __syncthreads();
  Co<Q, LIND, E> co;
  co_load(co, th, base + C::VEC);
  if (transpose) apply<Q, LIND, E, true>(th, co, base, x, y);
  else apply<Q, LIND, E, false>(th, co, base, x, y);
#pragma unroll
  for (int e = 0; e < E; e++) yg[(size_t)blockIdx.x * C::DIM + e * C::T + th.tid] = y[e];
}

}  // namespace qq
