/*
 * qb_qubit.cuh -- sweep kernels specialised for systems of two-level oscillators (qubits) without Jkl coupling:
 * BASELINE configs C3 (N=8, E=2) and C4 (N=16, E=4; E=2 when a sweep has no more initial conditions than SMs).
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
 *   - T = dim/E threads per initial condition (T=32: a single warp, no block barrier anywhere in the sweep;
 *     T=64: two warps, one barrier per reduction; measured on C4: E=4/T=64 is fastest, 2 CTAs per SM).  Reductions:
 *     in-thread over E, butterfly over the warp, warp partials in fixed order.
 *   - The trajectory leaves / enters the SM through the TMA bulk-copy engine (cp.async.bulk + mbarrier): the forward
 *     sweep's exchange buffer IS the source of the bulk store, the adjoint sweep prefetches its primal input (x_{n-1} or
 *     the stored stage) one step ahead into a two-slot shared-memory ring.  The per-step control record (dt, p_k, q_k, carrier phases, B-spline
 *     values/indices) is streamed the same way, so a step issues no global load at all.
 *   - the small GMRES recurrences (Givens rotations, Hessenberg, back substitution) are computed redundantly by all
 *     lanes from warp-private shared arrays (broadcast loads): no divergence, no flag broadcast.
 */
#pragma once
#include "qb_kernels.cuh"
#include "qb_qubit_args.h"

namespace qq {

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
  static constexpr int NRED = (2 * Q > 4 ? 2 * Q : 4);
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
 * Same accumulation order as qbk::apply_rhs (oscillators in order, bra then ket; zero-coefficient terms dropped). */
template <int Q, bool LIND, int E, bool TR>
__device__ __forceinline__ void apply(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, unsigned vs, const double2 (&x)[E], double2 (&y)[E]) {
  using C = K<Q, LIND, E>;
#pragma unroll
  for (int e = 0; e < E; e++) {
    const unsigned eo = (unsigned)e * C::ESTR;
    double hre = th.dA[e] * x[e].y, him = -th.dA[e] * x[e].x, hre2 = 0.0, him2 = 0.0;
#pragma unroll
    for (int k = 0; k < Q; k++) {
#pragma unroll
      for (int side = 0; side < (LIND ? 2 : 1); side++) {
        const int m = side == 0 ? Q - 1 - k : 2 * Q - 1 - k;
        double2 nb;
        double b;
        if (m < C::LT) { nb = lds2(vs + th.nboff[m] + eo); b = co.qr[m]; }
        else { const int lb = m >= C::LT ? m - C::LT : 0; nb = x[e ^ (1 << lb)]; b = ((e >> lb) & 1) ? -co.q[k] : co.q[k]; }
        const double a = side == 0 ? co.p[k] : -co.p[k];
        hre = fma(a, nb.y, hre);
        hre2 = fma(b, nb.x, hre2);
        him = fma(-a, nb.x, him);
        him2 = fma(b, nb.y, him2);
      }
    }
    hre += hre2; him += him2;
    double dre = 0.0, dim = 0.0;
    if (LIND) {
      dre = th.dB[e] * x[e].x; dim = th.dB[e] * x[e].y;
#pragma unroll
      for (int k = 0; k < Q; k++) {
        const int rb = Q - 1 - k, cb = 2 * Q - 1 - k;
        /* forward: rho_{i+1,i'+1} feeds rho_{i,i'} (both bits 0 here); transpose: the opposite shift (both bits 1) */
        bool ok = true;
        int emask = 0;
        if (rb >= C::LT) { const int lb = rb >= C::LT ? rb - C::LT : 0; if (((e >> lb) & 1) != (TR ? 1 : 0)) ok = false; emask |= 1 << lb; }
        if (cb >= C::LT) { const int lb = cb >= C::LT ? cb - C::LT : 0; if (((e >> lb) & 1) != (TR ? 1 : 0)) ok = false; emask |= 1 << lb; }
        if (ok) {
          const double l1 = TR ? th.gt[k] : th.gf[k];
          double2 nl;
          if (rb >= C::LT && cb >= C::LT) nl = x[e ^ emask];
          else nl = lds2(vs + th.l1off[k] + (unsigned)(e ^ emask) * C::ESTR);
          dre = fma(l1, nl.x, dre);
          dim = fma(l1, nl.y, dim);
        }
      }
    }
    y[e] = TR ? make_double2(dre - hre, dim - him) : make_double2(dre + hre, dim + him);
  }
}

/* ---- CTA-wide sums ------------------------------------------------------------------------------------------ */
template <int NW>
struct Sm {
  unsigned base;      /* shared address of the dynamic shared memory */
  unsigned red;       /* reduction half the NEXT reduction uses */
  unsigned red_other; /* the other half (the two are swapped after every use) */
  unsigned giv;       /* this warp's Givens scratch */
  unsigned hh;        /* this warp's Hessenberg (ldh = mcap doubles) */
  unsigned xb, v, vec;
};

template <int NV, int NW>
__device__ __forceinline__ void allsum(double (&v)[NV], Sm<NW>& sm, int lane, int warp) {
  qbk::warp_sum_n<NV>(v);
  if constexpr (NW > 1) {
    const unsigned r = sm.red;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NV; i++) sts1(r + (unsigned)(i * NW + warp) * 8u, v[i]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = qbk::partial_sum<NW>(r + (unsigned)(i * NW) * 8u);
    sm.red = sm.red_other;
    sm.red_other = r;
  }
}

template <int E>
__device__ __forceinline__ double dot_own(const double2 (&a)[E], const double2 (&b)[E]) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int e = 0; e < E; e++) { s0 = fma(a[e].x, b[e].x, s0); s1 = fma(a[e].y, b[e].y, s1); }
  return s0 + s1;
}

/* ---- GMRES: same algorithm and stopping rule as qbk::gmres_solve (KSPGMRES defaults, src/timestepper.cpp:541-550) ---- */
template <int Q, bool LIND, int E, bool TR>
__device__ __forceinline__ void gmres(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, Sm<K<Q, LIND, E>::NW>& sm,
                                      const double2 (&rhs)[E], double alpha, double2 (&y)[E], SolveStats& stats) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  const unsigned gcc = sm.giv, gss = sm.giv + 256, ggrs = sm.giv + 512, grd = sm.giv + 768, gnrs = sm.giv + 1024, ghc = sm.giv + 1280;
  const unsigned ldh = (unsigned)P.mcap * 8u;
  const bool l0 = th.lane == 0;
#pragma unroll
  for (int e = 0; e < E; e++) y[e] = make_double2(0.0, 0.0);
  int its = 0;
  bool converged = false, first = true;
  double ttol = 0.0;
  while (!converged && its < P.maxiter) {
    double2 v[E];
    if (first) {
#pragma unroll
      for (int e = 0; e < E; e++) v[e] = rhs[e];
    } else {
      /* restart: r = b - A y */
#pragma unroll
      for (int e = 0; e < E; e++) sts2(sm.xb + th.own + (unsigned)e * C::ESTR, y[e]);
      csync<NW>();
      double2 my[E];
      apply<Q, LIND, E, TR>(th, co, sm.xb, y, my);
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
    for (int e = 0; e < E; e++) { v[e].x *= rinv; v[e].y *= rinv; sts2(sm.v + th.own + (unsigned)e * C::ESTR, v[e]); }
    if (l0) sts1(ggrs, res);
    if (res <= ttol) { converged = true; break; }
    csync<NW>();
    int it = 0;
    while (!converged && it < P.mcap && its < P.maxiter) {
      /* w = A v_it */
      double2 w[E];
      {
        double2 mv[E];
        apply<Q, LIND, E, TR>(th, co, sm.v + (unsigned)it * sm.vec, v, mv);
        stats.evals++;
#pragma unroll
        for (int e = 0; e < E; e++) { w[e].x = v[e].x - alpha * mv[e].x; w[e].y = v[e].y - alpha * mv[e].y; }
      }
      /* classical Gram-Schmidt + the previous Givens rotations applied to the new Hessenberg column on the fly */
      double hprev = 0.0;
      if (it < 4) {
        double d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          d[u] = 0.0;
          if (u <= it) {
            double2 vj[E];
#pragma unroll
            for (int e = 0; e < E; e++) vj[e] = lds2(sm.v + (unsigned)u * sm.vec + th.own + (unsigned)e * C::ESTR);
            d[u] = dot_own<E>(w, vj);
          }
        }
        if (it == 0) { double t1[1] = {d[0]}; allsum<1, NW>(t1, sm, th.lane, th.warp); d[0] = t1[0]; }
        else if (it == 1) { double t2[2] = {d[0], d[1]}; allsum<2, NW>(t2, sm, th.lane, th.warp); d[0] = t2[0]; d[1] = t2[1]; }
        else if (it == 2) { double t3[3] = {d[0], d[1], d[2]}; allsum<3, NW>(t3, sm, th.lane, th.warp); d[0] = t3[0]; d[1] = t3[1]; d[2] = t3[2]; }
        else allsum<4, NW>(d, sm, th.lane, th.warp);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (u <= it) {
            const double h = d[u];
#pragma unroll
            for (int e = 0; e < E; e++) {
              const double2 vj = lds2(sm.v + (unsigned)u * sm.vec + th.own + (unsigned)e * C::ESTR);
              w[e].x -= h * vj.x;
              w[e].y -= h * vj.y;
            }
            if (u == 0) hprev = h;
            else {
              const double cj = lds1(gcc + 8u * (u - 1)), sj = lds1(gss + 8u * (u - 1));
              const double t = hprev;
              const double hj = cj * t + sj * h;
              hprev = -sj * t + cj * h;
              if (l0) sts1(sm.hh + (unsigned)it * ldh + 8u * (u - 1), hj);
            }
          }
        }
      } else {
        /* generic column: inner products in batches of four through the warp-private column buffer */
        for (int j0 = 0; j0 <= it; j0 += 4) {
          double d[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int j = (j0 + u <= it) ? j0 + u : it;
            double2 vj[E];
#pragma unroll
            for (int e = 0; e < E; e++) vj[e] = lds2(sm.v + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR);
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
        for (int j = 0; j <= it; j++) {
          const double h = lds1(ghc + 8u * (unsigned)j);
#pragma unroll
          for (int e = 0; e < E; e++) {
            const double2 vj = lds2(sm.v + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR);
            w[e].x -= h * vj.x;
            w[e].y -= h * vj.y;
          }
          if (j == 0) hprev = h;
          else {
            const double cj = lds1(gcc + 8u * (unsigned)(j - 1)), sj = lds1(gss + 8u * (unsigned)(j - 1));
            const double t = hprev;
            const double hj = cj * t + sj * h;
            hprev = -sj * t + cj * h;
            if (l0) sts1(sm.hh + (unsigned)it * ldh + 8u * (unsigned)(j - 1), hj);
          }
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
      if (it + 1 < P.mcap) {
#pragma unroll
        for (int e = 0; e < E; e++) sts2(sm.v + (unsigned)(it + 1) * sm.vec + th.own + (unsigned)e * C::ESTR, w[e]);
      }
#pragma unroll
      for (int e = 0; e < E; e++) v[e] = w[e];
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
      csync<NW>();   /* V[it+1] and this warp's recurrence arrays visible */
      it++;
      its++;
      if (flag != 0) converged = true;
    }
    /* KSPGMRESBuildSoln: back substitution (all lanes redundantly), then the combination */
    if (it > 0) {
      if (it <= 4) {
        double nr[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int kk = 3; kk >= 0; kk--) {
          if (kk < it) {
            double tsum = lds1(ggrs + 8u * kk);
#pragma unroll
            for (int j = kk + 1; j < 4; j++)
              if (j < it) tsum -= lds1(sm.hh + (unsigned)j * ldh + 8u * kk) * nr[j];
            nr[kk] = tsum * lds1(grd + 8u * kk);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (j < it) {
#pragma unroll
            for (int e = 0; e < E; e++) {
              const double2 vj = lds2(sm.v + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR);
              y[e].x += nr[j] * vj.x;
              y[e].y += nr[j] * vj.y;
            }
          }
        }
      } else {
        for (int kk = it - 1; kk >= 0; kk--) {
          double tsum = lds1(ggrs + 8u * (unsigned)kk);
          for (int j = kk + 1; j < it; j++) tsum -= lds1(sm.hh + (unsigned)j * ldh + 8u * (unsigned)kk) * lds1(gnrs + 8u * (unsigned)j);
          if (l0) sts1(gnrs + 8u * (unsigned)kk, tsum * lds1(grd + 8u * (unsigned)kk));
          __syncwarp();
        }
        for (int j = 0; j < it; j++) {
          const double cj = lds1(gnrs + 8u * (unsigned)j);
#pragma unroll
          for (int e = 0; e < E; e++) {
            const double2 vj = lds2(sm.v + (unsigned)j * sm.vec + th.own + (unsigned)e * C::ESTR);
            y[e].x += cj * vj.x;
            y[e].y += cj * vj.y;
          }
        }
      }
    }
  }
  stats.iters += its;
}

/* ImplMidpoint::NeumannSolve, src/timestepper.cpp:697-727 */
template <int Q, bool LIND, int E, bool TR>
__device__ __forceinline__ void neumann(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, Sm<K<Q, LIND, E>::NW>& sm,
                                        const double2 (&rhs)[E], double alpha, double2 (&y)[E], SolveStats& stats) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  unsigned cur = sm.v, nxt = sm.v + sm.vec;
#pragma unroll
  for (int e = 0; e < E; e++) { y[e] = rhs[e]; sts2(cur + th.own + (unsigned)e * C::ESTR, y[e]); }
  csync<NW>();
  double errnorm0 = 0.0;
  int iter;
  for (iter = 0; iter < P.maxiter; iter++) {
    double2 my[E];
    apply<Q, LIND, E, TR>(th, co, cur, y, my);
    stats.evals++;
    double e2 = 0.0, e3 = 0.0;
#pragma unroll
    for (int e = 0; e < E; e++) {
      const double2 yn = make_double2(1.0 * rhs[e].x + alpha * my[e].x, 1.0 * rhs[e].y + alpha * my[e].y);
      const double ex = y[e].x - yn.x, ey = y[e].y - yn.y;
      e2 = fma(ex, ex, e2); e3 = fma(ey, ey, e3);
      y[e] = yn;
      sts2(nxt + th.own + (unsigned)e * C::ESTR, yn);
    }
    double s[1] = {e2 + e3};
    allsum<1, NW>(s, sm, th.lane, th.warp);
    if constexpr (NW == 1) __syncwarp();      /* nxt visible (NW > 1: the barrier inside allsum) */
    const double errnorm = s[0] > 0.0 ? s[0] * fast_rsqrt(s[0]) : 0.0;
    const unsigned t = cur; cur = nxt; nxt = t;
    if (iter == 0) errnorm0 = errnorm;
    if (errnorm < P.abstol) break;
    if (errnorm < P.reltol * errnorm0) break;
  }
  stats.iters += iter;
}

template <int Q, bool LIND, int E, bool TR>
__device__ __forceinline__ void linsolve(const Th<Q, LIND, E>& th, const Co<Q, LIND, E>& co, const SolvePar& P, bool gm, Sm<K<Q, LIND, E>::NW>& sm,
                                         const double2 (&rhs)[E], double alpha, double2 (&y)[E], SolveStats& stats) {
  if (gm) gmres<Q, LIND, E, TR>(th, co, P, sm, rhs, alpha, y, stats);
  else neumann<Q, LIND, E, TR>(th, co, P, sm, rhs, alpha, y, stats);
}

template <int Q, bool LIND, int E>
__device__ __forceinline__ void sm_init(Sm<K<Q, LIND, E>::NW>& sm, const QLayout& L, int warp, int mcap) {
  using C = K<Q, LIND, E>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  sm.base = (unsigned)__cvta_generic_to_shared(smem_raw);
  sm.red = sm.base + L.red;
  sm.red_other = sm.red + (unsigned)(C::NRED * C::NW * 8);
  sm.giv = sm.base + L.giv + (unsigned)warp * kGivW;
  sm.hh = sm.base + L.hh + (unsigned)warp * (unsigned)(mcap * mcap) * 8u;
  sm.xb = sm.base + L.xb;
  sm.v = sm.base + L.v;
  sm.vec = C::VEC;
}

/* ---- forward sweep: TimeStepper::solveODE + ImplMidpoint::evolveFWD ---------------------------------------- */
template <int Q, bool LIND, int E>
__global__ void __launch_bounds__(K<Q, LIND, E>::T) qfwd_kernel(const __grid_constant__ DevSys S, const __grid_constant__ QArgs A) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  const bool store_stage = A.stage_traj != nullptr;
  const QLayout L = qlayout(NW, C::VEC, C::NRED, A.mcap, A.nvec, false, false, store_stage, A.rec_fwd_bytes, 0);
  Th<Q, LIND, E> th;
  th_init(th, S);
  Sm<NW> sm;
  sm_init<Q, LIND, E>(sm, L, th.warp, A.mcap);
  const int ic = blockIdx.x;
  const int ntime = S.ntime;
  const SolvePar P = {S.maxiter, A.mcap, S.abstol, S.reltol};
  const bool gm = S.linsolve == QB_LINSOLVE_GMRES;
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
  double2* trajg = A.store ? A.traj + (size_t)ic * (ntime + 1) * C::DIM : nullptr;
  double2* stageg = store_stage ? A.stage_traj + (size_t)ic * ntime * C::DIM : nullptr;
  const unsigned sb0 = sm.base + L.sb;
  __syncthreads();   /* mbarrier initialisation visible to every waiting thread */

  Co<Q, LIND, E> co;
  const unsigned xb0 = sm.xb;
  for (int n = 0; n < ntime; n++) {
    const int slot = n % kNR;
    mbar_wait(bars + 8u * slot, (unsigned)((n / kNR) & 1));
    const double dtn = co_load(co, th, recring + (unsigned)slot * A.rec_fwd_bytes);
    const unsigned xb = xb0 + (unsigned)(n & 1) * C::VEC;
    sm.xb = xb0 + (unsigned)((n + 1) & 1) * C::VEC;   /* GMRES restart scratch: the buffer no bulk store is reading */
#pragma unroll
    for (int e = 0; e < E; e++) sts2(xb + th.own + (unsigned)e * C::ESTR, x[e]);
    if (trajg) fence_proxy_async();
    csync<NW>();
    if (th.tid == 0) {
      if (trajg) {
        bulk_s2g(trajg + (size_t)n * C::DIM, xb, C::VEC);   /* x_n: shared -> HBM through the bulk-copy engine */
        /* the buffers of step n-1 may be overwritten at step n+1 (groups in flight: x_n [, stage_{n-1}]) */
        if (stageg) bulk_wait_read2(); else bulk_wait_read1();
      }
      if (n + kNR < ntime) {                                /* every thread has consumed this record before the barrier above */
        mbar_expect_tx(bars + 8u * slot, A.rec_fwd_bytes);
        bulk_g2s(recring + (unsigned)slot * A.rec_fwd_bytes, recg + (size_t)(n + kNR) * recstride_b, A.rec_fwd_bytes, bars + 8u * slot);
      }
    }
    double2 rhs[E], k[E];
    apply<Q, LIND, E, false>(th, co, xb, x, rhs);
    stats.evals++;
    linsolve<Q, LIND, E, false>(th, co, P, gm, sm, rhs, dtn / 2.0, k, stats);
    if (stageg) {
      /* the stage ImplMidpoint::evolveBWD re-solves for (src/timestepper.cpp:674-680), kept for the adjoint sweep */
      const unsigned sb = sb0 + (unsigned)(n & 1) * C::VEC;
#pragma unroll
      for (int e = 0; e < E; e++) sts2(sb + th.own + (unsigned)e * C::ESTR, make_double2(x[e].x + dtn / 2.0 * k[e].x, x[e].y + dtn / 2.0 * k[e].y));
      fence_proxy_async();
      csync<NW>();
      if (th.tid == 0) { bulk_s2g(stageg + (size_t)n * C::DIM, sb, C::VEC); bulk_wait_read2(); }
    }
#pragma unroll
    for (int e = 0; e < E; e++) { x[e].x += dtn * k[e].x; x[e].y += dtn * k[e].y; }
  }
#pragma unroll
  for (int e = 0; e < E; e++) {
    if (trajg) trajg[(size_t)ntime * C::DIM + e * C::T + th.tid] = x[e];
    A.state_out[(size_t)ic * C::DIM + e * C::T + th.tid] = x[e];
  }
  if (th.tid == 0) {
    bulk_wait_all();
    atomicAdd(&A.counters[0], (unsigned long long)stats.iters);
    atomicAdd(&A.counters[1], (unsigned long long)stats.evals);
  }
}

/* ---- adjoint sweep: TimeStepper::solveAdjointODE + ImplMidpoint::evolveBWD + dRHSdp + evalControl_diff ------ */
/* STORED: the primal ring carries the stages kept by the forward sweep (QB_FLAG_STORE_STAGE); the re-solving, averaging and
 * Schroedinger paths are compiled out of that instantiation (fewer live registers in the hot loop). */
template <int Q, bool LIND, int E, bool STORED>
__global__ void __launch_bounds__(K<Q, LIND, E>::T) qadj_kernel(const __grid_constant__ DevSys S, const __grid_constant__ QArgs A) {
  using C = K<Q, LIND, E>;
  constexpr int NW = C::NW;
  constexpr bool stored = STORED;       /* primal ring carries stored stages instead of states */
  const bool have_traj = A.traj != nullptr || stored;
  const QLayout L = qlayout(NW, C::VEC, C::NRED, A.mcap, A.nvec, true, have_traj, false, A.rec_bytes, S.nd);
  Th<Q, LIND, E> th;
  th_init(th, S);
  Sm<NW> sm;
  sm_init<Q, LIND, E>(sm, L, th.warp, A.mcap);
  const int ic = blockIdx.x;
  const int ntime = S.ntime, nd = S.nd, nupd = S.nupd;
  const SolvePar P = {S.maxiter, A.mcap, S.abstol, S.reltol};
  const bool gm = S.linsolve == QB_LINSOLVE_GMRES;
  const bool recompute = !stored && (S.flags & (QB_FLAG_RECOMPUTE_STAGE | QB_FLAG_STORE_STAGE)) != 0;
  SolveStats stats = {0, 0};
  const unsigned rbars = sm.base + L.bars, tbars = sm.base + L.bars + 8u * kNR;
  const unsigned recring = sm.base + L.rec, trajring = sm.base + L.traj, sgrad = sm.base + L.grad, supd = sm.base + L.upd;
  const char* recg = reinterpret_cast<const char*>(A.rec);
  const size_t recstride_b = (size_t)A.recstride * 8;
  /* primal ring: load q reads prim[-q]; states: q = 0 is x_ntime and step n = ntime-r uses q = r+1 (x_{n-1});
   * stored stages: step n = ntime-r uses q = r (stage of step n-1 -> n) */
  const double2* prim = stored ? A.stage_traj + ((size_t)ic * ntime + (ntime - 1)) * C::DIM
                               : (A.traj ? A.traj + ((size_t)ic * (ntime + 1) + ntime) * C::DIM : nullptr);
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
    if (have_traj) {
      for (int q = 0; q < kNT && q <= qlast; q++) {
        mbar_expect_tx(tbars + 8u * q, C::VEC);
        bulk_g2s(trajring + (unsigned)q * C::VEC, prim - (size_t)q * C::DIM, C::VEC, tbars + 8u * q);
      }
    }
  }
  for (int i = th.tid; i < nd; i += C::T) sts1(sgrad + 8u * i, 0.0);
  for (int i = th.tid; i < nupd * 4; i += C::T) asm volatile("st.shared.s32 [%0], %1;" ::"r"(supd + 4u * i), "r"(A.upd[i]) : "memory");
  double2 xadj[E], xcur[E], xprev[E];
#pragma unroll
  for (int e = 0; e < E; e++) xadj[e] = A.state[(size_t)ic * C::DIM + e * C::T + th.tid];
  __syncthreads();
  if constexpr (stored) {
#pragma unroll
    for (int e = 0; e < E; e++) xcur[e] = make_double2(0.0, 0.0);
  } else if (have_traj) {
    mbar_wait(tbars, 0u);     /* q = 0: x_ntime */
#pragma unroll
    for (int e = 0; e < E; e++) xcur[e] = lds2(trajring + th.own + (unsigned)e * C::ESTR);
    __syncthreads();
    if (th.tid == 0 && kNT <= qlast) {   /* slot 0 is free again: load q = kNT */
      mbar_expect_tx(tbars, C::VEC);
      bulk_g2s(trajring, prim - (size_t)kNT * C::DIM, C::VEC, tbars);
    }
  } else {
#pragma unroll
    for (int e = 0; e < E; e++) xcur[e] = A.final_state[(size_t)ic * C::DIM + e * C::T + th.tid];
  }

  Co<Q, LIND, E> co;
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
    if (stored || have_traj) {
      /* primal state at n-1 from the stored trajectory (Lindblad), src/timestepper.cpp:230 */
      mbar_wait(tbars + 8u * tslot, (unsigned)((q / kNT) & 1));
#pragma unroll
      for (int e = 0; e < E; e++) xprev[e] = lds2(trajring + (unsigned)tslot * C::VEC + th.own + (unsigned)e * C::ESTR);
    } else if constexpr (!stored) {
      /* Schroedinger: one IMR step backwards, src/timestepper.cpp:231 */
#pragma unroll
      for (int e = 0; e < E; e++) sts2(sm.xb + th.own + (unsigned)e * C::ESTR, xcur[e]);
      csync<NW>();
      double2 rhs[E], k[E];
      apply<Q, LIND, E, false>(th, co, sm.xb, xcur, rhs);
      stats.evals++;
      linsolve<Q, LIND, E, false>(th, co, P, gm, sm, rhs, -dtn / 2.0, k, stats);
#pragma unroll
      for (int e = 0; e < E; e++) xprev[e] = make_double2(xcur[e].x - dtn * k[e].x, xcur[e].y - dtn * k[e].y);
    }
    /* adjoint stage: (I - dtn/2 M)^T kbar = xadj ; kbar *= dtn */
    double2 kbar[E];
    linsolve<Q, LIND, E, true>(th, co, P, gm, sm, xadj, dtn / 2.0, kbar, stats);
#pragma unroll
    for (int e = 0; e < E; e++) { kbar[e].x *= dtn; kbar[e].y *= dtn; }
    /* primal stage x_{n-1} + dtn/2 k */
    double2 stage[E];
    if constexpr (stored) {
#pragma unroll
      for (int e = 0; e < E; e++) stage[e] = xprev[e];
    } else if (recompute) {
      csync<NW>();    /* xb may still be read by a slower warp (restart path / previous step) */
#pragma unroll
      for (int e = 0; e < E; e++) sts2(sm.xb + th.own + (unsigned)e * C::ESTR, xprev[e]);
      csync<NW>();
      double2 rhs[E], k[E];
      apply<Q, LIND, E, false>(th, co, sm.xb, xprev, rhs);
      stats.evals++;
      linsolve<Q, LIND, E, false>(th, co, P, gm, sm, rhs, dtn / 2.0, k, stats);
#pragma unroll
      for (int e = 0; e < E; e++) stage[e] = make_double2(xprev[e].x + dtn / 2.0 * k[e].x, xprev[e].y + dtn / 2.0 * k[e].y);
    } else {
#pragma unroll
      for (int e = 0; e < E; e++) stage[e] = make_double2(0.5 * (xprev[e].x + xcur[e].x), 0.5 * (xprev[e].y + xcur[e].y));
    }
    csync<NW>();
    const unsigned xs = sm.xb, ys = sm.xb + C::VEC;
#pragma unroll
    for (int e = 0; e < E; e++) { sts2(xs + th.own + (unsigned)e * C::ESTR, stage[e]); sts2(ys + th.own + (unsigned)e * C::ESTR, kbar[e]); }
    csync<NW>();
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
      qbk::warp_sum_n<2 * Q>(cpq);
      const unsigned red = sm.red;
      if (th.lane == 0) {
#pragma unroll
        for (int i = 0; i < 2 * Q; i++) sts1(red + (unsigned)(i * NW + th.warp) * 8u, cpq[i]);
      }
      /* xadj += M^T kbar */
      double2 mt[E];
      apply<Q, LIND, E, true>(th, co, ys, kbar, mt);
      stats.evals++;
#pragma unroll
      for (int e = 0; e < E; e++) { xadj[e].x += mt[e].x; xadj[e].y += mt[e].y; }
      csync<NW>();
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
          const double cp = qbk::partial_sum<NW>(red + (unsigned)((2 * uk) * NW) * 8u);
          const double cq = qbk::partial_sum<NW>(red + (unsigned)((2 * uk + 1) * NW) * 8u);
          const double pbar = 1.0 * cp, qbar = 1.0 * cq;
          const double B1bar = g_sn * qbar + g_cs * pbar;
          const double B2bar = g_cs * qbar - g_sn * pbar;
          const unsigned gaddr = sgrad + 8u * (unsigned)(S.osc_off[uk] + gi_idx + uf * gi_ns * 2);
          sts1(gaddr, lds1(gaddr) + g_bval * B1bar);
          sts1(gaddr + 8u * gi_ns, lds1(gaddr + 8u * gi_ns) + g_bval * B2bar);
        }
      }
      sm.red = sm.red_other;
      sm.red_other = red;
    }
    csync<NW>();   /* record, trajectory slot, exchange buffers and gradient of this step are done with */
    if (th.tid == 0) {
      if (r + kNR < ntime) {
        mbar_expect_tx(rbars + 8u * rslot, A.rec_bytes);
        bulk_g2s(recring + (unsigned)rslot * A.rec_bytes, recg + (size_t)(ntime - 1 - (r + kNR)) * recstride_b, A.rec_bytes, rbars + 8u * rslot);
      }
      if (have_traj && q + kNT <= qlast) {
        mbar_expect_tx(tbars + 8u * tslot, C::VEC);
        bulk_g2s(trajring + (unsigned)tslot * C::VEC, prim - (size_t)(q + kNT) * C::DIM, C::VEC, tbars + 8u * tslot);
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
