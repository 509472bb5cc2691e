/*
 * qb_mono.cuh -- pieces of the one-reduction linear solver shared by the qubit-system kernels (qb_qubit.cuh) and the
 * large-state kernels (qb_big.cuh): the scalar part (L D L^T recurrence on the Gram matrix of the monomial Krylov basis,
 * residual norms, solution coefficients) and the recursive-halving warp reduction.
 */
#pragma once
#include <type_traits>
#include "qb_kernels.cuh"

namespace qmono {

constexpr int kSMax = 6;     /* longest Krylov basis the one-reduction solver builds; longer solves take the Arnoldi fallback */
constexpr int kSBlk = 4;     /* most basis vectors the qubit kernels build before the first reduction */
/* smallest pivot / diagonal ratio (= sin^2 of the angle between a new basis vector and the span of the earlier ones) the
 * factorisation accepts: the pivot carries a relative rounding error of about 1e-16 / ratio, so 1e-6 keeps the iterate within
 * 1e-10 of the exact GMRES iterate; below it the solve goes to the Arnoldi fallback */
constexpr double kPivotMin = 1e-6;

/* compile-time loop: f(std::integral_constant<int, i>) for i in [B, EN) */
template <int B, int EN, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < EN) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, EN>(f);
  }
}

/* Recursive-halving warp reduction of CNT values: at every butterfly stage a lane keeps one half of its values and
 * sends the other half, so NV values cost about NV shuffles in total instead of 5*NV.  After the five stages value
 * `halver_slot<NV>(lane)` is complete on exactly one lane. */
template <int CNT, int MASK>
struct Halver {
  static __device__ __forceinline__ double run(const double (&v)[CNT], int lane) {
    constexpr int H = (CNT + 1) / 2;
    const bool up = (lane & MASK) != 0;
    double nv[H];
#pragma unroll
    for (int i = 0; i < H; i++) {
      const double lo = v[i];
      const double hi = (H + i < CNT) ? v[H + i] : 0.0;
      const double keep = up ? hi : lo;
      const double send = up ? lo : hi;
      nv[i] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
    }
    if constexpr (MASK == 1) return nv[0];
    else return Halver<H, MASK / 2>::run(nv, lane);
  }
};
template <int NV>
__device__ __forceinline__ int halver_slot(int lane) {
  /* w: width of the value array at this stage (the same on every lane); real: how many of this lane's values are not padding */
  int w = NV, real = NV, idx = 0;
#pragma unroll
  for (int mask = 16; mask > 0; mask >>= 1) {
    const int h = (w + 1) / 2;
    if (lane & mask) { idx += h; real = real > h ? real - h : 0; }
    else real = real < h ? real : h;
    w = h;
  }
  return real >= 1 ? idx : -1;
}

/* ---- the one-reduction linear solver ----------------------------------------------------------------------------
 * (I - alpha M) y = b, zero initial guess, GMRES iterate y_m = argmin over K_m(A, b) of |b - A y| with PETSc's stopping rule
 * (KSPGMRES + KSPConvergedDefault as set up in src/timestepper.cpp:541-550: stop at the first m with rho_m <= max(rtol |b|, atol)).
 * K_m(A, b) = span{u_0 .. u_{m-1}} with the Neumann-scaled monomial basis u_j = (alpha M)^j b, and A u_j = u_j - u_{j+1}.  The
 * residual of y = sum_{j<m} c_j u_j is r = sum_{j<=m} e_j u_j with sum_j e_j = 1 (e_0 = 1 - c_0, e_j = c_{j-1} - c_j, e_m = c_{m-1}),
 * every such e is reachable, hence with the Gram matrix G_ij = <u_i, u_j> (leading (m+1) x (m+1) block G_m)
 *     rho_m^2 = min { e^T G_m e : 1^T e = 1 } = 1 / (1^T G_m^{-1} 1),     e = G_m^{-1} 1 * rho_m^2,     c_j = sum_{i>j} e_i.
 * With G = L D L^T (unit lower L) and z = L^{-1} 1:  1 / rho_m^2 = sum_{j<=m} z_j^2 / D_j, a running sum of positive terms
 * that is nested in m -- no cancellation although rho_m / |b| ~ 1e-10.  All of it is O(m^3) scalar work that every thread
 * repeats in registers; the only communication of a solve is ONE CTA-wide reduction of the (m+1)(m+2)/2 Gram entries
 * (classical Arnoldi: two reductions per iteration).  The basis is built speculatively to the length the previous solve of
 * the same kind needed; if that is not enough it is extended one vector (and one reduction) at a time.  Ill-conditioned bases
 * (pivot below kPivotMin of its diagonal entry) and solves longer than kSMax vectors go to the Arnoldi fallback `gmres_g`, which
 * restarts the solve from scratch exactly as the reference would run it.
 * A by-product: alpha M y = sum_{j<m} c_j u_{j+1} needs no further operator application (the adjoint step's M^T kbar). */
struct Mono {
  double Dinv[kSMax + 1], zp[kSMax + 1], L[(kSMax + 1) * kSMax / 2];   /* L[r][i] at r(r-1)/2 + i */
  double sigma;
};

/* one more row of the L D L^T factorisation from Gram row R = g[GOFF .. GOFF+R]; false: pivot unreliable */
template <int R, int GOFF, int NG>
__device__ __forceinline__ bool ldl_row(Mono& M, const double (&g)[NG]) {
  static_assert(GOFF + R < NG, "Gram row");
  double t[R > 0 ? R : 1];
  double d = g[GOFF + R];
#pragma unroll
  for (int i = 0; i < R; i++) {
    double ti = g[GOFF + i];
#pragma unroll
    for (int k = 0; k < i; k++) ti = fma(-t[k], M.L[i * (i - 1) / 2 + k], ti);
    t[i] = ti;
    const double l = ti * M.Dinv[i];
    M.L[R * (R - 1) / 2 + i] = l;
    d = fma(-ti, l, d);
  }
  const bool ok = d > kPivotMin * g[GOFF + R];
  const double dinv = qbk::fast_rcp(d);
  M.Dinv[R] = dinv;
  double z = 1.0;
#pragma unroll
  for (int i = 0; i < R; i++) z = fma(-M.L[R * (R - 1) / 2 + i], M.zp[i], z);
  M.zp[R] = z;
  if constexpr (R == 0) M.sigma = dinv;
  else M.sigma = fma(z * z, dinv, M.sigma);
  return ok;
}

/* c_j (j < m) of y_m = sum c_j u_j from the factorisation of G_m */
__device__ __forceinline__ void mono_coeffs(const Mono& M, int m, double (&c)[kSMax]) {
  double e[kSMax + 1];
#pragma unroll
  for (int j = kSMax; j >= 0; j--) {
    double s = M.zp[j] * M.Dinv[j];
#pragma unroll
    for (int i = j + 1; i <= kSMax; i++)
      if (i <= m) s = fma(-M.L[i * (i - 1) / 2 + j], e[i], s);
    e[j] = (j <= m) ? s : 0.0;
  }
  const double rho2 = qbk::fast_rcp(M.sigma);
  double suf = 0.0;
#pragma unroll
  for (int j = kSMax - 1; j >= 0; j--) {
    suf += e[j + 1];
    c[j] = (j < m) ? suf * rho2 : 0.0;
  }
}

}  // namespace qmono
