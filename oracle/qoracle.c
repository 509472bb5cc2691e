/*
 * qoracle.c -- CPU restatement of the reference's hot path in plain C.  TEST INFRASTRUCTURE ONLY (see qoracle.h).
 *
 * Every function names the reference lines it follows (paths relative to the reference root).  The reference
 * instantiates one template per level tuple (src/mastereq.cpp:3039-3151); this restatement is generic in the number
 * of oscillators and levels and keeps the reference's loop order (ket indices outer, bra indices inner, last
 * oscillator fastest, src/mastereq.cpp:2123-2130) and the order in which the terms are added to each output element.
 */
#include "qoracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int qo_N(const qb_system_desc* d) {
  int N = 1;
  for (int k = 0; k < d->nosc; k++) N *= d->nlevels[k];
  return N;
}
int qo_dim(const qb_system_desc* d) {
  int N = qo_N(d);
  return d->lindblad == QB_LINDBLAD_NONE ? N : N * N;
}
static int qo_osc_nparams(const qb_system_desc* d, int k) {
  int n = 0, nc = d->carrier_off[k + 1] - d->carrier_off[k];
  for (int s = d->seg_off[k]; s < d->seg_off[k + 1]; s++) n += 2 * d->segments[s].nsplines * nc;
  return n;
}
int qo_ndesign(const qb_system_desc* d) {
  int n = 0;
  for (int k = 0; k < d->nosc; k++) n += qo_osc_nparams(d, k);
  return n;
}
static int qo_osc_offset(const qb_system_desc* d, int k) {
  int n = 0;
  for (int j = 0; j < k; j++) n += qo_osc_nparams(d, j);
  return n;
}

/* BSpline2nd::basisfunction, src/controlbasis.cpp:81-96 (tcenter from the constructor, :20-33) */
static double bspline2_basis(const qb_segment* s, int id, double t) {
  double dtknot = (s->tstop - s->tstart) / (double)(s->nsplines - 2);
  double width = 3.0 * dtknot;
  double tcenter = s->tstart + dtknot * ((id + 1) - 1.5);
  double tau = (t - tcenter) / width;
  if (tau < -1. / 2. || tau >= 1. / 2.) return 0.0;
  double val = 0.0;
  if (-1. / 2. <= tau && tau < -1. / 6.) val = 9. / 8. + 9. / 2. * tau + 9. / 2. * pow(tau, 2);
  else if (-1. / 6. <= tau && tau < 1. / 6.) val = 3. / 4. - 9. * pow(tau, 2);
  else if (1. / 6. <= tau && tau < 1. / 2.) val = 9. / 8. - 9. / 2. * tau + 9. / 2. * pow(tau, 2);
  return val;
}

/* ControlBasis::evaluate: BSpline2nd src/controlbasis.cpp:47-66, BSpline0 :229-243 */
static void basis_evaluate(const qb_segment* s, double t, const double* coeff, int f, double* B1, double* B2) {
  int ns = s->nsplines;
  if (s->type == QB_CTRL_BSPLINE2ND) {
    double sum1 = 0.0, sum2 = 0.0;
    for (int l = 0; l < ns; l++) {
      if (s->enforce_bc && (l <= 1 || l >= ns - 2)) continue;
      double Blt = bspline2_basis(s, l, t);
      sum1 += coeff[s->skip + f * ns * 2 + l] * Blt;
      sum2 += coeff[s->skip + f * ns * 2 + l + ns] * Blt;
    }
    *B1 = sum1; *B2 = sum2;
  } else { /* BSpline0 */
    double dtknot = (s->tstop - s->tstart) / (ns - 1.0);
    int id = (int)ceil((t - s->tstart) / dtknot - 0.5);
    if (id < 0 || id >= ns) { *B1 = 0.0; *B2 = 0.0; }
    else { *B1 = coeff[s->skip + f * ns * 2 + id]; *B2 = coeff[s->skip + f * ns * 2 + id + ns]; }
  }
}
/* ControlBasis::derivative: BSpline2nd src/controlbasis.cpp:68-79, BSpline0 :245-255 */
static void basis_derivative(const qb_segment* s, double t, double* cd, double v1, double v2, int f) {
  int ns = s->nsplines;
  if (s->type == QB_CTRL_BSPLINE2ND) {
    for (int l = 0; l < ns; l++) {
      if (s->enforce_bc && (l <= 1 || l >= ns - 2)) continue;
      double Blt = bspline2_basis(s, l, t);
      cd[s->skip + f * ns * 2 + l] += Blt * v1;
      cd[s->skip + f * ns * 2 + l + ns] += Blt * v2;
    }
  } else {
    double dtknot = (s->tstop - s->tstart) / (ns - 1.0);
    int id = (int)ceil((t - s->tstart) / dtknot - 0.5);
    if (id >= 0 && id < ns) { cd[s->skip + f * ns * 2 + id] += v1; cd[s->skip + f * ns * 2 + id + ns] += v2; }
  }
}

/* Oscillator::evalControl, src/oscillator.cpp:281-337 */
void qo_eval_control(const qb_system_desc* d, const double* alpha, int k, double t, double* p, double* q) {
  *p = 0.0; *q = 0.0;
  const double* params = alpha + qo_osc_offset(d, k);
  int nc = d->carrier_off[k + 1] - d->carrier_off[k];
  const double* om = d->carrier_freq + d->carrier_off[k];
  if (qo_osc_nparams(d, k) > 0) {
    for (int s = d->seg_off[k]; s < d->seg_off[k + 1]; s++) {
      const qb_segment* seg = &d->segments[s];
      if (seg->tstart <= t && seg->tstop >= t) {
        double sum_p = 0.0, sum_q = 0.0;
        for (int f = 0; f < nc; f++) {
          double B1 = 0.0, B2 = 0.0;
          basis_evaluate(seg, t, params, f, &B1, &B2);
          double c = cos(om[f] * t), sn = sin(om[f] * t);
          sum_p += c * B1 - sn * B2;
          sum_q += sn * B1 + c * B2;
        }
        *p = sum_p; *q = sum_q;
        break;
      }
    }
  }
  for (int i = d->pipulse_off[k]; i < d->pipulse_off[k + 1]; i++) {
    const double* pp = d->pipulse + 3 * i;
    if (pp[0] <= t && t <= pp[1]) { double a = pp[2] / sqrt(2.0); *p = a; *q = a; }
  }
}

/* Oscillator::evalControl_diff, src/oscillator.cpp:339-381 */
void qo_eval_control_diff(const qb_system_desc* d, int k, double t, double* grad, double pbar, double qbar) {
  double* g = grad + qo_osc_offset(d, k);
  int nc = d->carrier_off[k + 1] - d->carrier_off[k];
  const double* om = d->carrier_freq + d->carrier_off[k];
  if (qo_osc_nparams(d, k) > 0) {
    for (int s = d->seg_off[k]; s < d->seg_off[k + 1]; s++) {
      const qb_segment* seg = &d->segments[s];
      if (seg->tstart <= t && seg->tstop >= t) {
        for (int f = 0; f < nc; f++) {
          double c = cos(om[f] * t), sn = sin(om[f] * t);
          double B1bar = sn * qbar + c * pbar;
          double B2bar = c * qbar - sn * pbar;
          basis_derivative(seg, t, g, B1bar, B2bar, f);
        }
        break;
      }
    }
  }
}

/* ---- matrix-free RHS ------------------------------------------------------------------------------------- */
typedef struct {
  int Q, N, dim, lind;
  int n[QB_MAX_OSC], np[QB_MAX_OSC];        /* levels of bra index i_k and of ket index i_k' (1 if Schroedinger) */
  int stride[QB_MAX_OSC], stridep[QB_MAX_OSC];
  double decay[QB_MAX_OSC], dephase[QB_MAX_OSC];
} qo_geom;

/* strides: TensorGetIndex, include/mastereq.hpp:362-364,446-448,472-474,500-502; decay/dephase switches
 * src/mastereq.cpp:2070-2077 */
static void geom_init(qo_geom* g, const qb_system_desc* d) {
  g->Q = d->nosc; g->N = qo_N(d); g->dim = qo_dim(d); g->lind = d->lindblad;
  int addT1 = (d->lindblad == QB_LINDBLAD_DECAY || d->lindblad == QB_LINDBLAD_BOTH);
  int addT2 = (d->lindblad == QB_LINDBLAD_DEPHASE || d->lindblad == QB_LINDBLAD_BOTH);
  int s = 1;
  for (int k = d->nosc - 1; k >= 0; k--) {
    g->n[k] = d->nlevels[k];
    g->np[k] = d->lindblad == QB_LINDBLAD_NONE ? 1 : d->nlevels[k];
    g->stride[k] = s;
    g->stridep[k] = s * g->N;
    s *= d->nlevels[k];
    g->decay[k] = (d->decay_time[k] > 1e-14 && addT1) ? 1. / d->decay_time[k] : 0.0;
    g->dephase[k] = (d->dephase_time[k] > 1e-14 && addT2) ? 1. / d->dephase_time[k] : 0.0;
  }
}

/* H_detune + H_selfkerr + H_crosskerr, include/mastereq.hpp:316-326,376-403,... */
static double hdiag(const qb_system_desc* d, const int* i) {
  double detune = 0.0, kerr = 0.0, cross = 0.0;
  for (int k = 0; k < d->nosc; k++) detune += d->detuning[k] * i[k];
  for (int k = 0; k < d->nosc; k++) kerr += -d->selfkerr[k] / 2.0 * i[k] * (i[k] - 1);
  int id = 0;
  for (int k = 0; k < d->nosc; k++)
    for (int l = k + 1; l < d->nosc; l++) { cross += -d->crosskerr[id] * i[k] * i[l]; id++; }
  return detune + kerr + cross;
}

static void rhs_apply_pq(const qb_system_desc* d, const qo_geom* g, double t, const double* p, const double* q, int transpose,
                         const double* x, double* y) {
  const int Q = g->Q, dim = g->dim;
  double cosv[QB_MAX_PAIRS], sinv[QB_MAX_PAIRS];
  int npairs = Q * (Q - 1) / 2;
  for (int id = 0; id < npairs; id++) { cosv[id] = cos(d->eta[id] * t); sinv[id] = sin(d->eta[id] * t); }
  int i[QB_MAX_OSC], ip[QB_MAX_OSC];
  for (int it = 0; it < dim; it++) {
    /* digits: it = row + N*col, last oscillator fastest */
    int row = it % g->N, col = it / g->N;
    for (int k = Q - 1; k >= 0; k--) { i[k] = row % g->n[k]; row /= g->n[k]; }
    for (int k = Q - 1; k >= 0; k--) { ip[k] = g->lind ? col % g->n[k] : 0; if (g->lind) col /= g->n[k]; }
    double xre = x[it], xim = x[it + dim];
    double hd = hdiag(d, i);
    double hdp = g->lind ? hdiag(d, ip) : 0.0;
    double yre, yim;
    if (!transpose) { yre = (hd - hdp) * xim; yim = (-hd + hdp) * xre; }
    else            { yre = (-hd + hdp) * xim; yim = (hd - hdp) * xre; }
    if (g->lind) {
      double l1diag = 0.0, l2 = 0.0;
      for (int k = 0; k < Q; k++) l1diag += -g->decay[k] / 2.0 * (i[k] + ip[k]);
      for (int k = 0; k < Q; k++) l2 += g->dephase[k] * (i[k] * ip[k] - 1. / 2. * (i[k] * i[k] + ip[k] * ip[k]));
      yre += (l2 + l1diag) * xre;
      yim += (l2 + l1diag) * xim;
    }
    /* Jkl_coupling / Jkl_coupling_T, include/mastereq.hpp:632-675,701-741 */
    int id = 0;
    for (int k = 0; k < Q; k++) for (int l = k + 1; l < Q; l++, id++) {
      double J = d->Jkl[id], c = cosv[id], s = sinv[id];
      if (!(fabs(J) > 1e-10)) continue;
      int ni = g->n[k], nj = g->n[l], nip = g->np[k], njp = g->np[l];
      int a = i[k], ap = ip[k], b = i[l], bp = ip[l];
      int si = g->stride[k], sip = g->stridep[k], sj = g->stride[l], sjp = g->stridep[l];
      if (!transpose) {
        if (a > 0 && b < nj - 1) { int itx = it - si + sj; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a * (b + 1));
          yre += J * sq * (c * xi + s * xr); yim += J * sq * (-c * xr + s * xi); }
        if (a < ni - 1 && b > 0) { int itx = it + si - sj; double xr = x[itx], xi = x[itx + dim], sq = sqrt(b * (a + 1));
          yre += J * sq * (c * xi - s * xr); yim += J * sq * (-c * xr - s * xi); }
        if (ap > 0 && bp < njp - 1) { int itx = it - sip + sjp; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap * (bp + 1));
          yre += J * sq * (-c * xi + s * xr); yim += J * sq * (c * xr + s * xi); }
        if (ap < nip - 1 && bp > 0) { int itx = it + sip - sjp; double xr = x[itx], xi = x[itx + dim], sq = sqrt(bp * (ap + 1));
          yre += J * sq * (-c * xi - s * xr); yim += J * sq * (c * xr - s * xi); }
      } else {
        if (a < ni - 1 && b > 0) { int itx = it + si - sj; double xr = x[itx], xi = x[itx + dim], sq = sqrt(b * (a + 1));
          yre += J * sq * (-c * xi + s * xr); yim += J * sq * (c * xr + s * xi); }
        if (a > 0 && b < nj - 1) { int itx = it - si + sj; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a * (b + 1));
          yre += J * sq * (-c * xi - s * xr); yim += J * sq * (c * xr - s * xi); }
        if (ap < nip - 1 && bp > 0) { int itx = it + sip - sjp; double xr = x[itx], xi = x[itx + dim], sq = sqrt(bp * (ap + 1));
          yre += J * sq * (c * xi + s * xr); yim += J * sq * (-c * xr + s * xi); }
        if (ap > 0 && bp < njp - 1) { int itx = it - sip + sjp; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap * (bp + 1));
          yre += J * sq * (c * xi - s * xr); yim += J * sq * (-c * xr - s * xi); }
      }
    }
    /* L1decay / L1decay_T, include/mastereq.hpp:744-769 */
    if (g->lind) {
      for (int k = 0; k < Q; k++) {
        double dec = g->decay[k];
        if (!(fabs(dec) > 1e-12)) continue;
        if (!transpose) {
          if (i[k] < g->n[k] - 1 && ip[k] < g->n[k] - 1) {
            double l1off = dec * sqrt((i[k] + 1) * (ip[k] + 1));
            int itx = it + g->stride[k] + g->stridep[k];
            yre += l1off * x[itx]; yim += l1off * x[itx + dim];
          }
        } else {
          if (i[k] > 0 && ip[k] > 0) {
            double l1off = dec * sqrt(i[k] * ip[k]);
            int itx = it - g->stride[k] - g->stridep[k];
            yre += l1off * x[itx]; yim += l1off * x[itx + dim];
          }
        }
      }
    }
    /* control / control_T, include/mastereq.hpp:772-855,858-912 */
    for (int k = 0; k < Q; k++) {
      int n = g->n[k], np = g->np[k], a = i[k], ap = ip[k], si = g->stride[k], sip = g->stridep[k];
      double pt = p[k], qt = q[k];
      if (!transpose) {
        if (a < n - 1) { int itx = it + si; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a + 1);
          yre += sq * (pt * xi + qt * xr); yim += sq * (-pt * xr + qt * xi); }
        if (ap < np - 1) { int itx = it + sip; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap + 1);
          yre += sq * (-pt * xi + qt * xr); yim += sq * (pt * xr + qt * xi); }
        if (a > 0) { int itx = it - si; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a);
          yre += sq * (pt * xi - qt * xr); yim += sq * (-pt * xr - qt * xi); }
        if (ap > 0) { int itx = it - sip; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap);
          yre += sq * (-pt * xi - qt * xr); yim += sq * (pt * xr - qt * xi); }
      } else {
        if (a > 0) { int itx = it - si; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a);
          yre += sq * (-pt * xi + qt * xr); yim += sq * (pt * xr + qt * xi); }
        if (ap > 0) { int itx = it - sip; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap);
          yre += sq * (pt * xi + qt * xr); yim += sq * (-pt * xr + qt * xi); }
        if (a < n - 1) { int itx = it + si; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a + 1);
          yre += sq * (-pt * xi - qt * xr); yim += sq * (pt * xr - qt * xi); }
        if (ap < np - 1) { int itx = it + sip; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap + 1);
          yre += sq * (pt * xi - qt * xr); yim += sq * (-pt * xr - qt * xi); }
      }
    }
    y[it] = yre; y[it + dim] = yim;
  }
}

void qo_rhs_apply(const qb_system_desc* d, const double* alpha, double t, int transpose, const double* x, double* y) {
  qo_geom g; geom_init(&g, d);
  double p[QB_MAX_OSC], q[QB_MAX_OSC];
  for (int k = 0; k < d->nosc; k++) qo_eval_control(d, alpha, k, t, &p[k], &q[k]);
  rhs_apply_pq(d, &g, t, p, q, transpose, x, y);
}

/* compute_dRHS_dParams_matfree + dRHSdp_getcoeffs, src/mastereq.cpp:970-1276, include/mastereq.hpp:553-604 */
static void drhs_coeffs(const qo_geom* g, const double* x, const double* xbar, double* coeff_p, double* coeff_q) {
  const int Q = g->Q, dim = g->dim;
  for (int k = 0; k < Q; k++) { coeff_p[k] = 0.0; coeff_q[k] = 0.0; }
  int i[QB_MAX_OSC], ip[QB_MAX_OSC];
  for (int it = 0; it < dim; it++) {
    int row = it % g->N, col = it / g->N;
    for (int k = Q - 1; k >= 0; k--) { i[k] = row % g->n[k]; row /= g->n[k]; }
    for (int k = Q - 1; k >= 0; k--) { ip[k] = g->lind ? col % g->n[k] : 0; if (g->lind) col /= g->n[k]; }
    double xbarre = xbar[it], xbarim = xbar[it + dim];
    for (int k = 0; k < Q; k++) {
      int n = g->n[k], np = g->np[k], a = i[k], ap = ip[k], si = g->stride[k], sip = g->stridep[k];
      double pre = 0.0, pim = 0.0, qre = 0.0, qim = 0.0;
      if (a < n - 1) { int itx = it + si; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a + 1);
        pre += sq * xi; pim += -sq * xr; qre += sq * xr; qim += sq * xi; }
      if (ap < np - 1) { int itx = it + sip; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap + 1);
        pre += -sq * xi; pim += sq * xr; qre += sq * xr; qim += sq * xi; }
      if (a > 0) { int itx = it - si; double xr = x[itx], xi = x[itx + dim], sq = sqrt(a);
        pre += sq * xi; pim += -sq * xr; qre += -sq * xr; qim += -sq * xi; }
      if (ap > 0) { int itx = it - sip; double xr = x[itx], xi = x[itx + dim], sq = sqrt(ap);
        pre += -sq * xi; pim += sq * xr; qre += -sq * xr; qim += -sq * xi; }
      coeff_p[k] += pre * xbarre + pim * xbarim;
      coeff_q[k] += qre * xbarre + qim * xbarim;
    }
  }
}

void qo_drhs_dparams(const qb_system_desc* d, double t, const double* x, const double* xbar, double scale, double* grad) {
  qo_geom g; geom_init(&g, d);
  double cp[QB_MAX_OSC], cq[QB_MAX_OSC];
  drhs_coeffs(&g, x, xbar, cp, cq);
  for (int k = 0; k < d->nosc; k++) qo_eval_control_diff(d, k, t, grad, scale * cp[k], scale * cq[k]);
}

/* ---- linear solvers -------------------------------------------------------------------------------------- */
typedef struct {
  const qb_system_desc* d; qo_geom g; const double* alpha;
  double p[QB_MAX_OSC], q[QB_MAX_OSC], t;     /* assembled RHS (MasterEq::assemble_RHS) */
  int n2;                                      /* 2*dim */
  double *stage, *rhs, *stage_adj, *tmp, *err, *vv, *r; /* work vectors */
  long iters, evals;
} qo_ws;

static void ws_init(qo_ws* w, const qb_system_desc* d, const double* alpha) {
  w->d = d; geom_init(&w->g, d); w->alpha = alpha; w->n2 = 2 * w->g.dim; w->iters = 0; w->evals = 0;
  size_t n = (size_t)w->n2;
  w->stage = calloc(n, 8); w->rhs = calloc(n, 8); w->stage_adj = calloc(n, 8); w->tmp = calloc(n, 8); w->err = calloc(n, 8);
  w->r = calloc(n, 8);
  w->vv = calloc(n * 32, 8);
}
static void ws_free(qo_ws* w) { free(w->stage); free(w->rhs); free(w->stage_adj); free(w->tmp); free(w->err); free(w->r); free(w->vv); }

static void assemble(qo_ws* w, double t) { /* MasterEq::assemble_RHS, src/mastereq.cpp:657-678 */
  w->t = t;
  for (int k = 0; k < w->d->nosc; k++) qo_eval_control(w->d, w->alpha, k, t, &w->p[k], &w->q[k]);
}
static void matmult(qo_ws* w, int transpose, const double* x, double* y) {
  rhs_apply_pq(w->d, &w->g, w->t, w->p, w->q, transpose, x, y);
  w->evals++;
}

/* ImplMidpoint::NeumannSolve, src/timestepper.cpp:697-727: y = (I - alpha A)^-1 b by fixed-point iteration */
static int neumann(qo_ws* w, const double* b, double* y, double alpha, int transpose) {
  int n = w->n2;
  double errnorm = 0.0, errnorm0 = 0.0;
  memcpy(y, b, n * 8);
  int iter;
  for (iter = 0; iter < w->d->linsolve_maxiter; iter++) {
    memcpy(w->err, y, n * 8);
    matmult(w, transpose, y, w->tmp);
    for (int i = 0; i < n; i++) y[i] = 1.0 * b[i] + alpha * w->tmp[i];
    double s = 0.0;
    for (int i = 0; i < n; i++) { double e = w->err[i] - y[i]; s += e * e; }
    errnorm = sqrt(s);
    if (iter == 0) errnorm0 = errnorm;
    if (errnorm < w->d->linsolve_abstol) break;
    if (errnorm / errnorm0 < w->d->linsolve_reltol) break;
  }
  return iter;
}

/* PETSc KSPGMRES as the reference configures it (src/timestepper.cpp:541-550): PCNONE, zero initial guess, restart 30,
 * classical Gram-Schmidt without refinement, KSPConvergedDefault with ttol = max(rtol*||b||, abstol) on the recurrence
 * residual.  PETSc is a third-party dependency that is not in the reference tree (version unpinned; README names
 * 3.22.1); this follows its published algorithm.  Operator: (I - alpha M) or its transpose. */
static int gmres(qo_ws* w, const double* b, double* x, double alpha, int transpose) {
  const int n = w->n2, m = 30, maxits = w->d->linsolve_maxiter;
  double hh[32 * 31], cc[31], ss[31], grs[32], lhh[32], nrs[31];
#define HH(r, c) hh[(c) * 32 + (r)]
  double* vv = w->vv; double* r = w->r;
  for (int i = 0; i < n; i++) x[i] = 0.0;
  int its = 0, converged = 0, first = 1;
  double ttol = 0.0;
  while (!converged && its < maxits) {
    if (first) memcpy(r, b, n * 8);
    else {
      matmult(w, transpose, x, r);
      for (int i = 0; i < n; i++) r[i] = b[i] - (x[i] - alpha * r[i]);
    }
    double res = 0.0;
    for (int i = 0; i < n; i++) res += r[i] * r[i];
    res = sqrt(res);
    if (first) { ttol = fmax(w->d->linsolve_reltol * res, w->d->linsolve_abstol); first = 0; }
    if (res == 0.0) { converged = 1; break; }
    for (int i = 0; i < n; i++) vv[i] = r[i] / res;
    grs[0] = res;
    if (res <= ttol) { converged = 1; break; }
    int it = 0;
    while (!converged && it < m && its < maxits) {
      double* vin = vv + (size_t)it * n; double* vnew = vv + (size_t)(it + 1) * n;
      matmult(w, transpose, vin, vnew);
      for (int i = 0; i < n; i++) vnew[i] = vin[i] - alpha * vnew[i];
      for (int j = 0; j <= it; j++) { double s = 0.0; const double* vj = vv + (size_t)j * n; for (int i = 0; i < n; i++) s += vnew[i] * vj[i]; lhh[j] = s; }
      for (int j = 0; j <= it; j++) { const double* vj = vv + (size_t)j * n; double h = lhh[j]; for (int i = 0; i < n; i++) vnew[i] -= h * vj[i]; HH(j, it) = h; }
      double tt = 0.0;
      for (int i = 0; i < n; i++) tt += vnew[i] * vnew[i];
      tt = sqrt(tt);
      if (tt > 0.0) for (int i = 0; i < n; i++) vnew[i] /= tt;
      HH(it + 1, it) = tt;
      for (int j = 0; j < it; j++) {
        double t = HH(j, it);
        HH(j, it) = cc[j] * t + ss[j] * HH(j + 1, it);
        HH(j + 1, it) = -ss[j] * t + cc[j] * HH(j + 1, it);
      }
      double t = sqrt(HH(it, it) * HH(it, it) + HH(it + 1, it) * HH(it + 1, it));
      if (t == 0.0) { converged = 1; it++; its++; break; }
      cc[it] = HH(it, it) / t; ss[it] = HH(it + 1, it) / t;
      grs[it + 1] = -ss[it] * grs[it];
      grs[it] = cc[it] * grs[it];
      HH(it, it) = cc[it] * HH(it, it) + ss[it] * HH(it + 1, it);
      res = fabs(grs[it + 1]);
      it++; its++;
      if (res <= ttol) converged = 1;
    }
    if (it > 0) {
      for (int kk = it - 1; kk >= 0; kk--) {
        double tsum = grs[kk];
        for (int j = kk + 1; j < it; j++) tsum -= HH(kk, j) * nrs[j];
        nrs[kk] = tsum / HH(kk, kk);
      }
      for (int j = 0; j < it; j++) { const double* vj = vv + (size_t)j * n; double c = nrs[j]; for (int i = 0; i < n; i++) x[i] += c * vj[i]; }
    }
  }
#undef HH
  return its;
}

static void linsolve(qo_ws* w, const double* b, double* y, double halfdt, int transpose) {
  if (w->d->linsolve_type == QB_LINSOLVE_GMRES) w->iters += gmres(w, b, y, halfdt, transpose);
  else w->iters += neumann(w, b, y, halfdt, transpose);
}

/* ImplMidpoint::evolveFWD, src/timestepper.cpp:584-629 */
static void evolve_fwd(qo_ws* w, double tstart, double tstop, double* x) {
  double dt = tstop - tstart;
  assemble(w, (tstart + tstop) / 2.0);
  matmult(w, 0, x, w->rhs);
  linsolve(w, w->rhs, w->stage, dt / 2.0, 0);
  for (int i = 0; i < w->n2; i++) x[i] += dt * w->stage[i];
}

/* ImplMidpoint::evolveBWD, src/timestepper.cpp:631-694 */
static void evolve_bwd(qo_ws* w, double tstop, double tstart, const double* x, double* x_adj, double* grad) {
  double dt = tstop - tstart;
  double thalf = (tstart + tstop) / 2.0;
  assemble(w, (tstart + tstop) / 2.0);
  matmult(w, 0, x, w->rhs);
  linsolve(w, x_adj, w->stage_adj, dt / 2.0, 1);
  for (int i = 0; i < w->n2; i++) w->stage_adj[i] *= dt;
  linsolve(w, w->rhs, w->stage, dt / 2.0, 0);
  for (int i = 0; i < w->n2; i++) w->stage[i] = x[i] + dt / 2.0 * w->stage[i];
  {
    double cp[QB_MAX_OSC], cq[QB_MAX_OSC];
    drhs_coeffs(&w->g, w->stage, w->stage_adj, cp, cq);
    w->evals++;
    for (int k = 0; k < w->d->nosc; k++) qo_eval_control_diff(w->d, k, thalf, grad, 1.0 * cp[k], 1.0 * cq[k]);
  }
  matmult(w, 1, w->stage_adj, w->tmp);
  for (int i = 0; i < w->n2; i++) x_adj[i] += w->tmp[i];
}

/* CompositionalImplMidpoint gamma coefficients, src/timestepper.cpp:735-760 */
static int stage_gammas(int stepper, double* gamma) {
  if (stepper == QB_STEPPER_IMR8) {
    static const double g8[15] = {0.74167036435061295344822780, -0.40910082580003159399730010, 0.19075471029623837995387626,
      -0.57386247111608226665638773, 0.29906418130365592384446354, 0.33462491824529818378495798, 0.31529309239676659663205666,
      -0.79688793935291635401978884, 0.31529309239676659663205666, 0.33462491824529818378495798, 0.29906418130365592384446354,
      -0.57386247111608226665638773, 0.19075471029623837995387626, -0.40910082580003159399730010, 0.74167036435061295344822780};
    memcpy(gamma, g8, sizeof(g8));
    return 15;
  }
  if (stepper == QB_STEPPER_IMR4) {
    gamma[0] = 1. / (2. - pow(2., 1. / 3.)); gamma[1] = -pow(2., 1. / 3.) * gamma[0]; gamma[2] = 1. / (2. - pow(2., 1. / 3.));
    return 3;
  }
  gamma[0] = 1.0;
  return 1;
}

/* evolveFWD of the configured stepper: ImplMidpoint or CompositionalImplMidpoint::evolveFWD (src/timestepper.cpp:786-804) */
static void step_fwd(qo_ws* w, double tstart, double tstop, double* x) {
  if (w->d->stepper == QB_STEPPER_IMR) { evolve_fwd(w, tstart, tstop, x); return; }
  double gamma[15]; int ns = stage_gammas(w->d->stepper, gamma);
  double dt = tstop - tstart, tcurr = tstart;
  for (int s = 0; s < ns; s++) { double dts = gamma[s] * dt; evolve_fwd(w, tcurr, tcurr + dts, x); tcurr = tcurr + dts; }
}
/* CompositionalImplMidpoint::evolveBWD, src/timestepper.cpp:806-826 */
static void step_bwd(qo_ws* w, double tstop, double tstart, const double* x, double* x_adj, double* grad) {
  if (w->d->stepper == QB_STEPPER_IMR) { evolve_bwd(w, tstop, tstart, x, x_adj, grad); return; }
  double gamma[15]; int ns = stage_gammas(w->d->stepper, gamma);
  int n = w->n2;
  double dt = tstop - tstart, tcurr = tstart;
  double* xs = malloc((size_t)n * 8 * (ns + 1));
  double* aux = xs + (size_t)n * ns;
  memcpy(aux, x, n * 8);
  for (int s = 0; s < ns; s++) { memcpy(xs + (size_t)s * n, aux, n * 8); double dts = gamma[s] * dt; evolve_fwd(w, tcurr, tcurr + dts, aux); tcurr = tcurr + dts; }
  for (int s = ns - 1; s >= 0; s--) { double dts = gamma[s] * dt; evolve_bwd(w, tcurr, tcurr - dts, xs + (size_t)s * n, x_adj, grad); tcurr = tcurr - gamma[s] * dt; }
  free(xs);
}

/* TimeStepper::solveODE, src/timestepper.cpp:96-181 (penalty terms off) */
long qo_solve_ode(const qb_system_desc* d, const double* alpha, const double* rho0, double* rhoT, double* traj, long* rhs_evals) {
  qo_ws w; ws_init(&w, d, alpha);
  int n2 = w.n2;
  double dt = d->total_time / d->ntime;
  double* x = malloc(n2 * 8);
  memcpy(x, rho0, n2 * 8);
  for (int n = 0; n < d->ntime; n++) {
    double tstart = n * dt, tstop = (n + 1) * dt;
    if (traj) memcpy(traj + (size_t)n * n2, x, n2 * 8);
    step_fwd(&w, tstart, tstop, x);
  }
  if (traj) memcpy(traj + (size_t)d->ntime * n2, x, n2 * 8);
  memcpy(rhoT, x, n2 * 8);
  free(x);
  long it = w.iters;
  if (rhs_evals) *rhs_evals += w.evals;
  ws_free(&w);
  return it;
}

/* TimeStepper::solveAdjointODE, src/timestepper.cpp:184-253 (penalty terms off) */
long qo_solve_adjoint(const qb_system_desc* d, const double* alpha, const double* rhoT_bar, const double* finalstate,
                      const double* traj, double* redgrad, double* xbar0, long* rhs_evals) {
  qo_ws w; ws_init(&w, d, alpha);
  int n2 = w.n2, nd = qo_ndesign(d);
  double dt = d->total_time / d->ntime;
  for (int i = 0; i < nd; i++) redgrad[i] = 0.0;
  double* xadj = malloc(n2 * 8); double* xprimal = malloc(n2 * 8);
  memcpy(xadj, rhoT_bar, n2 * 8);
  memcpy(xprimal, finalstate, n2 * 8);
  for (int n = d->ntime; n > 0; n--) {
    double tstop = n * dt, tstart = (n - 1) * dt;
    if (traj) memcpy(xprimal, traj + (size_t)(n - 1) * n2, n2 * 8);
    else step_fwd(&w, tstop, tstart, xprimal);
    step_bwd(&w, tstop, tstart, xprimal, xadj, redgrad);
  }
  if (xbar0) memcpy(xbar0, xadj, n2 * 8);
  free(xadj); free(xprimal);
  long it = w.iters;
  if (rhs_evals) *rhs_evals += w.evals;
  ws_free(&w);
  return it;
}

int qo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

long qo_batch_forward(const qb_system_desc* d, const double* alpha, const double* rho0, int ninit, double* rhoT, int nthreads,
                      long* rhs_evals) {
  int n2 = 2 * qo_dim(d);
  long iters = 0, evals = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : iters, evals)
#endif
  for (int i = 0; i < ninit; i++) {
    long ev = 0;
    iters += qo_solve_ode(d, alpha, rho0 + (size_t)i * n2, rhoT + (size_t)i * n2, NULL, &ev);
    evals += ev;
  }
  if (rhs_evals) *rhs_evals += evals;
  return iters;
}

/* the IC loop of OptimProblem::evalGradF, src/optimproblem.cpp:386-444 (Lindblad) and :495-519 (Schroedinger);
 * G += redgrad is done in initial-condition order after the parallel loop. */
long qo_batch_gradient(const qb_system_desc* d, const double* alpha, const double* rho0, int ninit, double* rhoT,
                       const double* rhoT_bar_in, qo_terminal_cb cb, void* user, double* grad, int nthreads, long* rhs_evals) {
  int n2 = 2 * qo_dim(d), nd = qo_ndesign(d);
  int lind = d->lindblad != QB_LINDBLAD_NONE;
  double* allgrad = calloc((size_t)ninit * (nd > 0 ? nd : 1), 8);
  long iters = 0, evals = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : iters, evals)
#endif
  for (int i = 0; i < ninit; i++) {
    long ev = 0;
    double* traj = lind ? malloc((size_t)(d->ntime + 1) * n2 * 8) : NULL;
    double* bar = malloc(n2 * 8);
    iters += qo_solve_ode(d, alpha, rho0 + (size_t)i * n2, rhoT + (size_t)i * n2, traj, &ev);
    if (cb) cb(i, rhoT + (size_t)i * n2, bar, user);
    else memcpy(bar, rhoT_bar_in + (size_t)i * n2, n2 * 8);
    iters += qo_solve_adjoint(d, alpha, bar, rhoT + (size_t)i * n2, traj, allgrad + (size_t)i * nd, NULL, &ev);
    evals += ev;
    free(bar); free(traj);
  }
  for (int j = 0; j < nd; j++) grad[j] = 0.0;
  for (int i = 0; i < ninit; i++) for (int j = 0; j < nd; j++) grad[j] += allgrad[(size_t)i * nd + j];
  free(allgrad);
  if (rhs_evals) *rhs_evals += evals;
  return iters;
}
