/*
 * shim.cpp -- implementation of the sequential PETSc/MPI/TAO stand-in declared in petsc_shim.h.
 * TEST INFRASTRUCTURE ONLY (see header).  Everything here is our own code; the only
 * algorithm restated from a third party is KSPGMRES (PETSc, version unpinned by the
 * reference; call sites /root/reference/src/timestepper.cpp:541-550,602,652,674).
 */
#include "petsc_shim.h"
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <sched.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

/* ======================================================================= MPI over fork() */
struct qshim_comm {
  std::vector<int> members; /* world ranks, ascending key order */
  int myidx = 0;            /* my index in members */
};
struct qshim_barrier {
  std::atomic<int> count;
  std::atomic<int> sense;
};
static const int QSHIM_MAXCOUNT = 1 << 16;
struct qshim_shared {
  qshim_barrier bars[64];
  int islots[64][4];
  double dslots[1]; /* np * QSHIM_MAXCOUNT doubles follow */
};
static qshim_shared* g_sh = nullptr;
static int g_np = 1, g_rank = 0;
static std::vector<pid_t> g_children;
static qshim_comm g_world, g_self;
MPI_Comm MPI_COMM_WORLD = &g_world;
MPI_Comm MPI_COMM_SELF = &g_self;
MPI_Comm PETSC_COMM_WORLD = &g_world;
static int g_local_sense[64];

static double* dslot(int r) { return g_sh->dslots + (size_t)r * QSHIM_MAXCOUNT; }

static void comm_barrier(MPI_Comm c) {
  int n = (int)c->members.size();
  if (n <= 1) return;
  int leader = c->members[0];
  qshim_barrier& b = g_sh->bars[leader];
  int s = g_local_sense[leader] = !g_local_sense[leader];
  if (b.count.fetch_add(1) + 1 == n) {
    b.count.store(0);
    b.sense.store(s);
  } else {
    while (b.sense.load() != s) sched_yield();
  }
}

int MPI_Init(int*, char***) {
  const char* e = getenv("QSHIM_NP");
  g_np = e ? atoi(e) : 1;
  if (g_np < 1) g_np = 1;
  if (g_np > 64) g_np = 64;
  g_rank = 0;
  if (g_np > 1) {
    size_t bytes = sizeof(qshim_shared) + (size_t)g_np * QSHIM_MAXCOUNT * sizeof(double);
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("qshim mmap"); exit(1); }
    g_sh = (qshim_shared*)p;
    for (int i = 0; i < 64; i++) { g_sh->bars[i].count.store(0); g_sh->bars[i].sense.store(0); g_local_sense[i] = 0; }
    fflush(stdout); fflush(stderr);
    for (int r = 1; r < g_np; r++) {
      pid_t pid = fork();
      if (pid < 0) { perror("qshim fork"); exit(1); }
      if (pid == 0) { g_rank = r; g_children.clear(); break; }
      g_children.push_back(pid);
    }
  }
  g_world.members.resize(g_np);
  for (int i = 0; i < g_np; i++) g_world.members[i] = i;
  g_world.myidx = g_rank;
  g_self.members.assign(1, g_rank);
  g_self.myidx = 0;
  return 0;
}
static void qshim_dump_stats();
int MPI_Finalize() {
  qshim_dump_stats();
  fflush(stdout); fflush(stderr);
  if (g_np > 1) {
    comm_barrier(MPI_COMM_WORLD);
    if (g_rank == 0) { for (pid_t c : g_children) { int st; waitpid(c, &st, 0); } }
  }
  return 0;
}
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = c->myidx; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = (int)c->members.size(); return 0; }
int MPI_Barrier(MPI_Comm c) { comm_barrier(c); return 0; }
int MPI_Abort(MPI_Comm, int code) { exit(code); }
double MPI_Wtime() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }

int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* out) {
  qshim_comm* nc = new qshim_comm();
  int n = (int)c->members.size();
  if (n <= 1) { nc->members = c->members; nc->myidx = 0; *out = nc; return 0; }
  g_sh->islots[g_rank][0] = color;
  g_sh->islots[g_rank][1] = key;
  comm_barrier(c);
  std::vector<std::pair<int, int>> grp; /* (key, world rank) */
  for (int i = 0; i < n; i++) { int w = c->members[i]; if (g_sh->islots[w][0] == color) grp.push_back({g_sh->islots[w][1], w}); }
  comm_barrier(c);
  std::stable_sort(grp.begin(), grp.end());
  for (size_t i = 0; i < grp.size(); i++) { nc->members.push_back(grp[i].second); if (grp[i].second == g_rank) nc->myidx = (int)i; }
  /* barriers are keyed on the lowest world rank of a group */
  std::vector<int> sorted = nc->members; std::sort(sorted.begin(), sorted.end());
  if (sorted[0] != nc->members[0]) { int me = nc->members[nc->myidx]; nc->members = sorted; for (size_t i = 0; i < sorted.size(); i++) if (sorted[i] == me) nc->myidx = (int)i; }
  *out = nc;
  return 0;
}
static size_t dt_size(MPI_Datatype dt) { return dt == MPI_DOUBLE ? sizeof(double) : sizeof(int); }
int MPI_Allreduce(const void* sb, void* rb, int count, MPI_Datatype dt, MPI_Op, MPI_Comm c) {
  int n = (int)c->members.size();
  if (n <= 1) { if (sb != rb) memmove(rb, sb, count * dt_size(dt)); return 0; }
  if (dt != MPI_DOUBLE || count > QSHIM_MAXCOUNT) { fprintf(stderr, "qshim: unsupported allreduce\n"); exit(1); }
  memcpy(dslot(g_rank), sb, count * sizeof(double));
  comm_barrier(c);
  double* r = (double*)rb;
  for (int i = 0; i < count; i++) r[i] = 0.0;
  for (int m = 0; m < n; m++) { const double* s = dslot(c->members[m]); for (int i = 0; i < count; i++) r[i] += s[i]; }
  comm_barrier(c);
  return 0;
}
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm c) {
  int n = (int)c->members.size();
  if (n <= 1) return 0;
  size_t bytes = count * dt_size(dt);
  if (bytes > QSHIM_MAXCOUNT * sizeof(double)) { fprintf(stderr, "qshim: bcast too large\n"); exit(1); }
  int wroot = c->members[root];
  if (g_rank == wroot) memcpy(dslot(wroot), buf, bytes);
  comm_barrier(c);
  if (g_rank != wroot) memcpy(buf, dslot(wroot), bytes);
  comm_barrier(c);
  return 0;
}
int MPI_Scatter(const void* sb, int sc, MPI_Datatype sdt, void* rb, int rc, MPI_Datatype, int root, MPI_Comm c) {
  int n = (int)c->members.size();
  size_t bytes = sc * dt_size(sdt);
  if (n <= 1) { memmove(rb, sb, bytes); return 0; }
  int wroot = c->members[root];
  if (bytes * n > QSHIM_MAXCOUNT * sizeof(double)) { fprintf(stderr, "qshim: scatter too large\n"); exit(1); }
  if (g_rank == wroot) memcpy(dslot(wroot), sb, bytes * n);
  comm_barrier(c);
  memcpy(rb, (char*)dslot(wroot) + bytes * c->myidx, bytes);
  (void)rc;
  comm_barrier(c);
  return 0;
}

/* ======================================================================= sys */
static long long g_matmult = 0;
extern "C" long long qshim_matmult_count() { return g_matmult; }
extern "C" void qshim_reset_counters() { g_matmult = 0; }
static void qshim_dump_stats() {
  const char* f = getenv("QSHIM_STATS");
  if (!f) return;
  char name[512];
  snprintf(name, 511, "%s.%d", f, g_rank);
  FILE* fp = fopen(name, "w");
  if (fp) { fprintf(fp, "{\"rank\": %d, \"shell_matmults\": %lld}\n", g_rank, g_matmult); fclose(fp); }
}

static _p_PetscViewer* g_viewer_dummy = nullptr;
PetscViewer PETSC_VIEWER_STDOUT_WORLD = nullptr;
PetscViewer PETSC_VIEWER_STDOUT_SELF = nullptr;
PetscErrorCode PetscInitialize(int*, char***, const char*, const char*) { return 0; }
PetscErrorCode PetscFinalize() { return 0; }
PetscErrorCode PetscOptionsSetValue(void*, const char*, const char*) { return 0; }
PetscErrorCode PetscPrintf(MPI_Comm c, const char* fmt, ...) {
  if (c->myidx == 0) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); }
  return 0;
}
PetscErrorCode PetscViewerPushFormat(PetscViewer, PetscViewerFormat) { return 0; }
PetscErrorCode PetscViewerPopFormat(PetscViewer) { return 0; }
PetscErrorCode PetscViewerCreate(MPI_Comm, PetscViewer* v) { *v = g_viewer_dummy; return 0; }
PetscErrorCode PetscViewerSetType(PetscViewer, PetscViewerType) { return 0; }
PetscErrorCode PetscViewerFileSetMode(PetscViewer, PetscFileMode) { return 0; }
PetscErrorCode PetscViewerFileSetName(PetscViewer, const char*) { return 0; }
PetscErrorCode PetscViewerDestroy(PetscViewer* v) { *v = nullptr; return 0; }
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm, const char*, PetscFileMode, PetscViewer* v) { *v = g_viewer_dummy; return 0; }

/* ======================================================================= IS */
struct _p_IS { int n, first, step; };
PetscErrorCode ISCreateStride(MPI_Comm, PetscInt n, PetscInt first, PetscInt step, IS* is) { *is = new _p_IS{n, first, step}; return 0; }
PetscErrorCode ISDestroy(IS* is) { delete *is; *is = nullptr; return 0; }

/* ======================================================================= Vec */
struct _p_Vec {
  int n = 0;
  double* a = nullptr;
  bool owns = true;
};
static void vec_alloc(Vec v) {
  if (v->a || v->n <= 0) return;
  v->a = (double*)calloc(v->n, sizeof(double));
  v->owns = true;
}
PetscErrorCode VecCreate(MPI_Comm, Vec* v) { *v = new _p_Vec(); return 0; }
PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec* v) { *v = new _p_Vec(); (*v)->n = n; vec_alloc(*v); return 0; }
PetscErrorCode VecSetSizes(Vec v, PetscInt nl, PetscInt ng) { v->n = (ng >= 0 ? ng : nl); return 0; }
PetscErrorCode VecSetFromOptions(Vec v) { vec_alloc(v); return 0; }
PetscErrorCode VecSetUp(Vec v) { vec_alloc(v); return 0; }
PetscErrorCode VecDuplicate(Vec v, Vec* w) { *w = new _p_Vec(); (*w)->n = v->n; vec_alloc(*w); return 0; }
PetscErrorCode VecDestroy(Vec* v) {
  if (!v || !*v) return 0;
  if ((*v)->owns) free((*v)->a);
  delete *v; *v = nullptr; return 0;
}
PetscErrorCode VecCopy(Vec x, Vec y) { vec_alloc(y); if (x->a != y->a) memcpy(y->a, x->a, sizeof(double) * x->n); return 0; }
PetscErrorCode VecZeroEntries(Vec v) { vec_alloc(v); memset(v->a, 0, sizeof(double) * v->n); return 0; }
PetscErrorCode VecSet(Vec v, PetscScalar s) { vec_alloc(v); for (int i = 0; i < v->n; i++) v->a[i] = s; return 0; }
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) { for (int i = 0; i < y->n; i++) y->a[i] += a * x->a[i]; return 0; }
PetscErrorCode VecAYPX(Vec y, PetscScalar b, Vec x) { for (int i = 0; i < y->n; i++) y->a[i] = x->a[i] + b * y->a[i]; return 0; }
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar al, PetscScalar be, PetscScalar ga, Vec x, Vec y) {
  for (int i = 0; i < z->n; i++) z->a[i] = al * x->a[i] + be * y->a[i] + ga * z->a[i];
  return 0;
}
PetscErrorCode VecScale(Vec v, PetscScalar s) { for (int i = 0; i < v->n; i++) v->a[i] *= s; return 0; }
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal* nrm) {
  double s = 0.0;
  if (t == NORM_1) { for (int i = 0; i < v->n; i++) s += fabs(v->a[i]); *nrm = s; }
  else if (t == NORM_INFINITY) { for (int i = 0; i < v->n; i++) s = std::max(s, fabs(v->a[i])); *nrm = s; }
  else { for (int i = 0; i < v->n; i++) s += v->a[i] * v->a[i]; *nrm = sqrt(s); }
  return 0;
}
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar* val) { double s = 0.0; for (int i = 0; i < x->n; i++) s += x->a[i] * y->a[i]; *val = s; return 0; }
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar* val) { return VecDot(x, y, val); }
PetscErrorCode VecGetArray(Vec v, PetscScalar** a) { vec_alloc(v); *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec, PetscScalar** a) { if (a) *a = nullptr; return 0; }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar** a) { vec_alloc(v); *a = v->a; return 0; }
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar** a) { if (a) *a = nullptr; return 0; }
PetscErrorCode VecGetValues(Vec v, PetscInt ni, const PetscInt ix[], PetscScalar y[]) { for (int i = 0; i < ni; i++) y[i] = v->a[ix[i]]; return 0; }
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode m) {
  vec_alloc(v);
  for (int i = 0; i < ni; i++) { if (ix[i] < 0) continue; if (m == ADD_VALUES) v->a[ix[i]] += y[i]; else v->a[ix[i]] = y[i]; }
  return 0;
}
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar y, InsertMode m) { return VecSetValues(v, 1, &i, &y, m); }
PetscErrorCode VecAssemblyBegin(Vec) { return 0; }
PetscErrorCode VecAssemblyEnd(Vec) { return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt* n) { *n = v->n; return 0; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n) { *n = v->n; return 0; }
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt* lo, PetscInt* hi) { if (lo) *lo = 0; if (hi) *hi = v->n; return 0; }
PetscErrorCode VecGetSubVector(Vec v, IS is, Vec* sub) {
  if (is->step != 1) { fprintf(stderr, "qshim: only unit-stride IS supported\n"); exit(1); }
  Vec s = new _p_Vec(); s->n = is->n; s->a = v->a + is->first; s->owns = false; *sub = s; return 0;
}
PetscErrorCode VecRestoreSubVector(Vec, IS, Vec* sub) { delete *sub; *sub = nullptr; return 0; }
PetscErrorCode VecISCopy(Vec vfull, IS is, ScatterMode mode, Vec vred) {
  for (int i = 0; i < is->n; i++) {
    if (mode == SCATTER_FORWARD) vfull->a[is->first + i * is->step] = vred->a[i];
    else vred->a[i] = vfull->a[is->first + i * is->step];
  }
  return 0;
}
PetscErrorCode VecView(Vec v, PetscViewer) { for (int i = 0; i < v->n; i++) printf("%1.14e\n", v->a[i]); return 0; }
PetscErrorCode VecScatterCreateToZero(Vec v, VecScatter* s, Vec* out) { *s = nullptr; return VecDuplicate(v, out); }
PetscErrorCode VecScatterBegin(VecScatter, Vec x, Vec y, InsertMode, ScatterMode) { return VecCopy(x, y); }
PetscErrorCode VecScatterEnd(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
PetscErrorCode VecScatterDestroy(VecScatter* s) { *s = nullptr; return 0; }

/* ======================================================================= Mat */
struct _p_Mat {
  enum Kind { AIJ, DENSE, SHELL } kind = AIJ;
  int m = 0, n = 0;
  std::vector<std::map<int, double>> rows; /* AIJ */
  std::vector<double> dense;               /* DENSE, row-major */
  void* ctx = nullptr;                     /* SHELL */
  int (*mult)(Mat, Vec, Vec) = nullptr;
  int (*multT)(Mat, Vec, Vec) = nullptr;
  double vscale = 1.0, vshift = 0.0;
  std::vector<int> rcols; std::vector<double> rvals; /* MatGetRow scratch */
  std::vector<double> tmp;
};
static void mat_setup(Mat A) {
  if (A->kind == _p_Mat::AIJ && (int)A->rows.size() != A->m) A->rows.resize(A->m);
  if (A->kind == _p_Mat::DENSE && (int)A->dense.size() != A->m * A->n) A->dense.assign((size_t)A->m * A->n, 0.0);
}
PetscErrorCode MatCreate(MPI_Comm, Mat* A) { *A = new _p_Mat(); return 0; }
PetscErrorCode MatCreateSeqDense(MPI_Comm, PetscInt m, PetscInt n, PetscScalar*, Mat* A) {
  *A = new _p_Mat(); (*A)->kind = _p_Mat::DENSE; (*A)->m = m; (*A)->n = n; mat_setup(*A); return 0;
}
PetscErrorCode MatCreateShell(MPI_Comm, PetscInt, PetscInt, PetscInt M, PetscInt N, void* ctx, Mat* A) {
  *A = new _p_Mat(); (*A)->kind = _p_Mat::SHELL; (*A)->m = M; (*A)->n = N; (*A)->ctx = ctx; return 0;
}
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void)) {
  if (op == MATOP_MULT) A->mult = (int (*)(Mat, Vec, Vec))f;
  else if (op == MATOP_MULT_TRANSPOSE) A->multT = (int (*)(Mat, Vec, Vec))f;
  return 0;
}
PetscErrorCode MatShellGetContext(Mat A, void* ctx) { *(void**)ctx = A->ctx; return 0; }
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N) { A->m = (M >= 0 ? M : m); A->n = (N >= 0 ? N : n); return 0; }
PetscErrorCode MatSetType(Mat A, MatType t) { if (t && !strcmp(t, MATSEQDENSE)) A->kind = _p_Mat::DENSE; return 0; }
PetscErrorCode MatSetFromOptions(Mat) { return 0; }
PetscErrorCode MatSetUp(Mat A) { mat_setup(A); return 0; }
PetscErrorCode MatSetOptionsPrefix(Mat, const char*) { return 0; }
PetscErrorCode MatSetOption(Mat, MatOption, PetscBool) { return 0; }
PetscErrorCode MatMPIAIJSetPreallocation(Mat A, PetscInt, const PetscInt*, PetscInt, const PetscInt*) { mat_setup(A); return 0; }
PetscErrorCode MatSeqAIJSetPreallocation(Mat A, PetscInt, const PetscInt*) { mat_setup(A); return 0; }
PetscErrorCode MatDestroy(Mat* A) { if (A && *A) { delete *A; *A = nullptr; } return 0; }
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt im[], PetscInt n, const PetscInt in[], const PetscScalar v[], InsertMode mode) {
  mat_setup(A);
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    int r = im[i], c = in[j]; double val = v[i * n + j];
    if (r < 0 || c < 0) continue;
    if (A->kind == _p_Mat::DENSE) { double& d = A->dense[(size_t)r * A->n + c]; if (mode == ADD_VALUES) d += val; else d = val; }
    else { if (mode == ADD_VALUES) A->rows[r][c] += val; else A->rows[r][c] = val; }
  }
  return 0;
}
PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode m) { return MatSetValues(A, 1, &i, 1, &j, &v, m); }
PetscErrorCode MatGetValues(Mat A, PetscInt m, const PetscInt im[], PetscInt n, const PetscInt in[], PetscScalar v[]) {
  mat_setup(A);
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    int r = im[i], c = in[j];
    if (A->kind == _p_Mat::DENSE) v[i * n + j] = A->dense[(size_t)r * A->n + c];
    else { auto it = A->rows[r].find(c); v[i * n + j] = (it == A->rows[r].end()) ? 0.0 : it->second; }
  }
  return 0;
}
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType) {
  /* PETSc's MatAssemblyEnd_Shell resets the scale and shift of a shell matrix; ImplMidpoint relies on it
     (/root/reference/src/timestepper.cpp:600-601,617-618,686-689). */
  if (A->kind == _p_Mat::SHELL) { A->vscale = 1.0; A->vshift = 0.0; }
  mat_setup(A);
  return 0;
}
PetscErrorCode MatGetRow(Mat A, PetscInt row, PetscInt* ncols, const PetscInt** cols, const PetscScalar** vals) {
  mat_setup(A);
  A->rcols.clear(); A->rvals.clear();
  if (A->kind == _p_Mat::DENSE) { for (int c = 0; c < A->n; c++) { A->rcols.push_back(c); A->rvals.push_back(A->dense[(size_t)row * A->n + c]); } }
  else { for (auto& kv : A->rows[row]) { A->rcols.push_back(kv.first); A->rvals.push_back(kv.second); } }
  if (ncols) *ncols = (int)A->rcols.size();
  if (cols) *cols = A->rcols.data();
  if (vals) *vals = A->rvals.data();
  return 0;
}
PetscErrorCode MatRestoreRow(Mat, PetscInt, PetscInt*, const PetscInt**, const PetscScalar**) { return 0; }
PetscErrorCode MatGetSize(Mat A, PetscInt* m, PetscInt* n) { if (m) *m = A->m; if (n) *n = A->n; return 0; }
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt* lo, PetscInt* hi) { if (lo) *lo = 0; if (hi) *hi = A->m; return 0; }
PetscErrorCode MatCreateVecs(Mat A, Vec* right, Vec* left) {
  if (right) VecCreateSeq(MPI_COMM_SELF, A->n, right);
  if (left) VecCreateSeq(MPI_COMM_SELF, A->m, left);
  return 0;
}
/* y (+)= op(A) x for assembled types; out may not alias x */
static void mat_apply(Mat A, const double* x, double* y, bool transpose, bool add) {
  mat_setup(A);
  int ny = transpose ? A->n : A->m;
  if (!add) for (int i = 0; i < ny; i++) y[i] = 0.0;
  if (A->kind == _p_Mat::DENSE) {
    for (int r = 0; r < A->m; r++) for (int c = 0; c < A->n; c++) {
      double v = A->dense[(size_t)r * A->n + c];
      if (transpose) y[c] += v * x[r]; else y[r] += v * x[c];
    }
  } else {
    for (int r = 0; r < A->m; r++) for (auto& kv : A->rows[r]) {
      if (transpose) y[kv.first] += kv.second * x[r]; else y[r] += kv.second * x[kv.first];
    }
  }
}
static PetscErrorCode mat_mult_any(Mat A, Vec x, Vec y, bool transpose) {
  if (A->kind == _p_Mat::SHELL) {
    g_matmult++;
    int (*f)(Mat, Vec, Vec) = transpose ? A->multT : A->mult;
    if (!f) { fprintf(stderr, "qshim: shell op missing\n"); exit(1); }
    if (x == y) { fprintf(stderr, "qshim: aliased shell mult\n"); exit(1); }
    f(A, x, y);
    if (A->vscale != 1.0 || A->vshift != 0.0) for (int i = 0; i < y->n; i++) y->a[i] = A->vscale * y->a[i] + A->vshift * x->a[i];
    return 0;
  }
  if (x->a == y->a) { A->tmp.assign(x->a, x->a + x->n); mat_apply(A, A->tmp.data(), y->a, transpose, false); }
  else mat_apply(A, x->a, y->a, transpose, false);
  return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return mat_mult_any(A, x, y, false); }
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y) { return mat_mult_any(A, x, y, true); }
static PetscErrorCode mat_multadd_any(Mat A, Vec x, Vec y, Vec z, bool transpose) {
  /* z = y + op(A) x ; z may alias y */
  int nz = z->n;
  std::vector<double> t(nz);
  _p_Vec tv; tv.n = nz; tv.a = t.data(); tv.owns = false;
  mat_mult_any(A, x, &tv, transpose);
  for (int i = 0; i < nz; i++) z->a[i] = y->a[i] + t[i];
  return 0;
}
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z) { return mat_multadd_any(A, x, y, z, false); }
PetscErrorCode MatMultTransposeAdd(Mat A, Vec x, Vec y, Vec z) { return mat_multadd_any(A, x, y, z, true); }
PetscErrorCode MatScale(Mat A, PetscScalar a) {
  if (A->kind == _p_Mat::SHELL) { A->vscale *= a; A->vshift *= a; }
  else if (A->kind == _p_Mat::DENSE) { for (auto& v : A->dense) v *= a; }
  else { mat_setup(A); for (auto& r : A->rows) for (auto& kv : r) kv.second *= a; }
  return 0;
}
PetscErrorCode MatShift(Mat A, PetscScalar a) {
  if (A->kind == _p_Mat::SHELL) { A->vshift += a; }
  else if (A->kind == _p_Mat::DENSE) { for (int i = 0; i < std::min(A->m, A->n); i++) A->dense[(size_t)i * A->n + i] += a; }
  else { mat_setup(A); for (int i = 0; i < std::min(A->m, A->n); i++) A->rows[i][i] += a; }
  return 0;
}
static double mat_get(Mat A, int r, int c) { double v; MatGetValues(A, 1, &r, 1, &c, &v); return v; }
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure) {
  mat_setup(Y); mat_setup(X);
  for (int r = 0; r < X->m; r++) {
    if (X->kind == _p_Mat::DENSE) { for (int c = 0; c < X->n; c++) { double v = a * X->dense[(size_t)r * X->n + c]; if (v != 0.0 || Y->kind == _p_Mat::DENSE) MatSetValue(Y, r, c, v, ADD_VALUES); } }
    else for (auto& kv : X->rows[r]) MatSetValue(Y, r, kv.first, a * kv.second, ADD_VALUES);
  }
  return 0;
}
PetscErrorCode MatNorm(Mat A, NormType, PetscReal* nrm) {
  mat_setup(A); double s = 0.0;
  if (A->kind == _p_Mat::DENSE) for (double v : A->dense) s += v * v;
  else for (auto& r : A->rows) for (auto& kv : r) s += kv.second * kv.second;
  *nrm = sqrt(s); return 0;
}
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat* B) {
  mat_setup(A);
  Mat C = new _p_Mat(); C->kind = A->kind; C->m = A->m; C->n = A->n; mat_setup(C);
  if (op == MAT_COPY_VALUES) { C->rows = A->rows; C->dense = A->dense; }
  *B = C; return 0;
}
PetscErrorCode MatConvert(Mat A, MatType, MatReuse, Mat* B) { return MatDuplicate(A, MAT_COPY_VALUES, B); }
PetscErrorCode MatTranspose(Mat A, MatReuse, Mat* B) {
  mat_setup(A);
  Mat C = new _p_Mat(); C->kind = A->kind; C->m = A->n; C->n = A->m; mat_setup(C);
  for (int r = 0; r < A->m; r++) {
    if (A->kind == _p_Mat::DENSE) for (int c = 0; c < A->n; c++) C->dense[(size_t)c * C->n + r] = A->dense[(size_t)r * A->n + c];
    else for (auto& kv : A->rows[r]) C->rows[kv.first][r] = kv.second;
  }
  *B = C; return 0;
}
PetscErrorCode MatTransposeMatMult(Mat A, Mat B, MatReuse, PetscReal, Mat* Cout) {
  mat_setup(A); mat_setup(B);
  Mat C = new _p_Mat(); C->kind = _p_Mat::DENSE; C->m = A->n; C->n = B->n; mat_setup(C);
  for (int i = 0; i < A->n; i++) for (int j = 0; j < B->n; j++) {
    double s = 0.0; for (int k = 0; k < A->m; k++) s += mat_get(A, k, i) * mat_get(B, k, j);
    C->dense[(size_t)i * C->n + j] = s;
  }
  *Cout = C; return 0;
}
PetscErrorCode MatIsTranspose(Mat A, Mat B, PetscReal tol, PetscBool* flg) {
  *flg = PETSC_TRUE;
  for (int r = 0; r < A->m; r++) for (int c = 0; c < A->n; c++) if (fabs(mat_get(A, r, c) - mat_get(B, c, r)) > tol) { *flg = PETSC_FALSE; return 0; }
  return 0;
}
PetscErrorCode MatIsSymmetric(Mat A, PetscReal tol, PetscBool* flg) { return MatIsTranspose(A, A, tol, flg); }
PetscErrorCode MatView(Mat A, PetscViewer) {
  mat_setup(A);
  for (int r = 0; r < A->m; r++) { for (int c = 0; c < A->n; c++) printf("%g ", mat_get(A, r, c)); printf("\n"); }
  return 0;
}
PetscErrorCode MatLoad(Mat, PetscViewer) { fprintf(stderr, "qshim: MatLoad unsupported\n"); return 1; }

/* ======================================================================= KSP (GMRES) */
struct _p_PC { int dummy; };
struct _p_KSP {
  Mat A = nullptr;
  double rtol = 1e-5, abstol = 1e-50, dtol = 1e4;
  int maxits = 10000;
  int restart = 30;
  double rnorm = 0.0;
  int its = 0;
  _p_PC pc;
  std::vector<std::vector<double>> vv; /* Krylov vectors */
};
PetscErrorCode KSPCreate(MPI_Comm, KSP* k) { *k = new _p_KSP(); return 0; }
PetscErrorCode KSPDestroy(KSP* k) { delete *k; *k = nullptr; return 0; }
PetscErrorCode KSPGetPC(KSP k, PC* pc) { *pc = &k->pc; return 0; }
PetscErrorCode PCSetType(PC, PCType) { return 0; }
PetscErrorCode KSPSetType(KSP, KSPType) { return 0; }
PetscErrorCode KSPSetTolerances(KSP k, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits) {
  if (rtol != (double)PETSC_DEFAULT) k->rtol = rtol;
  if (abstol != (double)PETSC_DEFAULT) k->abstol = abstol;
  if (dtol != (double)PETSC_DEFAULT) k->dtol = dtol;
  if (maxits != PETSC_DEFAULT) k->maxits = maxits;
  return 0;
}
PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat) { k->A = A; return 0; }
PetscErrorCode KSPSetFromOptions(KSP) { return 0; }
PetscErrorCode KSPGetResidualNorm(KSP k, PetscReal* r) { *r = k->rnorm; return 0; }
PetscErrorCode KSPGetIterationNumber(KSP k, PetscInt* its) { *its = k->its; return 0; }

/* Restated from PETSc KSPGMRES (KSPSolve_GMRES / KSPGMRESCycle / KSPGMRESClassicalGramSchmidtOrthogonalization with
 * KSP_GMRES_CGS_REFINE_NEVER / KSPGMRESUpdateHessenberg / KSPGMRESBuildSoln) with PCNONE, zero initial guess and
 * KSPConvergedDefault: ttol = max(rtol*||b||, abstol); converged when the recurrence residual <= ttol. */
static PetscErrorCode gmres_solve(KSP k, Vec b, Vec x, bool transpose) {
  const int n = b->n;
  const int m = k->restart;
  std::vector<double> hh((size_t)(m + 2) * (m + 1), 0.0); /* rotated Hessenberg, column-major hh[row + col*(m+2)] */
  std::vector<double> cc(m + 1), ss(m + 1), grs(m + 2), lhh(m + 2);
  auto HH = [&](int r, int c) -> double& { return hh[(size_t)c * (m + 2) + r]; };
  if ((int)k->vv.size() < m + 2) k->vv.resize(m + 2);
  for (auto& v : k->vv) if ((int)v.size() != n) v.assign(n, 0.0);
  _p_Vec vin, vout; vin.owns = vout.owns = false; vin.n = vout.n = n;

  for (int i = 0; i < n; i++) x->a[i] = 0.0;
  k->its = 0; k->rnorm = 0.0;
  double ttol = 0.0;
  bool converged = false, first = true;
  std::vector<double> r(n);
  while (!converged && k->its < k->maxits) {
    /* residual r = b - A x (x = 0 on the first cycle) */
    if (first) { for (int i = 0; i < n; i++) r[i] = b->a[i]; }
    else {
      vin.a = x->a; vout.a = r.data();
      mat_mult_any(k->A, &vin, &vout, transpose);
      for (int i = 0; i < n; i++) r[i] = b->a[i] - r[i];
    }
    double res = 0.0; for (int i = 0; i < n; i++) res += r[i] * r[i]; res = sqrt(res);
    k->rnorm = res;
    if (first) { ttol = std::max(k->rtol * res, k->abstol); first = false; }
    if (res == 0.0) { converged = true; break; }
    for (int i = 0; i < n; i++) k->vv[0][i] = r[i] / res;
    grs[0] = res;
    if (res <= ttol) { converged = true; break; }
    int it = 0;
    while (!converged && it < m && k->its < k->maxits) {
      vin.a = k->vv[it].data(); vout.a = k->vv[it + 1].data();
      mat_mult_any(k->A, &vin, &vout, transpose);
      /* classical Gram-Schmidt, one pass */
      double* vnew = k->vv[it + 1].data();
      for (int j = 0; j <= it; j++) { double s = 0.0; const double* vj = k->vv[j].data(); for (int i = 0; i < n; i++) s += vnew[i] * vj[i]; lhh[j] = s; }
      for (int j = 0; j <= it; j++) { const double* vj = k->vv[j].data(); double h = lhh[j]; for (int i = 0; i < n; i++) vnew[i] -= h * vj[i]; HH(j, it) = h; }
      double tt = 0.0; for (int i = 0; i < n; i++) tt += vnew[i] * vnew[i]; tt = sqrt(tt);
      if (tt > 0.0) for (int i = 0; i < n; i++) vnew[i] /= tt;
      HH(it + 1, it) = tt;
      /* apply previous rotations to the new column, then form the new one */
      for (int j = 0; j < it; j++) {
        double t = HH(j, it);
        HH(j, it) = cc[j] * t + ss[j] * HH(j + 1, it);
        HH(j + 1, it) = -ss[j] * t + cc[j] * HH(j + 1, it);
      }
      double t = sqrt(HH(it, it) * HH(it, it) + HH(it + 1, it) * HH(it + 1, it));
      if (t == 0.0) { converged = true; it++; k->its++; break; }
      cc[it] = HH(it, it) / t; ss[it] = HH(it + 1, it) / t;
      grs[it + 1] = -ss[it] * grs[it];
      grs[it] = cc[it] * grs[it];
      HH(it, it) = cc[it] * HH(it, it) + ss[it] * HH(it + 1, it);
      res = fabs(grs[it + 1]);
      it++; k->its++;
      k->rnorm = res;
      if (res <= ttol) converged = true;
    }
    /* build solution: solve upper triangular system and add the combination */
    if (it > 0) {
      std::vector<double> nrs(it);
      for (int kk = it - 1; kk >= 0; kk--) {
        double tsum = grs[kk];
        for (int j = kk + 1; j < it; j++) tsum -= HH(kk, j) * nrs[j];
        nrs[kk] = tsum / HH(kk, kk);
      }
      for (int j = 0; j < it; j++) { const double* vj = k->vv[j].data(); double c = nrs[j]; for (int i = 0; i < n; i++) x->a[i] += c * vj[i]; }
    }
  }
  return 0;
}
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x) { return gmres_solve(k, b, x, false); }
PetscErrorCode KSPSolveTranspose(KSP k, Vec b, Vec x) { return gmres_solve(k, b, x, true); }

/* ======================================================================= Tao (monitor-at-x0 stub) */
struct _p_Tao {
  Vec x = nullptr, xl = nullptr, xu = nullptr, g = nullptr;
  PetscErrorCode (*objgrad)(Tao, Vec, PetscReal*, Vec, void*) = nullptr; void* objgrad_ctx = nullptr;
  PetscErrorCode (*monitor)(Tao, void*) = nullptr; void* monitor_ctx = nullptr;
  int its = 0, maxits = 0; double f = 0, gnorm = 0, step = 1.0;
  TaoConvergedReason reason = TAO_CONTINUE_ITERATING;
};
PetscErrorCode TaoCreate(MPI_Comm, Tao* t) { *t = new _p_Tao(); return 0; }
PetscErrorCode TaoDestroy(Tao* t) { if ((*t)->g) VecDestroy(&(*t)->g); delete *t; *t = nullptr; return 0; }
PetscErrorCode TaoSetType(Tao, TaoType) { return 0; }
PetscErrorCode TaoSetMaximumIterations(Tao t, PetscInt m) { t->maxits = m; return 0; }
PetscErrorCode TaoSetTolerances(Tao, PetscReal, PetscReal, PetscReal) { return 0; }
PetscErrorCode TaoMonitorSet(Tao t, PetscErrorCode (*mon)(Tao, void*), void* ctx, PetscErrorCode (*)(void**)) { t->monitor = mon; t->monitor_ctx = ctx; return 0; }
PetscErrorCode TaoSetVariableBounds(Tao t, Vec xl, Vec xu) { t->xl = xl; t->xu = xu; return 0; }
PetscErrorCode TaoSetFromOptions(Tao) { return 0; }
PetscErrorCode TaoSetObjective(Tao, PetscErrorCode (*)(Tao, Vec, PetscReal*, void*), void*) { return 0; }
PetscErrorCode TaoSetGradient(Tao, Vec, PetscErrorCode (*)(Tao, Vec, Vec, void*), void*) { return 0; }
PetscErrorCode TaoSetObjectiveAndGradient(Tao t, Vec, PetscErrorCode (*f)(Tao, Vec, PetscReal*, Vec, void*), void* ctx) { t->objgrad = f; t->objgrad_ctx = ctx; return 0; }
PetscErrorCode TaoSetSolution(Tao t, Vec x) { t->x = x; return 0; }
PetscErrorCode TaoGetSolution(Tao t, Vec* x) { *x = t->x; return 0; }
PetscErrorCode TaoGetSolutionStatus(Tao t, PetscInt* its, PetscReal* f, PetscReal* gnorm, PetscReal* cnorm, PetscReal* xdiff, TaoConvergedReason* reason) {
  if (its) *its = t->its; if (f) *f = t->f; if (gnorm) *gnorm = t->gnorm; if (cnorm) *cnorm = 0.0; if (xdiff) *xdiff = t->step; if (reason) *reason = t->reason;
  return 0;
}
PetscErrorCode TaoSetConvergedReason(Tao t, TaoConvergedReason r) { t->reason = r; return 0; }
PetscErrorCode TaoSolve(Tao t) {
  /* Bounded solvers first project the start vector into [xl,xu] (TaoSolve_BNLS -> TaoBNKInitialize: VecMedian). */
  if (t->xl && t->xu) for (int i = 0; i < t->x->n; i++) t->x->a[i] = std::min(std::max(t->x->a[i], t->xl->a[i]), t->xu->a[i]);
  if (!t->g) VecDuplicate(t->x, &t->g);
  t->objgrad(t, t->x, &t->f, t->g, t->objgrad_ctx);
  /* projected gradient norm (VecBoundGradientProjection) */
  double s = 0.0;
  for (int i = 0; i < t->x->n; i++) {
    double gi = t->g->a[i];
    if (t->xl && t->x->a[i] <= t->xl->a[i] && gi > 0.0) gi = 0.0;
    if (t->xu && t->x->a[i] >= t->xu->a[i] && gi < 0.0) gi = 0.0;
    s += gi * gi;
  }
  t->gnorm = sqrt(s); t->its = 0; t->step = 1.0;
  if (t->monitor) t->monitor(t, t->monitor_ctx);
  if (MPI_COMM_WORLD->myidx == 0) printf("qshim: TAO BQNLS is not restated; stopped after the evaluation at the (projected) start vector.\n");
  return 0;
}
