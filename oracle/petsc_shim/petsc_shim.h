/*
 * petsc_shim.h -- single-node PETSc/MPI/TAO stand-in used ONLY to compile the UNMODIFIED
 * reference sources (/root/reference/src/*.cpp) into oracle/_ref/ for parity checking and
 * for the CPU baseline.  TEST INFRASTRUCTURE: nothing in the product path (quandary_b200/)
 * includes or links this.
 *
 * Why it exists: the reference needs PETSc + MPI, which are not installed and cannot be
 * fetched (no network).  The reference's hot path uses a small, enumerable part of PETSc
 * (SURVEY.md section 8c "Option A"); this header declares exactly that surface with
 * sequential semantics.  "MPI" is implemented over fork()+shared memory so that
 * `QSHIM_NP=8 quandary cfg` distributes initial conditions over 8 processes the way
 * `mpirun -np 8 quandary cfg` does (src/main.cpp:133-175).
 *
 * Third-party arithmetic restated here (absent from /root/reference, PETSc version unpinned,
 * README mentions 3.22.1): KSPGMRES (restart 30, classical Gram-Schmidt without refinement,
 * Givens-rotation residual recurrence, KSPConvergedDefault test) -- see shim.cpp.
 * TAO BQNLS is NOT restated: TaoSolve() projects the start vector into the bounds, evaluates
 * objective+gradient once and calls the monitor with iter=0 (enough to reproduce row 0 of
 * tests/regression/cnot/base/optim_history.dat).
 */
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <sstream>
#include <iostream>
#include <fstream>

#define PETSC_VERSION_MAJOR 3
#define PETSC_VERSION_MINOR 22

typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;

#define PETSC_DEFAULT (-2)
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_SUCCESS 0
#define CHKERRQ(ierr) do { if (ierr) return ierr; } while (0)
#define PetscCall(x) do { PetscErrorCode _e = (x); if (_e) return _e; } while (0)

/* ------------------------------------------------------------------ MPI */
struct qshim_comm;
typedef qshim_comm* MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
extern MPI_Comm MPI_COMM_WORLD;
extern MPI_Comm MPI_COMM_SELF;
extern MPI_Comm PETSC_COMM_WORLD;
#define PETSC_COMM_SELF MPI_COMM_SELF
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPIU_INT MPI_INT
#define MPIU_SCALAR MPI_DOUBLE
#define MPIU_REAL MPI_DOUBLE
#define MPI_SUM 1
#define MPI_SUCCESS 0

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize();
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sdt, void* recvbuf, int recvcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime();

/* ------------------------------------------------------------------ objects */
struct _p_Vec;  typedef _p_Vec* Vec;
struct _p_Mat;  typedef _p_Mat* Mat;
struct _p_IS;   typedef _p_IS* IS;
struct _p_KSP;  typedef _p_KSP* KSP;
struct _p_PC;   typedef _p_PC* PC;
struct _p_Tao;  typedef _p_Tao* Tao;
struct _p_PetscViewer; typedef _p_PetscViewer* PetscViewer;
struct _p_VecScatter;  typedef _p_VecScatter* VecScatter;

typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { SCATTER_FORWARD = 0, SCATTER_REVERSE = 1 } ScatterMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX, MAT_IGNORE_MATRIX, MAT_INPLACE_MATRIX } MatReuse;
typedef enum { MAT_DO_NOT_COPY_VALUES, MAT_COPY_VALUES, MAT_SHARE_NONZERO_PATTERN } MatDuplicateOption;
typedef enum { DIFFERENT_NONZERO_PATTERN, SUBSET_NONZERO_PATTERN, SAME_NONZERO_PATTERN, UNKNOWN_NONZERO_PATTERN } MatStructure;
typedef enum { MAT_NEW_NONZERO_ALLOCATION_ERR = 19, MAT_SYMMETRIC = 1 } MatOption;
typedef enum { MATOP_MULT = 3, MATOP_MULT_TRANSPOSE = 5 } MatOperation;
typedef const char* MatType;
typedef const char* VecType;
typedef const char* KSPType;
typedef const char* PCType;
typedef const char* TaoType;
typedef const char* PetscViewerType;
#define MATMPIAIJ "mpiaij"
#define MATSEQAIJ "seqaij"
#define MATAIJ "aij"
#define MATSAME "same"
#define MATSHELL "shell"
#define MATSEQDENSE "seqdense"
#define KSPGMRES "gmres"
#define PCNONE "none"
#define TAOBQNLS "bqnls"
#define TAOBLMVM "blmvm"
#define PETSCVIEWERASCII "ascii"
#define PETSCVIEWERBINARY "binary"
typedef enum { PETSC_VIEWER_DEFAULT, PETSC_VIEWER_ASCII_MATLAB, PETSC_VIEWER_ASCII_DENSE } PetscViewerFormat;
typedef enum { FILE_MODE_READ, FILE_MODE_WRITE } PetscFileMode;
extern PetscViewer PETSC_VIEWER_STDOUT_WORLD;
extern PetscViewer PETSC_VIEWER_STDOUT_SELF;

typedef enum {
  TAO_CONVERGED_GATOL = 3, TAO_CONVERGED_GRTOL = 4, TAO_CONVERGED_GTTOL = 5, TAO_CONVERGED_STEPTOL = 6,
  TAO_CONVERGED_MINF = 7, TAO_CONVERGED_USER = 8, TAO_DIVERGED_MAXITS = -2, TAO_DIVERGED_NAN = -4,
  TAO_DIVERGED_MAXFCN = -5, TAO_DIVERGED_LS_FAILURE = -6, TAO_DIVERGED_TR_REDUCTION = -7, TAO_DIVERGED_USER = -8,
  TAO_CONTINUE_ITERATING = 0
} TaoConvergedReason;

/* ------------------------------------------------------------------ sys */
PetscErrorCode PetscInitialize(int* argc, char*** argv, const char* file, const char* help);
PetscErrorCode PetscFinalize();
PetscErrorCode PetscOptionsSetValue(void* options, const char* name, const char* value);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char* fmt, ...);
template <typename T> inline PetscErrorCode PetscMalloc1(size_t n, T** p) { *p = (T*)malloc((n ? n : 1) * sizeof(T)); return 0; }
template <typename T> inline PetscErrorCode PetscFreeImpl(T*& p) { free((void*)p); p = nullptr; return 0; }
#define PetscFree(p) PetscFreeImpl(p)
inline PetscReal PetscAbsReal(PetscReal a) { return fabs(a); }

PetscErrorCode PetscViewerPushFormat(PetscViewer v, PetscViewerFormat f);
PetscErrorCode PetscViewerPopFormat(PetscViewer v);
PetscErrorCode PetscViewerCreate(MPI_Comm comm, PetscViewer* v);
PetscErrorCode PetscViewerSetType(PetscViewer v, PetscViewerType t);
PetscErrorCode PetscViewerFileSetMode(PetscViewer v, PetscFileMode m);
PetscErrorCode PetscViewerFileSetName(PetscViewer v, const char* name);
PetscErrorCode PetscViewerDestroy(PetscViewer* v);
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char* name, PetscFileMode m, PetscViewer* v);

/* ------------------------------------------------------------------ IS */
PetscErrorCode ISCreateStride(MPI_Comm comm, PetscInt n, PetscInt first, PetscInt step, IS* is);
PetscErrorCode ISDestroy(IS* is);

/* ------------------------------------------------------------------ Vec */
PetscErrorCode VecCreate(MPI_Comm comm, Vec* v);
PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec* v);
PetscErrorCode VecSetSizes(Vec v, PetscInt nlocal, PetscInt nglobal);
PetscErrorCode VecSetFromOptions(Vec v);
PetscErrorCode VecSetUp(Vec v);
PetscErrorCode VecDuplicate(Vec v, Vec* w);
PetscErrorCode VecDestroy(Vec* v);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecAYPX(Vec y, PetscScalar b, Vec x);
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar alpha, PetscScalar beta, PetscScalar gamma, Vec x, Vec y);
PetscErrorCode VecScale(Vec v, PetscScalar a);
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal* nrm);
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar* val);
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar* val);
PetscErrorCode VecGetArray(Vec v, PetscScalar** a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar** a);
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar** a);
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar** a);
PetscErrorCode VecGetValues(Vec v, PetscInt ni, const PetscInt ix[], PetscScalar y[]);
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode m);
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar y, InsertMode m);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecGetSize(Vec v, PetscInt* n);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n);
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt* lo, PetscInt* hi);
PetscErrorCode VecGetSubVector(Vec v, IS is, Vec* sub);
PetscErrorCode VecRestoreSubVector(Vec v, IS is, Vec* sub);
PetscErrorCode VecISCopy(Vec vfull, IS is, ScatterMode mode, Vec vreduced);
PetscErrorCode VecView(Vec v, PetscViewer viewer);
PetscErrorCode VecScatterCreateToZero(Vec v, VecScatter* s, Vec* out);
PetscErrorCode VecScatterBegin(VecScatter s, Vec x, Vec y, InsertMode m, ScatterMode sm);
PetscErrorCode VecScatterEnd(VecScatter s, Vec x, Vec y, InsertMode m, ScatterMode sm);
PetscErrorCode VecScatterDestroy(VecScatter* s);

/* ------------------------------------------------------------------ Mat */
PetscErrorCode MatCreate(MPI_Comm comm, Mat* A);
PetscErrorCode MatCreateSeqDense(MPI_Comm comm, PetscInt m, PetscInt n, PetscScalar* data, Mat* A);
PetscErrorCode MatCreateShell(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, void* ctx, Mat* A);
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void));
PetscErrorCode MatShellGetContext(Mat A, void* ctx);
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N);
PetscErrorCode MatSetType(Mat A, MatType t);
PetscErrorCode MatSetFromOptions(Mat A);
PetscErrorCode MatSetUp(Mat A);
PetscErrorCode MatSetOptionsPrefix(Mat A, const char* prefix);
PetscErrorCode MatSetOption(Mat A, MatOption op, PetscBool flg);
PetscErrorCode MatMPIAIJSetPreallocation(Mat A, PetscInt dnz, const PetscInt* dnnz, PetscInt onz, const PetscInt* onnz);
PetscErrorCode MatSeqAIJSetPreallocation(Mat A, PetscInt nz, const PetscInt* nnz);
PetscErrorCode MatDestroy(Mat* A);
PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode m);
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode mode);
PetscErrorCode MatGetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], PetscScalar v[]);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatGetRow(Mat A, PetscInt row, PetscInt* ncols, const PetscInt** cols, const PetscScalar** vals);
PetscErrorCode MatRestoreRow(Mat A, PetscInt row, PetscInt* ncols, const PetscInt** cols, const PetscScalar** vals);
PetscErrorCode MatGetSize(Mat A, PetscInt* m, PetscInt* n);
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt* lo, PetscInt* hi);
PetscErrorCode MatCreateVecs(Mat A, Vec* right, Vec* left);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z);
PetscErrorCode MatMultTransposeAdd(Mat A, Vec x, Vec y, Vec z);
PetscErrorCode MatScale(Mat A, PetscScalar a);
PetscErrorCode MatShift(Mat A, PetscScalar a);
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure s);
PetscErrorCode MatNorm(Mat A, NormType t, PetscReal* nrm);
PetscErrorCode MatTranspose(Mat A, MatReuse reuse, Mat* B);
PetscErrorCode MatTransposeMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat* C);
PetscErrorCode MatConvert(Mat A, MatType t, MatReuse reuse, Mat* B);
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat* B);
PetscErrorCode MatIsTranspose(Mat A, Mat B, PetscReal tol, PetscBool* flg);
PetscErrorCode MatIsSymmetric(Mat A, PetscReal tol, PetscBool* flg);
PetscErrorCode MatView(Mat A, PetscViewer viewer);
PetscErrorCode MatLoad(Mat A, PetscViewer viewer);

/* ------------------------------------------------------------------ KSP */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP* ksp);
PetscErrorCode KSPDestroy(KSP* ksp);
PetscErrorCode KSPGetPC(KSP ksp, PC* pc);
PetscErrorCode PCSetType(PC pc, PCType t);
PetscErrorCode KSPSetType(KSP ksp, KSPType t);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPSolveTranspose(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal* rnorm);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt* its);

/* ------------------------------------------------------------------ Tao */
PetscErrorCode TaoCreate(MPI_Comm comm, Tao* tao);
PetscErrorCode TaoDestroy(Tao* tao);
PetscErrorCode TaoSetType(Tao tao, TaoType t);
PetscErrorCode TaoSetMaximumIterations(Tao tao, PetscInt maxits);
PetscErrorCode TaoSetTolerances(Tao tao, PetscReal gatol, PetscReal grtol, PetscReal gttol);
PetscErrorCode TaoMonitorSet(Tao tao, PetscErrorCode (*mon)(Tao, void*), void* ctx, PetscErrorCode (*destroy)(void**));
PetscErrorCode TaoSetVariableBounds(Tao tao, Vec xl, Vec xu);
PetscErrorCode TaoSetFromOptions(Tao tao);
PetscErrorCode TaoSetObjective(Tao tao, PetscErrorCode (*f)(Tao, Vec, PetscReal*, void*), void* ctx);
PetscErrorCode TaoSetGradient(Tao tao, Vec g, PetscErrorCode (*f)(Tao, Vec, Vec, void*), void* ctx);
PetscErrorCode TaoSetObjectiveAndGradient(Tao tao, Vec g, PetscErrorCode (*f)(Tao, Vec, PetscReal*, Vec, void*), void* ctx);
PetscErrorCode TaoSetSolution(Tao tao, Vec x);
PetscErrorCode TaoGetSolution(Tao tao, Vec* x);
PetscErrorCode TaoSolve(Tao tao);
PetscErrorCode TaoGetSolutionStatus(Tao tao, PetscInt* its, PetscReal* f, PetscReal* gnorm, PetscReal* cnorm, PetscReal* xdiff, TaoConvergedReason* reason);
PetscErrorCode TaoSetConvergedReason(Tao tao, TaoConvergedReason reason);

/* ------------------------------------------------------------------ shim statistics (not PETSc) */
extern "C" {
long long qshim_matmult_count();       /* number of shell MatMult + MatMultTranspose applications so far */
void qshim_reset_counters();
}
