/*
 * qb_dense.cu -- the tensor-core formulation of the operator for Hilbert dimension N = 64 (BASELINE config 5 names a
 * "tensor-core H*rho path"), built to be MEASURED against the stencil kernels, plus the fp64 tensor-pipe peak microbenchmark.
 *
 * y = M x with x = vec(rho):  -i [H(t), rho] as dense products on the fp64 tensor pipe (mma.sync.aligned.m8n8k4 f64 = SASS DMMA;
 * tcgen05 has no fp64 kind), the dissipator (diagonal + decay shift, src/mastereq.cpp:1713-1862 for three oscillators; dense
 * form docs/mkdocs/user_guide.md:28-58) as a short element-wise epilogue.  H(t) is assembled in shared memory from the level
 * digits:   H[i][i] = h_d(i),   H[i][i+e_k] = sqrt(i_k+1) (p_k + i q_k),   H[i][i-e_k] = sqrt(i_k) (p_k - i q_k),
 *           H[i][i-e_k+e_l] = J_kl sqrt(i_k (i_l+1)) (cos eta t + i sin eta t),   H[i][i+e_k-e_l] = its conjugate transpose entry
 * (the same coefficients the stencil kernels use, include/mastereq.hpp:316-531).
 * One CTA of 8 warps per state; warp w owns the 8 x 64 strip of rows 8w..8w+7 of K = H rho - rho H; 16 k-steps x 8 column tiles
 * x 8 real products = 1024 DMMA per warp and application (4.19 MFLOP per application against 0.57 MFLOP of the stencil).
 */
#include "qb_common.h"
#include "qb_kernels.cuh"

namespace qdense {

constexpr int N = 64;
constexpr int LD = 72;      /* padded row length (doubles): the four k-rows of a B fragment fall into two disjoint bank halves */

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dense_apply_kernel(const __grid_constant__ DevSys S, const double* __restrict__ row, const double2* __restrict__ xg,
                                                           double2* __restrict__ yg, int dimp, int reps) {
  extern __shared__ __align__(16) double sm[];
  double* Hr = sm;
  double* Hi = Hr + N * LD;
  double* Rr = Hi + N * LD;
  double* Ri = Rr + N * LD;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int Q = S.Q;
  const double2* x = xg + (size_t)blockIdx.x * dimp;
  double2* y = yg + (size_t)blockIdx.x * dimp;
  /* rho[r][c] = x[r + c*N] (vec is column-major, src/util.cpp:150-152) */
  for (int it = tid; it < N * N; it += 256) {
    const double2 v = x[it];
    const int r = it % N, c = it / N;
    Rr[r * LD + c] = v.x; Ri[r * LD + c] = v.y;
  }
  /* H(t) from the level digits */
  for (int idx = tid; idx < N * N; idx += 256) {
    const int i = idx / N, j = idx % N;
    int di[kMaxOsc], dj[kMaxOsc];
    { int a = i, b = j; for (int k = Q - 1; k >= 0; k--) { di[k] = a % S.nlev[k]; a /= S.nlev[k]; dj[k] = b % S.nlev[k]; b /= S.nlev[k]; } }
    int ndiff = 0, k0 = -1, k1 = -1;
    for (int k = 0; k < Q; k++) if (di[k] != dj[k]) { if (ndiff == 0) k0 = k; else k1 = k; ndiff++; }
    double hr = 0.0, hi = 0.0;
    if (ndiff == 0) hr = qbk::hdiag_dev(S, di);
    else if (ndiff == 1) {
      const double p = row[1 + k0], q = row[1 + Q + k0];
      if (dj[k0] == di[k0] + 1) { const double s = sqrt((double)(di[k0] + 1)); hr = s * p; hi = s * q; }
      else if (dj[k0] == di[k0] - 1) { const double s = sqrt((double)di[k0]); hr = s * p; hi = -s * q; }
    } else if (ndiff == 2) {
      const int id = k0 * Q - k0 * (k0 + 1) / 2 + (k1 - k0 - 1);     /* pair index in the order (0,1),(0,2),...,(1,2),... */
      const double J = S.Jkl[id];
      if (fabs(J) > 1e-10) {
        const int np = Q * (Q - 1) / 2;
        const double cs = row[1 + 2 * Q + id], sn = row[1 + 2 * Q + np + id];
        /* j = i - e_k0 + e_k1 */
        if (dj[k0] == di[k0] - 1 && dj[k1] == di[k1] + 1) { const double c = J * sqrt((double)(di[k0] * (di[k1] + 1))); hr = c * cs; hi = c * sn; }
        /* j = i + e_k0 - e_k1 */
        else if (dj[k0] == di[k0] + 1 && dj[k1] == di[k1] - 1) { const double c = J * sqrt((double)((di[k0] + 1) * di[k1])); hr = c * cs; hi = -c * sn; }
      }
    }
    Hr[i * LD + j] = hr; Hi[i * LD + j] = hi;
  }
  __syncthreads();
  const int r = lane >> 2, c = lane & 3;
  const int arow = (8 * w + r) * LD;
  for (int rep = 0; rep < reps; rep++) {
    double Kr[8][2], Ki[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) { Kr[nt][0] = Kr[nt][1] = 0.0; Ki[nt][0] = Ki[nt][1] = 0.0; }
#pragma unroll 4
    for (int kk = 0; kk < 16; kk++) {
      const int k0 = 4 * kk;
      const double aHr = Hr[arow + k0 + c], aHi = Hi[arow + k0 + c], aRr = Rr[arow + k0 + c], aRi = Ri[arow + k0 + c];
      const int brow = (k0 + c) * LD + r;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const double bRr = Rr[brow + 8 * nt], bRi = Ri[brow + 8 * nt], bHr = Hr[brow + 8 * nt], bHi = Hi[brow + 8 * nt];
        /* K = H rho - rho H */
        dmma(Kr[nt], aHr, bRr); dmma(Kr[nt], -aHi, bRi);
        dmma(Ki[nt], aHr, bRi); dmma(Ki[nt], aHi, bRr);
        dmma(Kr[nt], -aRr, bHr); dmma(Kr[nt], aRi, bHi);
        dmma(Ki[nt], -aRr, bHi); dmma(Ki[nt], -aRi, bHr);
      }
    }
    /* y = -i K + dissipator.  Accumulator layout of m8n8k4: row = lane/4, columns 2*(lane%4) + {0,1} */
    const int i = 8 * w + r;
    int di[kMaxOsc];
    { int a = i; for (int k = Q - 1; k >= 0; k--) { di[k] = a % S.nlev[k]; a /= S.nlev[k]; } }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int ip = 8 * nt + 2 * c + h;
        int dp[kMaxOsc];
        { int a = ip; for (int k = Q - 1; k >= 0; k--) { dp[k] = a % S.nlev[k]; a /= S.nlev[k]; } }
        double l1diag = 0.0, l2 = 0.0;
        for (int k = 0; k < Q; k++) {
          l1diag += -S.decay[k] / 2.0 * (di[k] + dp[k]);
          l2 += S.dephase[k] * (di[k] * dp[k] - 1. / 2. * (di[k] * di[k] + dp[k] * dp[k]));
        }
        const double dB = l2 + l1diag;
        double dre = dB * Rr[i * LD + ip], dim = dB * Ri[i * LD + ip];
        for (int k = 0; k < Q; k++) {
          const double dec = S.decay[k];
          if (fabs(dec) > 1e-12 && di[k] < S.nlev[k] - 1 && dp[k] < S.nlev[k] - 1) {
            const double l1 = dec * sqrt((double)((di[k] + 1) * (dp[k] + 1)));
            dre = fma(l1, Rr[(i + S.stride[k]) * LD + ip + S.stride[k]], dre);
            dim = fma(l1, Ri[(i + S.stride[k]) * LD + ip + S.stride[k]], dim);
          }
        }
        y[i + ip * N] = make_double2(dre + Ki[nt][h], dim - Kr[nt][h]);
      }
    }
  }
}

/* fp64 tensor-pipe peak: 8 independent accumulator pairs per thread, 512 flops per DMMA */
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a0, double b0) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = i * 0.5; }
  const double a = a0 + threadIdx.x * 1e-12, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace qdense

extern "C" {
bool qb_dense_supported(const DevSys& S) { return S.lind != 0 && S.N == qdense::N && S.dim == qdense::N * qdense::N; }

cudaError_t qb_dense_apply_launch(const DevSys& S, const double* d_row, const double2* x, double2* y, int dimp, int nbatch, int reps, cudaStream_t stream) {
  const size_t smem = sizeof(double) * 4 * qdense::N * qdense::LD;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute((const void*)qdense::dense_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  qdense::dense_apply_kernel<<<nbatch, 256, smem, stream>>>(S, d_row, x, y, dimp, reps);
  return cudaGetLastError();
}

cudaError_t qb_dmma_peak(int device, double* tflops) {
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return err;
  cudaDeviceProp prop;
  if ((err = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return err;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 13;
  double* d_out = nullptr;
  if ((err = cudaMalloc(&d_out, sizeof(double) * blocks * threads)) != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    qdense::dmma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    if ((err = cudaEventSynchronize(e1)) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = best;
  return err;
}
}  /* extern "C" (declared inside engine.cu's extern "C" block) */
