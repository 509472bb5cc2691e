/* qb_qubit.cu -- selection of the qubit-system sweep kernels (instantiated in qb_qubit_v*.cu, one variant per translation unit). */
#include "qb_qubit_args.h"

void qq_set_v0(QKernelSet* ks);   /* 3 qubits, 2 elements per thread (one warp per initial condition) */
void qq_set_v1(QKernelSet* ks);   /* 4 qubits, 4 elements per thread (two warps) */
void qq_set_v2(QKernelSet* ks);   /* 4 qubits, 2 elements per thread (four warps) */
void qq_set_v3(QKernelSet* ks);   /* 4 qubits, 1 element per thread (eight warps) */

/* Kernels exist for Lindblad systems of 3 and 4 qubits (dim 64 and 256).  epT: elements per thread (0 = default). */
void qq_pick(int Q, bool lind, int epT, QKernelSet* ks) {
  ks->fwd[0] = ks->fwd[1] = nullptr; ks->adj[0] = ks->adj[1] = nullptr; ks->app = nullptr;
  if (!lind) return;
  if (Q == 3) { qq_set_v0(ks); return; }
  if (Q == 4) {
    if (epT == 2) qq_set_v2(ks);
    else if (epT == 1) qq_set_v3(ks);
    else qq_set_v1(ks);
  }
}

size_t qq_smem(const QKernelSet& ks, bool adj, bool with_stage, unsigned recbytes, int nd) {
  return qq::qlayout(ks.nw, ks.vec_bytes, adj, adj, with_stage, recbytes, nd).total;
}
