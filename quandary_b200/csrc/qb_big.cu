/* qb_big.cu -- host-side helpers of the large-state kernels */
#include "qb_big.cuh"
size_t qbig_smem(const BKernelSet& ks, int mcap, int L, bool adj, int nd) { return qbig::blayout(ks.nred, mcap, L, ks.dimp, adj, nd).total; }
