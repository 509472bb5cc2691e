/* Kernel instantiations for systems with 1 oscillator(s); see qb_kernels.cuh. */
#include "qb_kernels.cuh"
template void qb_pick<1>(bool, bool, bool, int, KernelSet*);
