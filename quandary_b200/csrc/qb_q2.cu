/* Kernel instantiations for systems with 2 oscillator(s); see qb_kernels.cuh. */
#include "qb_kernels.cuh"
template void qb_pick<2>(bool, bool, bool, int, KernelSet*);
