/* Kernel instantiations for systems with 4 oscillator(s); see qb_kernels.cuh. */
#include "qb_kernels.cuh"
template void qb_pick<4>(bool, bool, bool, int, KernelSet*);
