/* Kernel instantiations for systems with 5 oscillator(s); see qb_kernels.cuh. */
#include "qb_kernels.cuh"
template void qb_pick<5>(bool, bool, bool, int, KernelSet*);
