/* Kernel instantiations for systems with 3 oscillator(s); see qb_kernels.cuh. */
#include "qb_kernels.cuh"
template void qb_pick<3>(bool, bool, bool, int, KernelSet*);
