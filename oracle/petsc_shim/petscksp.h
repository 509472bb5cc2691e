/* Forwarding header: the single-process PETSc/MPI shim lives in petsc_shim.h (test infrastructure, see oracle/README.md). */
#pragma once
#include "petsc_shim.h"
