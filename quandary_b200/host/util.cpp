/* util.cpp -- index maps between essential and full level spaces, vector file reader (reference: src/util.cpp:150-278,566-640) */
#include <cstdio>

#include "defs.hpp"

long mapEssToFull(long i, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential) {
  long id = 0, index = i;
  for (size_t iosc = 0; iosc + 1 < nlevels.size(); iosc++) {
    long postdim = 1, postdim_ess = 1;
    for (size_t j = iosc + 1; j < nlevels.size(); j++) { postdim *= (long)nlevels[j]; postdim_ess *= (long)nessential[j]; }
    const long iblock = index / postdim_ess;
    index = index % postdim_ess;
    id += iblock * postdim;
  }
  return id + index;
}

long mapFullToEss(long i, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential) {
  long id = 0, index = i;
  for (size_t iosc = 0; iosc < nlevels.size(); iosc++) {
    long postdim = 1, postdim_ess = 1;
    for (size_t j = iosc + 1; j < nlevels.size(); j++) { postdim *= (long)nlevels[j]; postdim_ess *= (long)nessential[j]; }
    const long iblock = index / postdim;
    index = index % postdim;
    if (iblock >= (long)nessential[iosc]) return -1;
    id += iblock * postdim_ess;
  }
  return id;
}

bool isEssential(long i, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential) {
  long index = i;
  for (size_t iosc = 0; iosc < nlevels.size(); iosc++) {
    long postdim = 1;
    for (size_t j = iosc + 1; j < nlevels.size(); j++) postdim *= (long)nlevels[j];
    if (index / postdim >= (long)nessential[iosc]) return false;
    index = index % postdim;
  }
  return true;
}

int read_vector(const char* filename, double* var, int dim, bool quietmode) {
  FILE* file = fopen(filename, "r");
  if (!file) {
    printf("ERROR: Can't open file %s\n", filename);
    return 0;
  }
  if (!quietmode) printf("Reading file %s\n", filename);
  for (int ix = 0; ix < dim; ix++) {
    double myval = 0.0;
    if (fscanf(file, "%lf", &myval) != 1) { printf("ERROR: EOF reached while reading %s (%d of %d values)\n", filename, ix, dim); break; }
    var[ix] = myval;
  }
  fclose(file);
  return 1;
}
