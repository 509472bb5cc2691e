/* defs.hpp -- enums shared by the host-side classes; same names and values as the reference's include/defs.hpp. */
#pragma once
#include <string>
#include <vector>

enum class LindbladType { NONE, DECAY, DEPHASE, BOTH };
enum class LinearSolverType { GMRES, NEUMANN };
enum class RunType { SIMULATION, GRADIENT, OPTIMIZATION, EVALCONTROLS, NONE };
enum class ControlType { NONE, BSPLINE, BSPLINEAMP, STEP, BSPLINE0 };
enum class InitialConditionType { FROMFILE, PURE, ENSEMBLE, DIAGONAL, BASIS, THREESTATES, NPLUSONE, PERFORMANCE };
enum class TargetType { GATE, PURE, FROMFILE };
enum class ObjectiveType { JFROBENIUS, JTRACE, JMEASURE };

/* The reference passes PETSc `Vec`s; here a state / design vector is a contiguous host array in the same layout
 * (state: [Re(dim) | Im(dim)], vec(rho) column-major; src/mastereq.cpp:1322-1323, src/util.cpp:150-152). */
typedef std::vector<double> Vec;

/* helpers with the reference's names (src/util.cpp:150-278) */
inline long getVecID(long row, long col, long dim) { return row + col * dim; }
template <typename T>
inline void copyLast(std::vector<T>& v, size_t n) {
  if (v.empty()) return;
  T last = v.back();
  while (v.size() < n) v.push_back(last);
}
long mapEssToFull(long i, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential);
long mapFullToEss(long i, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential);
bool isEssential(long i, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential);
int read_vector(const char* filename, double* var, int dim, bool quietmode);
