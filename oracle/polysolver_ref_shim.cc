// Shim that lets the reference's OWN PolySolver::getCoefficients (plain arithmetic, lines 50-149 of
// Modules/Mapping/PolySolver.cc) be compiled without Ceres/Eigen: oracle/Makefile extracts those
// lines from /root/reference into oracle/_ref/polysolver_getcoefficients.inc (git-ignored, never
// committed) and this file includes them.  TEST INFRASTRUCTURE ONLY: used to pin the oracle's
// polynomial coefficients and to generate tests/golden/polysolver_ref.npz.
#include <cmath>
using std::pow;
namespace defSLAM {
struct PolySolver {
  static void getCoefficients(double a, double b, double c, double d, double t1, double t2, double e1, double e2,
                              double x1, double y1, double x2, double y2, int i, double *eq);
};
#include "_ref/polysolver_getcoefficients.inc"
}  // namespace defSLAM

extern "C" void ref_polysolver_coefficients(double a, double b, double c, double d, double t1, double t2, double e1,
                                            double e2, double x1, double y1, double x2, double y2, double *eq1,
                                            double *eq2) {
  defSLAM::PolySolver::getCoefficients(a, b, c, d, t1, t2, e1, e2, x1, y1, x2, y2, 0, eq1);
  defSLAM::PolySolver::getCoefficients(a, b, c, d, t1, t2, e1, e2, x1, y1, x2, y2, 1, eq2);
}
