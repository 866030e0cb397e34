// TEST INFRASTRUCTURE: stand-in for eigen_matrix_utils' <eigen_matrix_utils/eiquadprog.hpp> (un-vendored dependency of the reference,
// rosdyn.rosinstall:7-9, used only by Chain::computeLocalIk / computeWeigthedLocalIk, primitives_impl.h:1398-1468).
//   solve_quadprog(G, g0, CE, ce0, CI, ci0, x):  min 1/2 x^T G x + g0^T x   s.t.  CE^T x + ce0 = 0,  CI^T x + ci0 >= 0
// The reference calls it with no equalities and CI = [I, -I] (primitives_impl.h:783-793), i.e. a box; the stand-in returns the minimiser
// of exactly that problem (oracle/box_qp.h, checked against brute-force enumeration in tests/test_ik.py) and refuses anything else.
#pragma once
#include <stdexcept>
#include <vector>

#include "../../box_qp.h"
#include "../mini_eigen.h"
namespace Eigen
{
inline double solve_quadprog(MatrixXd& G, VectorXd& g0, const MatrixXd& CE, const VectorXd&, const MatrixXd& CI, const VectorXd& ci0, VectorXd& x)
{
  const int n = (int)g0.size();
  if (CE.cols() != 0 || CI.rows() != n || CI.cols() != 2 * n || n > OR_IK_MAXN) throw std::logic_error("solve_quadprog stand-in: box constraints only");
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 2 * n; k++)
    {
      const double want = (k == i) ? 1.0 : ((k == n + i) ? -1.0 : 0.0);
      if (CI(i, k) != want) throw std::logic_error("solve_quadprog stand-in: CI must be [I, -I]");
    }
  std::vector<double> H(n * n), f(n), lo(n), hi(n), sol(n);
  for (int i = 0; i < n; i++)
  {
    for (int k = 0; k < n; k++) H[i * n + k] = G(i, k);
    f[i] = g0(i);
    lo[i] = -ci0(i);
    hi[i] = ci0(n + i);
  }
  oracle_box_qp_impl(n, H.data(), f.data(), lo.data(), hi.data(), sol.data());
  x.resize(n);
  double v = 0.0;
  for (int i = 0; i < n; i++)
  {
    x(i) = sol[i];
    double s = 0.0;
    for (int k = 0; k < n; k++) s += H[i * n + k] * sol[k];
    v += sol[i] * (0.5 * s + f[i]);
  }
  return v;
}
}  // namespace Eigen
