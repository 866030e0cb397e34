// TEST INFRASTRUCTURE: stand-in for eigen_matrix_utils' <eigen_matrix_utils/eiquadprog.hpp> (un-vendored dependency of the reference,
// used only by Chain::computeLocalIk, primitives_impl.h:1398-1468, which is off the hot path).  Declared so that the reference's header
// compiles; calling it reports that the QP solver is absent.
#pragma once
#include <stdexcept>

#include "../mini_eigen.h"
namespace Eigen
{
inline double solve_quadprog(MatrixXd&, VectorXd&, const MatrixXd&, const VectorXd&, const MatrixXd&, const VectorXd&, VectorXd&)
{
  throw std::logic_error("solve_quadprog: eigen_matrix_utils is not available in the checker build");
}
}  // namespace Eigen
