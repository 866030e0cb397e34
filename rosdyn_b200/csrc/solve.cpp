// solve.cpp -- the step right after the hot path in identification: solve the normal equations
// G pi = b accumulated by rdb_regressor_gram_batch.  The reference holds no code for it (the consumer, rosdyn_identification, is an
// external package: reference README.md:15).  Phi is rank deficient in the standard parameters (only the base parameters are
// identifiable), so the solve is the minimum-norm least-squares solution through a symmetric eigen-decomposition of G
// (cyclic Jacobi, fp64): pi = V_r diag(1/lambda_r) V_r^T b over the eigenvalues lambda > rel_tol * lambda_max.
// Host code on purpose: G is (10 nJ + Pc)^2 <= a few hundred squared, microseconds of work next to billions of samples.
#include <algorithm>
#include <cmath>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/rosdyn_b200.h"

namespace rdb
{
rdb_status set_error(rdb_status s, const std::string& what);

// eigen-decomposition of the symmetric n x n matrix a (column-major, destroyed): a = V diag(w) V^T, V column-major
static void jacobi_eig(int n, std::vector<double>& a, std::vector<double>& V, std::vector<double>& w)
{
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  auto A = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };
  for (int sweep = 0; sweep < 100; sweep++)
  {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) (i == j ? diag : off) += A(i, j) * A(i, j);
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++)
      {
        const double apq = A(p, q);
        if (apq == 0.0) continue;
        const double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++)
        {
          const double akp = A(k, p), akq = A(k, q);
          A(k, p) = c * akp - s * akq;
          A(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++)
        {
          const double apk = A(p, k), aqk = A(q, k);
          A(p, k) = c * apk - s * aqk;
          A(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++)
        {
          const double vkp = V[(size_t)p * n + k], vkq = V[(size_t)q * n + k];
          V[(size_t)p * n + k] = c * vkp - s * vkq;
          V[(size_t)q * n + k] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A(i, i);
}
}  // namespace rdb

extern "C" rdb_status rdb_normal_equations_solve(int32_t P, const double* gram, const double* rhs, double tau_sq, double rel_tol,
                                                 double* parameters, double* eigenvalues, int32_t* rank, double* residual_sq)
{
  using namespace rdb;
  if (P <= 0 || !gram || !rhs || !parameters) return set_error(RDB_ERR_INVALID_ARG, "normal_equations_solve: bad argument");
  if (!(rel_tol > 0)) rel_tol = 1e-10;
  std::vector<double> a((size_t)P * P), V, w;
  for (int j = 0; j < P; j++)
    for (int i = 0; i < P; i++) a[(size_t)j * P + i] = 0.5 * (gram[(size_t)j * P + i] + gram[(size_t)i * P + j]);
  const std::vector<double> G(a);
  jacobi_eig(P, a, V, w);
  std::vector<int> order(P);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int x, int y) { return w[x] > w[y]; });
  const double wmax = std::max(w[order[0]], 0.0);
  int r = 0;
  std::fill(parameters, parameters + P, 0.0);
  for (int k = 0; k < P; k++)
  {
    const int e = order[k];
    if (eigenvalues) eigenvalues[k] = w[e];
    if (!(w[e] > rel_tol * wmax) || wmax == 0.0) continue;
    r++;
    double vb = 0.0;
    for (int i = 0; i < P; i++) vb += V[(size_t)e * P + i] * rhs[i];
    const double coef = vb / w[e];
    for (int i = 0; i < P; i++) parameters[i] += coef * V[(size_t)e * P + i];
  }
  if (rank) *rank = r;
  if (residual_sq)
  {
    // || Phi pi - tau ||^2 = tau^T tau - 2 pi^T b + pi^T G pi
    double pb = 0.0, pGp = 0.0;
    for (int j = 0; j < P; j++)
    {
      pb += parameters[j] * rhs[j];
      double s = 0.0;
      for (int i = 0; i < P; i++) s += G[(size_t)j * P + i] * parameters[i];
      pGp += parameters[j] * s;
    }
    *residual_sq = tau_sq - 2.0 * pb + pGp;
  }
  return RDB_OK;
}

// Chain::getMultiplicity (primitives_impl.h:1470-1517): the multi-turn images of q inside the joint limits
extern "C" rdb_status rdb_multiplicity(int32_t n, const int32_t* joint_type_of_input, const double* q, const double* q_min, const double* q_max,
                                       double* out, int64_t capacity, int64_t* count)
{
  if (n < 0 || !count || (n > 0 && (!joint_type_of_input || !q || !q_min || !q_max)))
    return rdb::set_error(RDB_ERR_INVALID_ARG, "multiplicity: null or negative argument");
  *count = 0;
  const double two_pi = 2.0 * 3.14159265358979323846;  // 2*M_PI
  // The reference loops until the limit is passed; with its "no limit" defaults (+-1e10, PI.h:92-93) that is 1.6e9 turns per joint and an
  // exponential number of vectors.  Refused instead of reproduced: non-finite values (a NaN q never passes a limit), a revolute joint whose
  // images number more than 2001 (limits wider than 10^3 turns, or q that far outside them), more than 2^22 vectors in total.
  constexpr double MAX_TURNS = 1.0e3;
  constexpr double MAX_TOTAL = 4194304.0;
  double total = 1.0;
  for (int i = 0; i < n; i++)
  {
    if (!std::isfinite(q[i])) return rdb::set_error(RDB_ERR_INVALID_ARG, "multiplicity: q is not finite");
    if (joint_type_of_input[i] != RDB_JOINT_REVOLUTE) continue;
    if (!std::isfinite(q_min[i]) || !std::isfinite(q_max[i])) return rdb::set_error(RDB_ERR_INVALID_ARG, "multiplicity: limits are not finite");
    const double up = std::max(0.0, (q_max[i] - q[i]) / two_pi), down = std::max(0.0, (q[i] - q_min[i]) / two_pi);
    if (!(up <= MAX_TURNS) || !(down <= MAX_TURNS))
      return rdb::set_error(RDB_ERR_INVALID_ARG, "multiplicity: more than 1000 turns between q and a joint limit");
    total *= 1.0 + std::floor(up) + std::floor(down);
    if (total > MAX_TOTAL) return rdb::set_error(RDB_ERR_INVALID_ARG, "multiplicity: more than 2^22 joint vectors");
  }
  try
  {
    std::vector<std::vector<double>> ax((size_t)n);
    for (int i = 0; i < n; i++)
    {
      ax[i].push_back(q[i]);
      if (joint_type_of_input[i] != RDB_JOINT_REVOLUTE) continue;
      for (double t = q[i] + two_pi; !(t > q_max[i]); t += two_pi) ax[i].push_back(t);  // while (true) { tmp += 2 pi; if (tmp > max) break; ... }
      for (double t = q[i] - two_pi; !(t < q_min[i]); t -= two_pi) ax[i].push_back(t);
    }
    std::vector<std::vector<double>> all;
    all.emplace_back(q, q + n);
    for (int i = 0; i < n; i++)
    {
      const size_t have = all.size();
      for (size_t is = 1; is < ax[i].size(); is++)
        for (size_t im = 0; im < have; im++)
        {
          std::vector<double> v = all[im];
          v[i] = ax[i][is];
          all.push_back(std::move(v));
        }
    }
    *count = (int64_t)all.size();
    if (!out || capacity < *count) return rdb::set_error(RDB_ERR_INVALID_ARG, "multiplicity: capacity too small (count is set)");
    for (size_t k = 0; k < all.size(); k++)
      for (int i = 0; i < n; i++) out[k * (size_t)n + i] = all[k][i];
  }
  catch (const std::bad_alloc&)  // nothing may cross the extern "C" boundary
  {
    return rdb::set_error(RDB_ERR_ALLOC, "multiplicity: out of host memory");
  }
  return RDB_OK;
}
