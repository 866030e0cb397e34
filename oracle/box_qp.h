/* TEST INFRASTRUCTURE (oracle): exact box-constrained QP used by the local-IK restatement (rosdyn_oracle.c) and by the stand-in of the
 * un-vendored eigen_matrix_utils QP solver in the reference build (shim/eigen_matrix_utils/eiquadprog.hpp). */
#ifndef ROSDYN_ORACLE_BOX_QP_H
#define ROSDYN_ORACLE_BOX_QP_H
#include <math.h>

/* The QP of PI.h:1421-1427: min 1/2 x^T H x + f^T x s.t. lo <= x <= hi (CI = [I, -I], ci0 = [sol - q_min; q_max - sol], no equalities).
 * Eigen::solve_quadprog lives in the un-vendored eigen_matrix_utils (rosdyn.rosinstall:7-9); what it returns is the minimiser, which a
 * primal active-set method finds as well.  Directions in which H is numerically singular are left where they are. */
#define OR_IK_MAXN 8
static inline void oracle_box_qp_impl(int n, const double* H, const double* f, const double* lo, const double* hi, double* x)
{
  int state[OR_IK_MAXN]; /* 0 free, 1 at lo, 2 at hi, 3 pinned */
  double hmx = 0.0, fmx = 0.0;
  for (int i = 0; i < n; i++)
  {
    hmx = fmax(hmx, H[i * n + i]);
    fmx = fmax(fmx, fabs(f[i]));
    x[i] = 0.0;
    state[i] = 0;
    if (lo[i] >= hi[i]) { x[i] = lo[i]; state[i] = 3; }
    else if (x[i] <= lo[i]) { x[i] = lo[i]; state[i] = 1; }
    else if (x[i] >= hi[i]) { x[i] = hi[i]; state[i] = 2; }
  }
  const double ptol = 1e-13 * hmx, gtol = 1e-12 * (fmx + hmx);
  for (int it = 0; it < 6 * OR_IK_MAXN + 8; it++)
  {
    double g[OR_IK_MAXN], d[OR_IK_MAXN], L[OR_IK_MAXN][OR_IK_MAXN], y[OR_IK_MAXN];
    int idx[OR_IK_MAXN], ok[OR_IK_MAXN], m = 0;
    for (int i = 0; i < n; i++)
    {
      double s = f[i];
      for (int k = 0; k < n; k++) s += H[i * n + k] * x[k];
      g[i] = s;
      d[i] = 0.0;
      if (state[i] == 0) idx[m++] = i;
    }
    for (int a = 0; a < m; a++)
      for (int b = 0; b <= a; b++)
      {
        double s = H[idx[a] * n + idx[b]];
        for (int k = 0; k < b; k++) s -= L[a][k] * L[b][k];
        if (a == b)
        {
          ok[a] = s > ptol;
          L[a][a] = ok[a] ? sqrt(s) : 1.0;
          if (!ok[a])
            for (int k = 0; k < a; k++) L[a][k] = 0.0;
        }
        else
          L[a][b] = ok[b] ? s / L[b][b] : 0.0;
      }
    for (int a = 0; a < m; a++)
    {
      double s = ok[a] ? -g[idx[a]] : 0.0;
      for (int k = 0; k < a; k++) s -= L[a][k] * y[k];
      y[a] = s / L[a][a];
    }
    for (int a = m - 1; a >= 0; a--)
    {
      double s = y[a];
      for (int k = a + 1; k < m; k++) s -= L[k][a] * d[idx[k]];
      d[idx[a]] = ok[a] ? s / L[a][a] : 0.0;
    }
    double alpha = 1.0;
    int blocking = -1, bstate = 0;
    for (int a = 0; a < m; a++)
    {
      int i = idx[a];
      if (d[i] > 0.0 && x[i] + d[i] > hi[i])
      {
        double s = (hi[i] - x[i]) / d[i];
        if (s < alpha) { alpha = s; blocking = i; bstate = 2; }
      }
      else if (d[i] < 0.0 && x[i] + d[i] < lo[i])
      {
        double s = (lo[i] - x[i]) / d[i];
        if (s < alpha) { alpha = s; blocking = i; bstate = 1; }
      }
    }
    for (int a = 0; a < m; a++) x[idx[a]] += alpha * d[idx[a]];
    if (blocking >= 0)
    {
      x[blocking] = bstate == 1 ? lo[blocking] : hi[blocking];
      state[blocking] = bstate;
      continue;
    }
    int worst = -1;
    double wv = gtol;
    for (int i = 0; i < n; i++)
    {
      if (state[i] != 1 && state[i] != 2) continue;
      double s = f[i];
      for (int k = 0; k < n; k++) s += H[i * n + k] * x[k];
      double viol = state[i] == 1 ? -s : s;
      if (viol > wv) { wv = viol; worst = i; }
    }
    if (worst < 0) break;
    state[worst] = 0;
  }
}

#endif
