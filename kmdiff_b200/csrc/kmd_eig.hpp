// Host-side dense symmetric eigensolver for the samples x samples Eigenstrat matrix (n ~ 10^2..10^3).
//
// Replaces eigvecs() of HAWK's smartpca (thirdparty/hawk/EIG6.0.1-Hawk/src/eigsubs.c:35-51, LAPACK
// dspev through eigx.c:95-114): all eigenvalues, descending, with unit eigenvectors.  The matrix is
// tiny next to the k-mer stream, so this stays on the host (north_star: "the samples x samples
// eigensolve stays on host"); LAPACKE headers are not part of this image, hence a self-contained
// Householder tridiagonalisation + implicit-shift QL iteration.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace kmd {

// a: n*n row-major symmetric.  evals[j] descending; evecs[j*n + i] = component i of eigenvector j,
// normalised so that its largest-magnitude component is positive (LAPACK leaves the sign unspecified).
// Returns 0, or 1 if the QL iteration failed to converge.
inline int sym_eig(const double* a, uint32_t n, double* evals, double* evecs)
{
  if (n == 0) return 0;
  std::vector<double> A(a, a + static_cast<size_t>(n) * n);
  std::vector<double> Q(static_cast<size_t>(n) * n, 0.0);  // accumulated reflectors, row-major
  for (uint32_t i = 0; i < n; i++) Q[static_cast<size_t>(i) * n + i] = 1.0;
  std::vector<double> v(n), p(n), d(n), e(n, 0.0);

  // ---- Householder: A <- H A H with H = I - 2 v v^T zeroing column k below the sub-diagonal ----
  for (uint32_t k = 0; k + 2 < n; k++)
  {
    const uint32_t m = n - k - 1;  // length of the sub-column
    double norm2 = 0.0;
    for (uint32_t i = 0; i < m; i++) { v[i] = A[static_cast<size_t>(k + 1 + i) * n + k]; norm2 += v[i] * v[i]; }
    const double norm = std::sqrt(norm2);
    double tail2 = norm2 - v[0] * v[0];
    if (norm == 0.0 || tail2 <= 0.0) continue;  // already tridiagonal in this column
    const double alpha = v[0] > 0 ? -norm : norm;
    v[0] -= alpha;
    double vn = std::sqrt(v[0] * v[0] + tail2);
    for (uint32_t i = 0; i < m; i++) v[i] /= vn;
    // p = A22 v, K = v^T p, q = p - K v, A22 -= 2 (v q^T + q v^T)
    double K = 0.0;
    for (uint32_t i = 0; i < m; i++)
    {
      const double* row = &A[static_cast<size_t>(k + 1 + i) * n + k + 1];
      double s = 0.0;
      for (uint32_t j = 0; j < m; j++) s += row[j] * v[j];
      p[i] = s;
      K += v[i] * s;
    }
    for (uint32_t i = 0; i < m; i++) p[i] -= K * v[i];
    for (uint32_t i = 0; i < m; i++)
    {
      double* row = &A[static_cast<size_t>(k + 1 + i) * n + k + 1];
      const double vi2 = 2.0 * v[i], pi2 = 2.0 * p[i];
      for (uint32_t j = 0; j < m; j++) row[j] -= vi2 * p[j] + pi2 * v[j];
    }
    // column / row k
    A[static_cast<size_t>(k + 1) * n + k] = alpha;
    A[static_cast<size_t>(k) * n + k + 1] = alpha;
    for (uint32_t i = 1; i < m; i++)
    {
      A[static_cast<size_t>(k + 1 + i) * n + k] = 0.0;
      A[static_cast<size_t>(k) * n + k + 1 + i] = 0.0;
    }
    // Q <- Q H (columns k+1.. of Q)
    for (uint32_t r = 0; r < n; r++)
    {
      double* row = &Q[static_cast<size_t>(r) * n + k + 1];
      double s = 0.0;
      for (uint32_t j = 0; j < m; j++) s += row[j] * v[j];
      s *= 2.0;
      for (uint32_t j = 0; j < m; j++) row[j] -= s * v[j];
    }
  }
  for (uint32_t i = 0; i < n; i++) d[i] = A[static_cast<size_t>(i) * n + i];
  for (uint32_t i = 0; i + 1 < n; i++) e[i] = A[static_cast<size_t>(i + 1) * n + i];
  e[n - 1] = 0.0;

  // ---- implicit-shift QL on (d, e), rotations applied to the columns of Q ----
  int rc = 0;
  const double eps = 2.220446049250313e-16;
  for (uint32_t l = 0; l < n; l++)
  {
    int iter = 0;
    uint32_t m;
    do
    {
      for (m = l; m + 1 < n; m++)
      {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= eps * dd) break;
      }
      if (m != l)
      {
        if (iter++ == 200) { rc = 1; break; }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, pp = 0.0;
        int i;
        for (i = static_cast<int>(m) - 1; i >= static_cast<int>(l); i--)
        {
          double f = s * e[i];
          const double b = c * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0)
          {
            d[i + 1] -= pp;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - pp;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (pp = s * r);
          g = c * r - b;
          for (uint32_t k = 0; k < n; k++)
          {
            double* row = &Q[static_cast<size_t>(k) * n];
            f = row[i + 1];
            row[i + 1] = s * row[i] + c * f;
            row[i] = c * row[i] - s * f;
          }
        }
        if (r == 0.0 && i >= static_cast<int>(l)) continue;
        d[l] -= pp;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }

  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return d[x] > d[y]; });
  for (uint32_t j = 0; j < n; j++)
  {
    const uint32_t col = order[j];
    evals[j] = d[col];
    uint32_t big = 0;
    for (uint32_t i = 1; i < n; i++)
      if (std::fabs(Q[static_cast<size_t>(i) * n + col]) > std::fabs(Q[static_cast<size_t>(big) * n + col])) big = i;
    const double sgn = Q[static_cast<size_t>(big) * n + col] < 0 ? -1.0 : 1.0;
    for (uint32_t i = 0; i < n; i++) evecs[static_cast<size_t>(j) * n + i] = sgn * Q[static_cast<size_t>(i) * n + col];
  }
  return rc;
}

}  // namespace kmd
