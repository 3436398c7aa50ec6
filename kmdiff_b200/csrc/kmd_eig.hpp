// Host-side dense symmetric eigensolver for the samples x samples Eigenstrat matrix (n ~ 10^2..10^3).
//
// Replaces eigvecs() of HAWK's smartpca (thirdparty/hawk/EIG6.0.1-Hawk/src/eigsubs.c:35-51, LAPACK
// dspev through eigx.c:95-114): all eigenvalues, descending, with unit eigenvectors.  The matrix is
// tiny next to the k-mer stream, so this stays on the host (north_star: "the samples x samples
// eigensolve stays on host"); LAPACKE headers are not part of this image, hence a self-contained
// Householder tridiagonalisation + implicit-shift QL iteration.  The reflectors and the QL rotations do not
// depend on the eigenvector matrix they are accumulated into, so they are recorded while (d, e) is reduced and
// then applied to Q in their original order by a few threads, each owning a block of rows of Q: every
// element of Q sees exactly the operation sequence of the textbook loop, only faster.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

namespace kmd {

// Far from the diagonal the accumulated eigenvector entries decay below 1e-308; x86 handles such subnormal
// operands in microcode (~100x slower), and they cannot influence a result that is printed with four
// decimals (pcs.evec), so the accumulation threads flush them to zero.
struct FlushSubnormals
{
#if defined(__x86_64__) || defined(__i386__)
  unsigned saved;
  FlushSubnormals() : saved(_mm_getcsr()) { _mm_setcsr(saved | 0x8040u); }   // FTZ | DAZ
  ~FlushSubnormals() { _mm_setcsr(saved); }
#endif
};

// Householder: A <- H A H with H = I - 2 v v^T zeroing column k below the sub-diagonal, k = 0 .. n-3; the reflectors
// are recorded (Q = H_0 H_1 ...).  Both O(m^2) sweeps of a step run along rows with independent elements, so the
// host compiler vectorises them without re-associating a single sum: p = A22 v is accumulated row by row as
// p += v_j * A22[j][:] — the trailing matrix stays exactly symmetric under the rank-2 update (its two products are
// 2 v_i p_j and 2 p_i v_j on either side of the diagonal, the sum commutes), so this is the dot-product form's sum in
// the same j order.
inline void tridiagonalize(std::vector<double>& A, uint32_t n, std::vector<double>& v, std::vector<double>& p,
                           std::vector<uint32_t>& refl_k, std::vector<double>& refl_v)
{
  for (uint32_t k = 0; k + 2 < n; k++)
  {
    const uint32_t m = n - k - 1;  // length of the sub-column
    double norm2 = 0.0;
    for (uint32_t i = 0; i < m; i++) { v[i] = A[static_cast<size_t>(k + 1 + i) * n + k]; norm2 += v[i] * v[i]; }
    const double norm = std::sqrt(norm2);
    const double tail2 = norm2 - v[0] * v[0];
    if (norm == 0.0 || tail2 <= 0.0) continue;  // already tridiagonal in this column
    const double alpha = v[0] > 0 ? -norm : norm;
    v[0] -= alpha;
    const double vn = std::sqrt(v[0] * v[0] + tail2);
    for (uint32_t i = 0; i < m; i++) v[i] /= vn;
    // p = A22 v, K = v^T p, q = p - K v, A22 -= 2 (v q^T + q v^T)
    double* __restrict__ pp = p.data();
    const double* __restrict__ vv = v.data();
    for (uint32_t i = 0; i < m; i++) pp[i] = 0.0;
    for (uint32_t j = 0; j < m; j++)
    {
      const double* __restrict__ row = &A[static_cast<size_t>(k + 1 + j) * n + k + 1];
      const double vj = vv[j];
      for (uint32_t i = 0; i < m; i++) pp[i] += row[i] * vj;
    }
    double K = 0.0;
    for (uint32_t i = 0; i < m; i++) K += vv[i] * pp[i];
    for (uint32_t i = 0; i < m; i++) pp[i] -= K * vv[i];
    for (uint32_t i = 0; i < m; i++)
    {
      double* __restrict__ row = &A[static_cast<size_t>(k + 1 + i) * n + k + 1];
      const double vi2 = 2.0 * vv[i], pi2 = 2.0 * pp[i];
      for (uint32_t j = 0; j < m; j++) row[j] -= vi2 * pp[j] + pi2 * vv[j];
    }
    // column / row k
    A[static_cast<size_t>(k + 1) * n + k] = alpha;
    A[static_cast<size_t>(k) * n + k + 1] = alpha;
    for (uint32_t i = 1; i < m; i++)
    {
      A[static_cast<size_t>(k + 1 + i) * n + k] = 0.0;
      A[static_cast<size_t>(k) * n + k + 1 + i] = 0.0;
    }
    // Q <- Q H (columns k+1.. of Q) is applied later: keep the reflector
    refl_k.push_back(k);
    refl_v.insert(refl_v.end(), v.begin(), v.begin() + m);
  }
}

struct Rot { double s, c; uint32_t i; };

// implicit-shift QL on (d, e); the rotations are recorded when `rot` is given.  Returns 1 if it failed to converge.
inline int ql_implicit(std::vector<double>& d, std::vector<double>& e, uint32_t n, std::vector<Rot>* rot)
{
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
          e[i + 1] = (r = std::sqrt(f * f + g * g));   // |f|, |g| <= ||A||: no overflow to guard against, unlike hypot
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
          if (rot) rot->push_back({s, c, static_cast<uint32_t>(i)});   // columns i, i+1 of Q, applied later
        }
        if (r == 0.0 && i >= static_cast<int>(l)) continue;
        d[l] -= pp;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return rc;
}

// a: n*n row-major symmetric.  evals[j] descending; evecs[j*n + i] = component i of eigenvector j,
// normalised so that its largest-magnitude component is positive (LAPACK leaves the sign unspecified).
// Returns 0, or 1 if the QL iteration failed to converge.
inline int sym_eig(const double* a, uint32_t n, double* evals, double* evecs)
{
  if (n == 0) return 0;
  std::vector<double> A(a, a + static_cast<size_t>(n) * n);
  // accumulated reflectors, stored TRANSPOSED (Q[c * n + r] = element (r, c)): every Householder update and every
  // QL rotation then works on contiguous rows; the arithmetic and its order are those of the plain layout
  std::vector<double> Q_store(static_cast<size_t>(n) * n + 8, 0.0);
  // 64-byte aligned, so that the threads' row blocks (multiples of 8 doubles) never share a cache line
  double* Q = Q_store.data() + ((64 - reinterpret_cast<uintptr_t>(Q_store.data()) % 64) % 64) / sizeof(double);
  std::vector<Rot> rot;
  std::vector<uint32_t> refl_k;
  std::vector<double> refl_v;
  rot.reserve(static_cast<size_t>(n) * n * 2);
  refl_v.reserve(static_cast<size_t>(n) * n / 2 + n);
  std::vector<double> v(n), p(n), d(n), e(n, 0.0);

  tridiagonalize(A, n, v, p, refl_k, refl_v);
  for (uint32_t i = 0; i < n; i++) d[i] = A[static_cast<size_t>(i) * n + i];
  for (uint32_t i = 0; i + 1 < n; i++) e[i] = A[static_cast<size_t>(i + 1) * n + i];
  e[n - 1] = 0.0;

  // ---- implicit-shift QL on (d, e), rotations applied to the columns of Q ----
  const int rc = ql_implicit(d, e, n, &rot);

  // ---- apply the recorded reflectors, then the rotations, to Q: rows [r0, r1) per thread ----
  {
    auto apply = [&](uint32_t r0, uint32_t r1) {
      FlushSubnormals ftz;
      const uint32_t w = r1 - r0;
      // private compact copy of the block (n x w, identity rows r0..r1): sequential, cache-resident sweeps
      std::vector<double> B(static_cast<size_t>(n) * w, 0.0);
      for (uint32_t r = r0; r < r1; r++) B[static_cast<size_t>(r) * w + (r - r0)] = 1.0;
      std::vector<double> sacc(w);
      size_t off = 0;
      for (uint32_t k : refl_k)
      {
        const uint32_t m = n - k - 1;
        const double* vk = &refl_v[off];
        off += m;
        // s_r = 2 sum_j Q(r, k+1+j) v_j in j order, then Q(r, k+1+j) -= s_r v_j
        std::fill(sacc.begin(), sacc.end(), 0.0);
        for (uint32_t j = 0; j < m; j++)
        {
          const double* col = &B[static_cast<size_t>(k + 1 + j) * w];
          const double vj = vk[j];
          for (uint32_t r = 0; r < w; r++) sacc[r] += col[r] * vj;
        }
        for (uint32_t r = 0; r < w; r++) sacc[r] *= 2.0;
        for (uint32_t j = 0; j < m; j++)
        {
          double* col = &B[static_cast<size_t>(k + 1 + j) * w];
          const double vj = vk[j];
          for (uint32_t r = 0; r < w; r++) col[r] -= sacc[r] * vj;
        }
      }
      for (const Rot& t : rot)
      {
        double* __restrict__ ci = &B[static_cast<size_t>(t.i) * w];
        double* __restrict__ ci1 = ci + w;
        const double s = t.s, c = t.c;
        for (uint32_t r = 0; r < w; r++)
        {
          const double f = ci1[r];
          const double g0 = ci[r];
          ci1[r] = s * g0 + c * f;
          ci[r] = c * g0 - s * f;
        }
      }
      for (uint32_t c = 0; c < n; c++) std::copy(&B[static_cast<size_t>(c) * w], &B[static_cast<size_t>(c) * w] + w, &Q[static_cast<size_t>(c) * n + r0]);
    };
    const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t n_thr = std::max(1u, std::min({hw, 16u, n / 24u}));
    if (n_thr == 1) apply(0, n);
    else
    {
      std::vector<std::thread> pool;
      const uint32_t per = ((n + n_thr - 1) / n_thr + 7) / 8 * 8;
      for (uint32_t t = 0; t < n_thr; t++)
      {
        const uint32_t r0 = std::min(n, t * per), r1 = std::min(n, r0 + per);
        if (r0 < r1) pool.emplace_back(apply, r0, r1);
      }
      for (auto& t : pool) t.join();
    }
  }

  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return d[x] > d[y]; });
  for (uint32_t j = 0; j < n; j++)
  {
    const uint32_t col = order[j];
    evals[j] = d[col];
    const double* qc = &Q[static_cast<size_t>(col) * n];
    uint32_t big = 0;
    for (uint32_t i = 1; i < n; i++)
      if (std::fabs(qc[i]) > std::fabs(qc[big])) big = i;
    const double sgn = qc[big] < 0 ? -1.0 : 1.0;
    for (uint32_t i = 0; i < n; i++) evecs[static_cast<size_t>(j) * n + i] = sgn * qc[i];
  }
  return rc;
}

// The k leading eigenpairs only (what smartpca's output needs: all eigenvalues, numeigs eigenvectors): eigenvalues by QL
// without accumulating rotations, eigenvectors of the tridiagonal matrix by inverse iteration (LU with partial
// pivoting, re-orthogonalised inside clusters), carried back through the recorded reflectors — O(n^2) per vector
// instead of the O(n^3) accumulation of all n.  evals: all n, descending; evecs[j*n + i], j < k, largest component
// positive.  Returns 0, or 1 if an iteration failed to converge.
inline int sym_eig_top(const double* a, uint32_t n, uint32_t k, double* evals, double* evecs)
{
  if (n == 0) return 0;
  k = std::min(k, n);
  std::vector<double> A(a, a + static_cast<size_t>(n) * n);
  std::vector<uint32_t> refl_k;
  std::vector<double> refl_v;
  refl_v.reserve(static_cast<size_t>(n) * n / 2 + n);
  std::vector<double> v(n), p(n), d(n), e(n, 0.0);
  tridiagonalize(A, n, v, p, refl_k, refl_v);
  for (uint32_t i = 0; i < n; i++) d[i] = A[static_cast<size_t>(i) * n + i];
  for (uint32_t i = 0; i + 1 < n; i++) e[i] = A[static_cast<size_t>(i + 1) * n + i];
  e[n - 1] = 0.0;
  const std::vector<double> d0(d), e0(e);
  int rc = ql_implicit(d, e, n, nullptr);
  std::sort(d.begin(), d.end(), [](double x, double y) { return x > y; });
  for (uint32_t j = 0; j < n; j++) evals[j] = d[j];
  if (k == 0) return rc;

  double tnorm = 0.0;
  for (uint32_t i = 0; i < n; i++) tnorm = std::max(tnorm, std::fabs(d0[i]) + (i ? std::fabs(e0[i - 1]) : 0.0) + std::fabs(e0[i]));
  if (tnorm == 0.0) tnorm = 1.0;
  const double eps = 2.220446049250313e-16;
  std::vector<double> dl(n), dd(n), du(n), du2(n), x(n);
  std::vector<uint8_t> swapped(n);
  std::vector<double> Y(static_cast<size_t>(k) * n);   // eigenvectors of the tridiagonal matrix
  uint64_t lcg = 0x9E3779B97F4A7C15ull;
  double prev_shift = 0.0;
  for (uint32_t j = 0; j < k; j++)
  {
    // shift: the eigenvalue, nudged off an exactly repeated one so that the factorisations differ
    double shift = evals[j];
    if (j && std::fabs(shift - prev_shift) < 10.0 * eps * tnorm) shift = prev_shift - 10.0 * eps * tnorm;
    prev_shift = shift;
    // LU of T - shift I with partial pivoting (tridiagonal: one fill-in diagonal)
    for (uint32_t i = 0; i < n; i++) { dd[i] = d0[i] - shift; du[i] = i + 1 < n ? e0[i] : 0.0; dl[i] = du[i]; du2[i] = 0.0; }
    for (uint32_t i = 0; i + 1 < n; i++)
    {
      if (std::fabs(dd[i]) >= std::fabs(dl[i]))
      {
        if (dd[i] == 0.0) dd[i] = eps * tnorm;
        const double f = dl[i] / dd[i];
        dl[i] = f;
        dd[i + 1] -= f * du[i];
        swapped[i] = 0;
      }
      else
      {
        const double f = dd[i] / dl[i];
        dd[i] = dl[i];
        dl[i] = f;
        const double t = du[i];
        du[i] = dd[i + 1];
        dd[i + 1] = t - f * dd[i + 1];
        if (i + 2 < n) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
        swapped[i] = 1;
      }
    }
    if (dd[n - 1] == 0.0) dd[n - 1] = eps * tnorm;
    for (uint32_t i = 0; i < n; i++)
    {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      x[i] = static_cast<double>(lcg >> 11) / 9007199254740992.0 - 0.5;
    }
    double* y = &Y[static_cast<size_t>(j) * n];
    for (int it = 0; it < 6; it++)
    {
      // solve (T - shift I) x' = x
      for (uint32_t i = 0; i + 1 < n; i++)
      {
        if (!swapped[i]) x[i + 1] -= dl[i] * x[i];
        else { const double t = x[i]; x[i] = x[i + 1]; x[i + 1] = t - dl[i] * x[i]; }
      }
      x[n - 1] /= dd[n - 1];
      if (n > 1) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / dd[n - 2];
      for (int i = static_cast<int>(n) - 3; i >= 0; i--) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / dd[i];
      // keep it orthogonal to the earlier vectors of its cluster
      for (uint32_t q = 0; q < j; q++)
      {
        if (std::fabs(evals[q] - evals[j]) > 1e-3 * tnorm) continue;
        const double* yq = &Y[static_cast<size_t>(q) * n];
        double dot = 0.0;
        for (uint32_t i = 0; i < n; i++) dot += yq[i] * x[i];
        for (uint32_t i = 0; i < n; i++) x[i] -= dot * yq[i];
      }
      double nrm = 0.0;
      for (uint32_t i = 0; i < n; i++) nrm += x[i] * x[i];
      nrm = std::sqrt(nrm);
      if (!(nrm > 0.0) || !std::isfinite(nrm)) { rc = 1; break; }
      for (uint32_t i = 0; i < n; i++) x[i] /= nrm;
      // converged when a further solve no longer turns the vector
      double diff = 0.0, sum = 0.0;
      for (uint32_t i = 0; i < n; i++) { diff += (x[i] - y[i]) * (x[i] - y[i]); sum += (x[i] + y[i]) * (x[i] + y[i]); }
      std::copy(x.begin(), x.end(), y);
      if (it >= 2 && std::min(diff, sum) < 1e-28) break;
    }
  }
  // eigenvectors of A: Q y with Q = H_0 H_1 ... (apply the reflectors last to first)
  std::vector<size_t> off(refl_k.size());
  {
    size_t o = 0;
    for (size_t r = 0; r < refl_k.size(); r++) { off[r] = o; o += n - refl_k[r] - 1; }
  }
  for (uint32_t j = 0; j < k; j++)
  {
    double* y = &Y[static_cast<size_t>(j) * n];
    for (size_t r = refl_k.size(); r-- > 0;)
    {
      const uint32_t kk = refl_k[r], m = n - kk - 1;
      const double* vk = &refl_v[off[r]];
      double sdot = 0.0;
      for (uint32_t i = 0; i < m; i++) sdot += vk[i] * y[kk + 1 + i];
      sdot *= 2.0;
      for (uint32_t i = 0; i < m; i++) y[kk + 1 + i] -= sdot * vk[i];
    }
    uint32_t big = 0;
    for (uint32_t i = 1; i < n; i++)
      if (std::fabs(y[i]) > std::fabs(y[big])) big = i;
    const double sgn = y[big] < 0 ? -1.0 : 1.0;
    for (uint32_t i = 0; i < n; i++) evecs[static_cast<size_t>(j) * n + i] = sgn * y[i];
  }
  return rc;
}

}  // namespace kmd
