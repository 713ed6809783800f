// All eigenvalues of a symmetric tridiagonal matrix by bisection on Sturm counts -- the values-only back end of
// the symmetric eigensolver (a2: PCA spectrum, R/SimilarityM.R:41-72; a14: Laplacian spectrum,
// R/Clustering.R:96-105, where the reference calls eigen(only.values = TRUE)).  Eigenvalue k is found
// independently of all others (one GPU thread each, the recurrence streams d and e^2 through L1 with every lane of
// a warp at the same index), to an absolute accuracy of a few ulp of ||T||: the accuracy class of the implicit QL
// iteration, without its O(n^2) sequential sweep on the host.  The arithmetic is shared with the CPU checks
// (rsoptsc_host_tridiag_bisect, tests/test_abi.py).
#pragma once
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD inline
#endif

namespace rs {
namespace bisect {

// number of eigenvalues of T strictly below x (LDL^T pivots q_i of T - x I; a vanishing pivot is pushed to
// -pivmin as in LAPACK dstebz, so the count never depends on a division by zero)
RS_HD int sturm_count(const double* d, const double* e2, int n, double x, double pivmin) {
  int cnt = 0;
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += q < 0.0;
  for (int i = 1; i < n; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;
}

// k-th smallest eigenvalue (k = 0 .. n-1) inside the Gershgorin interval [gl, gu]
RS_HD double kth_eigenvalue(const double* d, const double* e2, int n, int k, double gl, double gu, double pivmin) {
  double lo = gl, hi = gu;
  const double tnorm = fmax(fabs(gl), fabs(gu));
  const double abstol = 2.0 * DBL_EPSILON * tnorm + 2.0 * pivmin;
  for (int it = 0; it < 128; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (hi - lo <= abstol + 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) || mid <= lo || mid >= hi) break;
    if (sturm_count(d, e2, n, mid, pivmin) > k) hi = mid;  // at least k+1 eigenvalues below mid
    else lo = mid;
  }
  return 0.5 * (lo + hi);
}

}  // namespace bisect
}  // namespace rs
