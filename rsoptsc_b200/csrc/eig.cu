// Dense symmetric eigensolver used by the SVT (a4), the PCA spectrum (a2) and the Laplacian
// spectrum (a14): cuSOLVER's 64-bit API (cusolverDnXsyevd).  This is the one library call that
// dominates the path; it is isolated here so that a native two-stage solver can
// replace it.  The 64-bit entry point removes the int overflow of the legacy workspace size for
// blocks above ~23,000 rows.
#include "internal.h"

namespace rs {

// A (n x n, lower triangle referenced) -> eigenvectors in A (when `vectors`), eigenvalues ascending
// in d_lam.  Throws when the solver reports failure.
void sym_eig(double* d_A, int64_t n, double* d_lam, bool vectors) {
  Context& c = ctx();
  RS_REQUIRE(n >= 1, "sym_eig: empty matrix");
  // probed on B200 / CUDA 12.9 (scratch/eig_limit_probe.cu): Xsyevd_bufferSize returns INVALID_VALUE above 32,768
  RS_REQUIRE(n <= 32768, "symmetric eigensolver: a connected block of " + std::to_string(n) +
                             " cells exceeds cuSOLVER's 32,768-row limit (largest supported dense block)");
  static cusolverDnParams_t params = nullptr;
  if (!params) RS_CUSOLVER(cusolverDnCreateParams(&params));
  const cusolverEigMode_t mode = vectors ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR;
  size_t wdev = 0, whost = 0;
  RS_CUSOLVER(cusolverDnXsyevd_bufferSize(c.cusolver, params, mode, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, d_A, n,
                                          CUDA_R_64F, d_lam, CUDA_R_64F, &wdev, &whost));
  DevBuf<char> work(std::max<size_t>(wdev, 1));
  std::vector<char> hwork(std::max<size_t>(whost, 1));
  DevBuf<int> info(1);
  RS_CUSOLVER(cusolverDnXsyevd(c.cusolver, params, mode, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, d_A, n, CUDA_R_64F,
                               d_lam, CUDA_R_64F, work.p, wdev, hwork.data(), whost, info.p));
  int hinfo = 0;
  RS_CUDA(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  RS_REQUIRE(hinfo == 0, "symmetric eigensolver did not converge (info=" + std::to_string(hinfo) + ")");
}

}  // namespace rs
