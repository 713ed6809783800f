// a4: singular value thresholding  J = U max(S - tau, 0) V^T  (R/SimilarityM.R:142-155).
//
// The reference takes a full SVD of the n x n matrix A = Z - Y3/mu.  Here the same J is formed
// from the Gram matrix:  G = A^T A = V diag(lambda) V^T,  sigma = sqrt(lambda),
//     J = A V_r diag(1 - tau/sigma_r) V_r^T          (r = #{sigma > tau})
// i.e. one syrk-shaped contraction, one symmetric eigen-decomposition and two GEMMs whose inner
// dimension shrinks with the rank of J.  In FP64 this matches the SVD route to ~1e-15 relative
// in Z (scratch/proto_admm.py); FP32-level Gram accuracy does NOT (7e-7), because squaring
// buries singular values below sqrt(eps)*sigma_1 -- hence the FP64(-equivalent) requirement, met
// on the tensor cores by the split-integer GEMM of gemm_ozaki.cu.
#include <cmath>
#include <thread>

#include "comm.h"
#include "gemm.h"
#include "internal.h"

namespace rs {

void SvtWorkspace::ensure(int64_t n_) {
  if (n == n_ && G.p) return;
  n = n_;
  G.alloc((size_t)n * n);
  T.alloc((size_t)n * n);
  lam.alloc(n);
}

// Second half of the SVT: eigenpairs of A^T A (V in the columns of d_G, lambda ascending in d_lam) ->
// J = A V_r diag(1 - tau / sigma_r) V_r^T over A.  d_G is used as scratch.  Returns rank(J).
static int64_t svt_reconstruct(double* d_A, int64_t n, double tau, double* d_G, double* d_lam, DevBuf<double>& T,
                               DevBuf<double>& B) {
  Context& c = ctx();
  if (!T.p || T.n < (size_t)n * n) T.alloc((size_t)n * n);
  std::vector<double> lam(n);
  RS_CUDA(cudaMemcpyAsync(lam.data(), d_lam, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  // shrink factors; eigenvalues ascending so the kept ones are a suffix
  std::vector<double> h(n, 0.0);
  int64_t r0 = n;
  for (int64_t i = n - 1; i >= 0; --i) {
    const double sig = lam[i] > 0 ? std::sqrt(lam[i]) : 0.0;
    if (sig > tau) { h[i] = 1.0 - tau / sig; r0 = i; } else break;
  }
  const int64_t r = n - r0;
  if (r == 0) {
    RS_CUDA(cudaMemsetAsync(d_A, 0, sizeof(double) * n * n, c.stream));
    return 0;
  }
  const double* Vr = d_G + r0 * n;
  Phase ph("recon");
  if (3 * r > 2 * n && n >= 1536) {
    // nearly full rank (saturated phase): J = A B with the SYMMETRIC B = V_r h V_r^T = W W^T,
    // W = V_r diag(sqrt h): one half-GEMM (n^2 r) + one GEMM (2 n^3) instead of two GEMMs (4 n^2 r)
    for (int64_t i = r0; i < n; ++i) h[i] = std::sqrt(h[i]);
    RS_CUDA(cudaMemcpyAsync(d_lam + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
    copy_scale_columns(Vr, n, r, d_lam + r0, T.p);   // W                     (n x r)
    B.alloc((size_t)n * n);
    gram_AAt(T.p, n, r, B.p);                        // B = W W^T  (lower)    (n x n)
    mirror_lower(B.p, n);
    gemm_nn(d_A, B.p, d_G, n, n, n);                 // J = A B, into G (V is no longer needed)
    RS_CUDA(cudaMemcpyAsync(d_A, d_G, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, c.stream));
    B.release();
  } else {
    RS_CUDA(cudaMemcpyAsync(d_lam + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
    gemm_nn(d_A, Vr, T.p, n, r, n);                 // T = A V_r          (n x r)
    scale_columns(T.p, n, r, d_lam + r0);        // T *= diag(h_r)
    gemm_nt(T.p, Vr, d_A, n, n, r);                 // J = T V_r^T        (n x n), over A
  }
  return r;
}

int64_t svt_inplace(SvtWorkspace& ws, double* d_A, int64_t n, double tau) {
  RS_REQUIRE(n <= 65535, "svt: block larger than 65,535 cells");
  ws.ensure(n);
  Context& c = ctx();
  {  // G = A^T A (lower triangle)
    Phase ph("gram");
    gram_AtA(d_A, n, n, ws.G.p);
  }
  {  // G -> V (columns), lam ascending
    Phase ph("eig");
    // only the eigenvectors with sqrt(lambda) > tau enter J (margin: the count below is re-derived from lambda)
    sym_eig(ws.G.p, n, ws.lam.p, /*vectors=*/true, tau * tau * (1.0 - 1e-9));
  }
  return svt_reconstruct(d_A, n, tau, ws.G.p, ws.lam.p, ws.T, ws.B);
}

// ---- several small blocks at once ---------------------------------------------------------------------------
// A connected block of ~1,000 cells keeps the GPU busy for a fraction of each of the ~10 us steps of its eigensolve
// (latency-bound tridiagonalisation inside cusolverDnXsyevd).  Independent small blocks therefore run their Gram +
// eigensolve side by side on two extra streams (own cuBLAS / cuSOLVER handles); the library stream waits for both
// and then does the reconstructions.  All buffers of the side work are allocated before it starts and released after
// the join, so the single-stream assumption of the caching allocator holds.
namespace {
struct SideLane {
  cudaStream_t stream = nullptr;
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  cudaEvent_t done = nullptr;
};
SideLane* side_lanes() {
  static SideLane lanes[2];
  static int lanes_device = -1;
  Context& c = ctx();
  if (lanes_device != c.device) {
    for (auto& l : lanes) {  // device switch: the old objects belong to the old device
      if (l.cusolver) cusolverDnDestroy(l.cusolver);
      if (l.cublas) cublasDestroy(l.cublas);
      if (l.done) cudaEventDestroy(l.done);
      if (l.stream) cudaStreamDestroy(l.stream);
      l = SideLane();
      RS_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
      RS_CUBLAS(cublasCreate(&l.cublas));
      RS_CUBLAS(cublasSetStream(l.cublas, l.stream));
      RS_CUSOLVER(cusolverDnCreate(&l.cusolver));
      RS_CUSOLVER(cusolverDnSetStream(l.cusolver, l.stream));
      RS_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
    }
    lanes_device = c.device;
  }
  return lanes;
}
struct SmallJob {
  DevBuf<double> G, lam;
  DevBuf<char> work;
  std::vector<char> hwork;
  DevBuf<int> info;
};
}  // namespace

void svt_small_batch(const std::vector<double*>& blocks, const std::vector<int64_t>& sizes, double tau) {
  Context& c = ctx();
  const size_t nb = blocks.size();
  SideLane* lanes = side_lanes();
  static cusolverDnParams_t params = nullptr;
  if (!params) RS_CUSOLVER(cusolverDnCreateParams(&params));
  std::vector<SmallJob> jobs(nb);
  {
    Phase ph("svt_small_eig");
    struct EventGuard {
      cudaEvent_t e = nullptr;
      ~EventGuard() { if (e) cudaEventDestroy(e); }
    } ready_guard;
    RS_CUDA(cudaEventCreateWithFlags(&ready_guard.e, cudaEventDisableTiming));
    const cudaEvent_t ready = ready_guard.e;
    RS_CUDA(cudaEventRecord(ready, c.stream));  // the blocks of A are complete on the library stream
    // buffers first, on this thread (the caching allocator is not thread-safe)
    std::vector<size_t> wdev(nb, 0), whost(nb, 0);
    for (size_t i = 0; i < nb; ++i) {
      const int64_t n = sizes[i];
      SmallJob& J = jobs[i];
      J.G.alloc((size_t)n * n); J.lam.alloc(n); J.info.alloc(1);
      RS_CUSOLVER(cusolverDnXsyevd_bufferSize(lanes[i & 1].cusolver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                              CUDA_R_64F, J.G.p, n, CUDA_R_64F, J.lam.p, CUDA_R_64F, &wdev[i], &whost[i]));
      J.work.alloc(std::max<size_t>(wdev[i], 1));
      J.hwork.resize(std::max<size_t>(whost[i], 1));
    }
    // cusolverDnXsyevd blocks its calling thread (host-side steps between its kernels), so each lane is driven by its
    // own host thread: the calling thread takes lane 0, one helper thread lane 1.  Failures are carried back as text.
    auto drive = [&](int lane, std::string* err) {
      try {
        RS_CUDA(cudaSetDevice(c.device));
        SideLane& L = lanes[lane];
        RS_CUDA(cudaStreamWaitEvent(L.stream, ready, 0));
        for (size_t i = lane; i < nb; i += 2) {
          const int64_t n = sizes[i];
          SmallJob& J = jobs[i];
          const double one = 1.0, zero = 0.0;
          RS_CUBLAS(cublasDsyrk(L.cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, (int)n, (int)n, &one, blocks[i], (int)n, &zero,
                                J.G.p, (int)n));                            // G = A^T A (lower)
          RS_CUSOLVER(cusolverDnXsyevd(L.cusolver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, J.G.p,
                                       n, CUDA_R_64F, J.lam.p, CUDA_R_64F, J.work.p, wdev[i], J.hwork.data(), whost[i],
                                       J.info.p));
        }
      } catch (const std::exception& e) {
        *err = e.what();
      } catch (...) {
        *err = "unknown error";
      }
    };
    std::string err0, err1;
    std::thread helper(drive, 1, &err1);
    drive(0, &err0);
    helper.join();
    RS_REQUIRE(err0.empty() && err1.empty(), "small-block SVT: " + (err0.empty() ? err1 : err0));
    for (int l = 0; l < 2; ++l) {
      RS_CUDA(cudaEventRecord(lanes[l].done, lanes[l].stream));
      RS_CUDA(cudaStreamWaitEvent(c.stream, lanes[l].done, 0));  // join: the library stream continues after both lanes
    }
  }
  DevBuf<double> T, B;
  for (size_t i = 0; i < nb; ++i) {
    int hinfo = 0;
    RS_CUDA(cudaMemcpyAsync(&hinfo, jobs[i].info.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RS_CUDA(cudaStreamSynchronize(c.stream));
    RS_REQUIRE(hinfo == 0, "symmetric eigensolver did not converge (info=" + std::to_string(hinfo) + ")");
    svt_reconstruct(blocks[i], sizes[i], tau, jobs[i].G.p, jobs[i].lam.p, T, B);
  }
  RS_CUDA(cudaStreamSynchronize(c.stream));  // the jobs' buffers go back to the allocator only after all their work
}

// ---- collective SVT of a block replicated on every rank (multi-GPU) ----------
//   Gram           every rank forms one column panel of the lower triangle of G = A^T A (panels of
//                  equal area), the panels are all-gathered over NVLink
//   eigensolve     sym_eig(collective): fused multi-GPU tridiagonalisation, divide-and-conquer merges
//                  and back-transformation split by columns
//   reconstruction rank q forms its own columns of J only:  J[:, q] = A (V_r h V_r[q, :]^T)
// so J never travels: the column-local ADMM kernels of rank q read exactly these columns.
int64_t svt_sharded(SvtWorkspace& ws, double* d_A, int64_t n, double tau, const std::vector<int64_t>& cb) {
  RS_REQUIRE(n <= 65535, "svt: block larger than 65,535 cells");
  ws.ensure(n);
  Context& c = ctx();
  const Comm& cm = comm();
  const int P = cm.nranks, rank = cm.rank;
  {  // G = A^T A, lower triangle, one panel per rank
    Phase ph("gram");
    const std::vector<int64_t> gb = split_lower_triangle(n, P, 128);
    const int64_t g0 = gb[rank], g1 = gb[rank + 1];
    if (g1 > g0)  // rows g0.. of the panel: G[g0:, g0:g1] = A[:, g0:]^T A[:, g0:g1]
      gemm_tn_ld(d_A + (size_t)g0 * n, n, d_A + (size_t)g0 * n, n, ws.G.p + g0 + (size_t)g0 * n, n, n - g0, g1 - g0, n);
    comm_allgather_cols(ws.G.p, n, gb);
  }
  {
    Phase ph("eig");
    sym_eig(ws.G.p, n, ws.lam.p, /*vectors=*/true, tau * tau * (1.0 - 1e-9), /*collective=*/true);
  }
  std::vector<double> lam(n);
  RS_CUDA(cudaMemcpyAsync(lam.data(), ws.lam.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  std::vector<double> h(n, 0.0);
  int64_t r0 = n;
  for (int64_t i = n - 1; i >= 0; --i) {
    const double sig = lam[i] > 0 ? std::sqrt(lam[i]) : 0.0;
    if (sig > tau) { h[i] = 1.0 - tau / sig; r0 = i; } else break;
  }
  const int64_t r = n - r0;
  const int64_t c0 = cb[rank], c1 = cb[rank + 1], w = c1 - c0;
  if (r == 0) {
    if (w > 0) RS_CUDA(cudaMemsetAsync(d_A + (size_t)c0 * n, 0, sizeof(double) * n * w, c.stream));
    return 0;
  }
  Phase ph("recon");
  if (w > 0) {
    const double* Vr = ws.G.p + (size_t)r0 * n;
    RS_CUDA(cudaMemcpyAsync(ws.lam.p + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
    copy_scale_columns(Vr, n, r, ws.lam.p + r0, ws.T.p);                  // T = V_r diag(h)          (n x r)
    DevBuf<double> Bq((size_t)n * w), Jq((size_t)n * w);
    gemm_nt_ld(ws.T.p, n, Vr + c0, n, Bq.p, n, n, w, r);                  // B[:, q] = T V_r[q, :]^T  (n x w)
    gemm_nn_ld(d_A, n, Bq.p, n, Jq.p, n, n, w, n);                        // J[:, q] = A B[:, q]
    RS_CUDA(cudaMemcpyAsync(d_A + (size_t)c0 * n, Jq.p, sizeof(double) * n * w, cudaMemcpyDeviceToDevice, c.stream));
  }
  return r;
}

}  // namespace rs
