// Probe: the cooperative-kernel tridiagonalisation (rsoptsc_b200/csrc/sytrd_kernel.cuh) against
// cusolverDnDsytrd -- same LAPACK conventions, so d, e, tau and the reflectors must agree to rounding.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo scratch/sytrd_probe.cu -o scratch/sytrd_probe -lcublas -lcusolver
//   ./scratch/sytrd_probe 7954 [reps]
#include <cusolverDn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../rsoptsc_b200/csrc/sytrd_kernel.cuh"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t st_ = (x);                                                        \
    if (st_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(st_), __FILE__, __LINE__); \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

__global__ void k_fill(double* A, long long n, unsigned long long seed) {
  // symmetric pseudo-random matrix: value depends on (min, max) of the indices
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const long long i = idx % n, j = idx / n;
  const unsigned long long a = i < j ? i : j, b = i < j ? j : i;
  unsigned long long h = seed + a * 0x9E3779B97F4A7C15ull + b * 0xBF58476D1CE4E5B9ull;
  h ^= h >> 31; h *= 0x94D049BB133111EBull; h ^= h >> 29; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
  double v = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  if (i == j) v += 1.0;
  if (i < j) v = nan("");  // the upper triangle must never be read
  A[idx] = v;
}

static double maxdiff(const std::vector<double>& a, const std::vector<double>& b, double* scale) {
  double m = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    if (a[i] != a[i] && b[i] != b[i]) continue;
    m = std::max(m, std::fabs(a[i] - b[i]));
    s = std::max(s, std::fabs(a[i]));
  }
  *scale = s;
  return m;
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 2000;
  const int reps = argc > 2 ? atoi(argv[2]) : 2;
  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  cublasHandle_t blas;
  cublasCreate(&blas);
  cusolverDnHandle_t sol;
  cusolverDnCreate(&sol);
  cusolverDnSetStream(sol, stream);

  double *A0, *A1, *A2, *d1, *e1, *t1, *d2, *e2, *t2, *work2;
  CK(cudaMalloc(&A0, sizeof(double) * n * n));
  CK(cudaMalloc(&A1, sizeof(double) * n * n));
  CK(cudaMalloc(&A2, sizeof(double) * n * n));
  for (double** p : {&d1, &e1, &t1, &d2, &e2, &t2}) CK(cudaMalloc(p, sizeof(double) * n));
  CK(cudaMalloc(&work2, sizeof(double) * rs::sytrd::workspace_doubles(n)));
  k_fill<<<(unsigned)((n * n + 255) / 256), 256, 0, stream>>>(A0, n, 12345);
  CK(cudaStreamSynchronize(stream));

  int lwork = 0;
  cusolverDnDsytrd_bufferSize(sol, CUBLAS_FILL_MODE_LOWER, (int)n, A1, (int)n, d1, e1, t1, &lwork);
  double* work1;
  int* info;
  CK(cudaMalloc(&work1, sizeof(double) * (size_t)lwork));
  CK(cudaMalloc(&info, sizeof(int)));

  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  float ms_ref = 1e30f, ms_own = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemcpyAsync(A1, A0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, stream));
    cudaEventRecord(ev0, stream);
    cusolverDnDsytrd(sol, CUBLAS_FILL_MODE_LOWER, (int)n, A1, (int)n, d1, e1, t1, work1, lwork, info);
    cudaEventRecord(ev1, stream);
    CK(cudaStreamSynchronize(stream));
    float ms;
    cudaEventElapsedTime(&ms, ev0, ev1);
    ms_ref = std::min(ms_ref, ms);
  }
  int64_t launches = 0;
  long long* prof = nullptr;
  if (getenv("SYTRD_PROF")) {
    CK(cudaMalloc(&prof, 8 * 1100));
    CK(cudaMemset(prof, 0, 8 * 1100));
  }
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemcpyAsync(A2, A0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, stream));
    cudaEventRecord(ev0, stream);
    launches = 0;
    CK(rs::sytrd::run(blas, stream, A2, n, n, d2, e2, t2, work2, &launches, prof));
    cudaEventRecord(ev1, stream);
    CK(cudaStreamSynchronize(stream));
    float ms;
    cudaEventElapsedTime(&ms, ev0, ev1);
    ms_own = std::min(ms_own, ms);
  }
  printf("n=%lld  cusolverDnDsytrd %.2f ms   own %.2f ms (%lld launches)   ratio %.2fx\n", n, ms_ref, ms_own,
         (long long)launches, ms_ref / ms_own);
  if (prof) {
    long long h[1100];
    CK(cudaMemcpy(h, prof, 8 * 1100, cudaMemcpyDeviceToHost));
    const char* names[7] = {"(unused)", "allreduce1 (norm, alpha)", "reflector + product phase", "allreduce2 (v^T A v)",
                            "row phase c (row loop)", "row phase a (panel dots)", "row phase b (w(j+1), alpha2)"};
    printf("  staged strips consumed by warp 0 of CTA 0: %lld of %lld columns\n", h[12], (long long)n);

    for (int k = 0; k < 6; ++k)
      printf("  prof %-28s %8.2f ms total over %d reps (cycles/1.9GHz)  %.2f us/column\n", names[k], h[k] / 1.9e6, reps,
             h[k] / 1.9e3 / reps / (double)n);
  }
  if (prof) {
    long long h[1100];
    CK(cudaMemcpy(h, prof, 8 * 1100, cudaMemcpyDeviceToHost));
    std::vector<std::pair<long long, int>> v;
    for (int c = 0; c < 148; ++c) v.push_back({h[16 + c], c});
    std::sort(v.begin(), v.end());
    printf("  phase-3 time per CTA (ms): min %.1f (cta %d)  p25 %.1f  median %.1f  p75 %.1f  max %.1f (cta %d)\n", v[0].first / 1.9e6,
           v[0].second, v[37].first / 1.9e6, v[74].first / 1.9e6, v[111].first / 1.9e6, v[147].first / 1.9e6, v[147].second);
    printf("  slowest 10 CTAs:");
    for (int k = 138; k < 148; ++k) printf(" %d(%.0f)", v[k].second, v[k].first / 1.9e6);
    printf("\n  fastest 10 CTAs:");
    for (int k = 0; k < 10; ++k) printf(" %d(%.0f)", v[k].second, v[k].first / 1.9e6);
    printf("\n");
  }
  const double bytes = 4.0 / 3.0 * (double)n * n * n;
  printf("  symv floor traffic %.1f GB -> own %.0f GB/s effective\n", bytes * 1e-9, bytes / (ms_own * 1e-3) * 1e-9);

  std::vector<double> hd1(n), he1(n - 1), ht1(n - 1), hd2(n), he2(n - 1), ht2(n - 1);
  CK(cudaMemcpy(hd1.data(), d1, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hd2.data(), d2, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(he1.data(), e1, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(he2.data(), e2, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ht1.data(), t1, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ht2.data(), t2, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost));
  double s;
  double m = maxdiff(hd1, hd2, &s);
  printf("  d   max|diff| %.3e  (scale %.3e)\n", m, s);
  m = maxdiff(he1, he2, &s);
  printf("  e   max|diff| %.3e  (scale %.3e)\n", m, s);
  m = maxdiff(ht1, ht2, &s);
  printf("  tau max|diff| %.3e  (scale %.3e)\n", m, s);
  if (n <= 12000) {
    std::vector<double> h1((size_t)n * n), h2((size_t)n * n);
    CK(cudaMemcpy(h1.data(), A1, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), A2, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    double mm = 0;
    long long bad = 0;
    for (long long j = 0; j < n; ++j)
      for (long long i = j + 1; i < n; ++i) {
        const double df = std::fabs(h1[i + j * n] - h2[i + j * n]);
        if (!(df <= 1e-9)) ++bad;
        if (df > mm) mm = df;
      }
    printf("  reflectors (strict lower) max|diff| %.3e, entries off by > 1e-9: %lld\n", mm, bad);
  }
  return 0;
}
