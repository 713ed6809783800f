// Which n does cusolverDnXsyevd accept? (buffer-size query only; scratch tool)
#include <cstdio>
#include <cuda_runtime.h>
#include <cusolverDn.h>
int main(){
  cusolverDnHandle_t cs; cusolverDnCreate(&cs); cusolverDnParams_t par; cusolverDnCreateParams(&par);
  double* dummy; cudaMalloc(&dummy, 1024);
  long ns[] = {24000, 30000, 32000, 32767, 32768, 33000, 36000, 40000, 46340, 46341, 50000, 60000};
  for (long n : ns) {
    size_t wd=0, wh=0;
    cusolverStatus_t st = cusolverDnXsyevd_bufferSize(cs, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dummy, n, CUDA_R_64F, dummy, CUDA_R_64F, &wd, &wh);
    int lw=0; cusolverStatus_t st2 = cusolverDnDsyevd_bufferSize(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, dummy, (int)n, dummy, &lw);
    printf("n=%ld Xsyevd status %d dev %.2f GB host %.2f MB | legacy status %d lwork %d\n", n, (int)st, wd/1e9, wh/1e6, (int)st2, lw);
  }
  return 0;
}
