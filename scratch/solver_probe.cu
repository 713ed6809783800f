// Probe: how fast are the candidate dense FP64 building blocks on this B200?
// (scratch tool, not part of the product)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do{auto e=(x); if(e!=0){printf("ERR %d at %s:%d\n",(int)e,__FILE__,__LINE__); exit(1);} }while(0)
__global__ void fill(double* a, size_t n, unsigned seed){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){ unsigned x=(unsigned)(i*2654435761u)^seed; x^=x>>13; x*=0x5bd1e995u; x^=x>>15; a[i]=((double)(x&0xffffff)/16777216.0-0.5);} }
int main(int argc,char**argv){
  int n = argc>1? atoi(argv[1]) : 4096;
  int do_svdp = argc>2? atoi(argv[2]) : 1;
  cublasHandle_t cb; cusolverDnHandle_t cs; CK(cublasCreate(&cb)); CK(cusolverDnCreate(&cs));
  double *A,*G,*W,*C; size_t nn=(size_t)n*n;
  CK(cudaMalloc(&A,nn*8)); CK(cudaMalloc(&G,nn*8)); CK(cudaMalloc(&C,nn*8)); CK(cudaMalloc(&W,n*8));
  fill<<<1024,256>>>(A,nn,1u);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  double one=1,zero=0;
  for(int rep=0;rep<2;rep++){
    cudaEventRecord(e0); CK(cublasDsyrk(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,n,n,&one,A,n,&zero,G,n)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("n=%d dsyrk %.2f ms  %.1f TF/s (n^3)\n",n,ms,1.0*n*n*n/ms/1e9);
    cudaEventRecord(e0); CK(cublasDgemm(cb,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,A,n,&zero,C,n)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("n=%d dgemm %.2f ms  %.1f TF/s\n",n,ms,2.0*n*n*n/ms/1e9);
  }
  int lwork=0; CK(cusolverDnDsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,W,&lwork));
  double* work; CK(cudaMalloc(&work,(size_t)lwork*8)); int* info; CK(cudaMalloc(&info,4));
  printf("syevd lwork %.1f MB\n", lwork*8.0/1e6);
  for(int rep=0;rep<2;rep++){
    CK(cublasDsyrk(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,n,n,&one,A,n,&zero,G,n));
    cudaEventRecord(e0); CK(cusolverDnDsyevd(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,W,work,lwork,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    int hi; cudaMemcpy(&hi,info,4,cudaMemcpyDeviceToHost); printf("n=%d dsyevd(V) %.1f ms info %d\n",n,ms,hi);
  }
  { // values only
    CK(cublasDsyrk(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,n,n,&one,A,n,&zero,G,n));
    cudaEventRecord(e0); CK(cusolverDnDsyevd(cs,CUSOLVER_EIG_MODE_NOVECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,W,work,lwork,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("n=%d dsyevd(noV) %.1f ms\n",n,ms);
  }
  cudaFree(work);
  if(do_svdp){
    cusolverDnParams_t par; CK(cusolverDnCreateParams(&par));
    double *U,*V,*S; CK(cudaMalloc(&U,nn*8)); CK(cudaMalloc(&V,nn*8)); CK(cudaMalloc(&S,n*8));
    size_t wd=0,wh=0; CK(cusolverDnXgesvdp_bufferSize(cs,par,CUSOLVER_EIG_MODE_VECTOR,0,n,n,CUDA_R_64F,C,n,CUDA_R_64F,S,CUDA_R_64F,U,n,CUDA_R_64F,V,n,CUDA_R_64F,&wd,&wh));
    void* dw; CK(cudaMalloc(&dw,wd)); std::vector<char> hw(wh+16); double herr=0;
    printf("gesvdp work %.1f MB dev, %.1f MB host\n", wd/1e6, wh/1e6);
    for(int rep=0;rep<2;rep++){
      CK(cudaMemcpy(C,A,nn*8,cudaMemcpyDeviceToDevice));
      cudaEventRecord(e0); CK(cusolverDnXgesvdp(cs,par,CUSOLVER_EIG_MODE_VECTOR,0,n,n,CUDA_R_64F,C,n,CUDA_R_64F,S,CUDA_R_64F,U,n,CUDA_R_64F,V,n,CUDA_R_64F,dw,wd,hw.data(),wh,info,&herr)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
      printf("n=%d gesvdp %.1f ms herr %.2e\n",n,ms,herr);
    }
  }
  // float syevd for comparison
  { float *Gf,*Wf; CK(cudaMalloc(&Gf,nn*4)); CK(cudaMalloc(&Wf,n*4)); cudaMemset(Gf,0,nn*4);
    // build an SPD-ish float matrix: Gf = (float)G lower
    CK(cublasDsyrk(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,n,n,&one,A,n,&zero,G,n));
    std::vector<double> tmp; // convert on device via simple kernel is overkill: use cublas? just reinterpret: run on garbage-free zeros + identity
    int lw=0; CK(cusolverDnSsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,Gf,n,Wf,&lw)); float* wk; CK(cudaMalloc(&wk,(size_t)lw*4));
    cudaEventRecord(e0); CK(cusolverDnSsyevd(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,Gf,n,Wf,wk,lw,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("n=%d ssyevd(V, zero matrix: lower bound only) %.1f ms\n",n,ms);
  }
  return 0;
}
