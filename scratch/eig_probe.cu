// Probe: legacy Dsyevd vs 64-bit Xsyevd vs Dsyevdx (range) vs Dsyevj on a Gram matrix (scratch tool)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do{auto e=(x); if(e!=0){printf("ERR %d at %s:%d\n",(int)e,__FILE__,__LINE__); exit(1);} }while(0)
__global__ void fill(double* a, size_t n, unsigned seed){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){ unsigned x=(unsigned)(i*2654435761u)^seed; x^=x>>13; x*=0x5bd1e995u; x^=x>>15; a[i]=((double)(x&0xffffff)/16777216.0-0.5);} }
int main(int argc,char**argv){
  int n = argc>1? atoi(argv[1]) : 8000;
  cublasHandle_t cb; cusolverDnHandle_t cs; CK(cublasCreate(&cb)); CK(cusolverDnCreate(&cs));
  double *A,*G,*W; size_t nn=(size_t)n*n; CK(cudaMalloc(&A,nn*8)); CK(cudaMalloc(&G,nn*8)); CK(cudaMalloc(&W,n*8));
  fill<<<1024,256>>>(A,nn,1u);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms; double one=1,zero=0; int* info; CK(cudaMalloc(&info,4));
  auto gram=[&]{ CK(cublasDsyrk(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_T,n,n,&one,A,n,&zero,G,n)); };
  { int lwork=0; CK(cusolverDnDsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,W,&lwork)); double* work; CK(cudaMalloc(&work,(size_t)lwork*8));
    for(int r=0;r<2;r++){ gram(); cudaEventRecord(e0); CK(cusolverDnDsyevd(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,W,work,lwork,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("n=%d Dsyevd %.1f ms\n",n,ms);} cudaFree(work); }
  { cusolverDnParams_t par; CK(cusolverDnCreateParams(&par)); size_t wd=0,wh=0; CK(cusolverDnXsyevd_bufferSize(cs,par,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,CUDA_R_64F,G,n,CUDA_R_64F,W,CUDA_R_64F,&wd,&wh)); void* dw; CK(cudaMalloc(&dw,wd)); std::vector<char> hw(wh+16);
    for(int r=0;r<2;r++){ gram(); cudaEventRecord(e0); CK(cusolverDnXsyevd(cs,par,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,CUDA_R_64F,G,n,CUDA_R_64F,W,CUDA_R_64F,dw,wd,hw.data(),wh,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("n=%d Xsyevd %.1f ms\n",n,ms);} cudaFree(dw); }
  { int lwork=0, found=0; CK(cusolverDnDsyevdx_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUSOLVER_EIG_RANGE_I,CUBLAS_FILL_MODE_LOWER,n,G,n,0,0,n/2+1,n,&found,W,&lwork)); double* work; CK(cudaMalloc(&work,(size_t)lwork*8));
    for(int r=0;r<2;r++){ gram(); cudaEventRecord(e0); CK(cusolverDnDsyevdx(cs,CUSOLVER_EIG_MODE_VECTOR,CUSOLVER_EIG_RANGE_I,CUBLAS_FILL_MODE_LOWER,n,G,n,0,0,n/2+1,n,&found,W,work,lwork,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("n=%d Dsyevdx(top half) %.1f ms found %d\n",n,ms,found);} cudaFree(work); }
  { int lwork=0; double* tau; CK(cudaMalloc(&tau,n*8)); double *D,*E; CK(cudaMalloc(&D,n*8)); CK(cudaMalloc(&E,n*8)); CK(cusolverDnDsytrd_bufferSize(cs,CUBLAS_FILL_MODE_LOWER,n,G,n,D,E,tau,&lwork)); double* work; CK(cudaMalloc(&work,(size_t)lwork*8));
    for(int r=0;r<2;r++){ gram(); cudaEventRecord(e0); CK(cusolverDnDsytrd(cs,CUBLAS_FILL_MODE_LOWER,n,G,n,D,E,tau,work,lwork,info)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("n=%d Dsytrd alone %.1f ms\n",n,ms);} }
  return 0;
}
