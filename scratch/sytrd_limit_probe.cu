// Do the legacy Dsytrd / Dormtr entry points accept n > 32768? (scratch probe)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cusolverDn.h>
__global__ void fill(double* a, size_t n, size_t ld){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){ size_t r=i%ld,c=i/ld; unsigned x=(unsigned)((r<c? c*ld+r : i)*2654435761u); x^=x>>13; x*=0x5bd1e995u; x^=x>>15; a[i]=((double)(x&0xffff)/65536.0-0.5) + (r==c? 4.0:0.0);} }
int main(int argc,char**argv){
  long n = argc>1? atol(argv[1]) : 34000;
  cusolverDnHandle_t cs; cusolverDnCreate(&cs);
  double *A,*D,*E,*tau; size_t nn=(size_t)n*n; if(cudaMalloc(&A,nn*8)!=cudaSuccess){printf("malloc fail\n");return 1;} cudaMalloc(&D,n*8); cudaMalloc(&E,n*8); cudaMalloc(&tau,n*8);
  fill<<<2048,256>>>(A,nn,(size_t)n);
  int lwork=0; cusolverStatus_t st = cusolverDnDsytrd_bufferSize(cs,CUBLAS_FILL_MODE_LOWER,(int)n,A,(int)n,D,E,tau,&lwork);
  printf("n=%ld sytrd_bufferSize status %d lwork %d\n", n,(int)st,lwork);
  if(st!=0) return 0;
  double* work; cudaMalloc(&work,(size_t)lwork*8); int* info; cudaMalloc(&info,4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  cudaEventRecord(e0); st = cusolverDnDsytrd(cs,CUBLAS_FILL_MODE_LOWER,(int)n,A,(int)n,D,E,tau,work,lwork,info); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
  int hi=-1; cudaMemcpy(&hi,info,4,cudaMemcpyDeviceToHost); printf("n=%ld Dsytrd status %d info %d  %.1f ms  (%s)\n",n,(int)st,hi,ms,cudaGetErrorString(cudaGetLastError()));
  // ormtr: apply Q to a thin matrix C (n x 64)
  double* C; cudaMalloc(&C,(size_t)n*64*8); cudaMemset(C,0,(size_t)n*64*8); int lw2=0;
  st = cusolverDnDormtr_bufferSize(cs,CUBLAS_SIDE_LEFT,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_N,(int)n,64,A,(int)n,tau,C,(int)n,&lw2);
  printf("ormtr_bufferSize status %d lwork %d\n",(int)st,lw2);
  if(st==0){ double* w2; cudaMalloc(&w2,(size_t)lw2*8); cudaEventRecord(e0); st=cusolverDnDormtr(cs,CUBLAS_SIDE_LEFT,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_N,(int)n,64,A,(int)n,tau,C,(int)n,w2,lw2,info); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); cudaMemcpy(&hi,info,4,cudaMemcpyDeviceToHost); printf("Dormtr(n x 64) status %d info %d %.1f ms\n",(int)st,hi,ms);}
  return 0;
}
