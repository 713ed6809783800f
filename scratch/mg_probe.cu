// Does cusolverMgSyevd (1 device) accept n > 32768, and how fast is it? (scratch probe)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverMg.h>
#define CK(x) do{auto e=(x); if(e!=0){printf("ERR %d at line %d\n",(int)e,__LINE__); return 1;} }while(0)
// local buffer of device d in the 1-D column block-cyclic layout (tile width TB): local column lc -> global column
__global__ void fill(double* a, size_t ld, size_t lcols, int d, int ndev, int TB){ size_t n=ld*lcols; size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){ size_t r=i%ld,lc=i/ld; size_t c=((lc/TB)*ndev+d)*TB + lc%TB; size_t k = r<c? c*ld+r : r*ld+c; unsigned x=(unsigned)(k*2654435761u); x^=x>>13; x*=0x5bd1e995u; x^=x>>15; a[i]=((double)(x&0xffff)/65536.0-0.5) + (r==c? 4.0:0.0);} }
int main(int argc,char**argv){
  int n = argc>1? atoi(argv[1]) : 8000;
  int ndev = argc>2? atoi(argv[2]) : 1;
  const int TB = 256;
  cusolverMgHandle_t h; CK(cusolverMgCreate(&h));
  std::vector<int> devs(ndev); for(int i=0;i<ndev;i++) devs[i]=i;
  CK(cusolverMgDeviceSelect(h, ndev, devs.data()));
  cudaLibMgGrid_t grid; CK(cusolverMgCreateDeviceGrid(&grid, 1, ndev, devs.data(), CUDALIBMG_GRID_MAPPING_COL_MAJOR));
  cudaLibMgMatrixDesc_t desc; CK(cusolverMgCreateMatrixDesc(&desc, n, n, n, TB, CUDA_R_64F, grid));
  // 1-D column block-cyclic distribution: with one device the local matrix is the whole matrix
  std::vector<void*> dA(ndev);
  int nblk = (n + TB - 1)/TB;
  for(int d=0; d<ndev; d++){ int myblk = (nblk - d + ndev - 1)/ndev; size_t cols=(size_t)myblk*TB; cudaSetDevice(d); if(cudaMalloc(&dA[d], (size_t)n*cols*8)!=cudaSuccess){printf("malloc fail\n"); return 1;}
    fill<<<2048,256>>>((double*)dA[d],(size_t)n,cols,d,ndev,TB); cudaDeviceSynchronize(); }
  std::vector<double> W(n);
  int64_t lwork=0; cusolverStatus_t st = cusolverMgSyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA.data(), 1, 1, desc, W.data(), CUDA_R_64F, CUDA_R_64F, &lwork);
  printf("n=%d ndev=%d MgSyevd_bufferSize status %d lwork %.2f GB per device\n", n, ndev, (int)st, lwork*8/1e9);
  if(st!=0) return 0;
  std::vector<void*> dW(ndev); for(int d=0; d<ndev; d++){ cudaSetDevice(d); if(cudaMalloc(&dW[d], (size_t)lwork*8)!=cudaSuccess){printf("work malloc fail\n"); return 1;} }
  int info=0; cudaSetDevice(0);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  cudaEventRecord(e0);
  st = cusolverMgSyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA.data(), 1, 1, desc, W.data(), CUDA_R_64F, CUDA_R_64F, dW.data(), lwork, &info);
  cudaDeviceSynchronize(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
  printf("n=%d MgSyevd status %d info %d  %.1f ms  W[0]=%.4f W[n-1]=%.4f\n", n,(int)st,info,ms,W[0],W[n-1]);
  return 0;
}
