// Dense symmetric eigensolver used by the SVT (a4), the PCA spectrum (a2) and the Laplacian
// spectrum (a14) -- the dominant cost of the path.
//   4,000 <= n <= 65,535, eigenvectors wanted : native solver -- cooperative-kernel tridiagonalisation
//                 (sytrd_kernel.cuh), tridiagonal divide and conquer (stedc.cu), back-transformation
//                 (cusolverDnDormtr up to 32,768 rows, backtransform.cu above)
//   otherwise, n <= 32,768 : cusolverDnXsyevd (64-bit API; the legacy entry point overflows its int
//                 workspace size above ~23,000 rows)
//   n  > 32,768 : cusolverDnXsyevd rejects the size (probed in scratch/eig_limit_probe.cu);
//                 cusolverMgSyevd on this process's device (scratch/mg_probe.cu: 33,000 rows in 36.8 s
//                 on one B200; the native solver needs 14.9 s) -- only for RSOPTSC_EIG=cusolver,
//                 value-only spectra and blocks above 65,535 rows
#include <cusolverMg.h>
#include <dlfcn.h>

#include "comm.h"
#include "gemm.h"
#include "internal.h"
#include "sytrd_kernel.cuh"
#include "tridiag_bisect.h"

namespace rs {

static int64_t mg_threshold() {
  static int64_t t = -1;
  if (t < 0) {
    const char* e = getenv("RSOPTSC_MG_MIN_N");  // tests lower it to exercise the large-block path
    t = e ? atoll(e) : 32768;
    if (t < 1) t = 1;
    if (t > 32768) t = 32768;
  }
  return t;
}

// libcusolverMg is loaded on demand: it needs the toolkit's own libcublas (12.9), and linking it
// eagerly breaks loading in processes where an older libcublas.so.12 is already resident (e.g. after
// `import torch`, whose wheels bundle 12.8).  Only blocks above 32,768 rows ever get here.
namespace mg {
decltype(&::cusolverMgCreate) Create;
decltype(&::cusolverMgDestroy) Destroy;
decltype(&::cusolverMgDeviceSelect) DeviceSelect;
decltype(&::cusolverMgCreateDeviceGrid) CreateDeviceGrid;
decltype(&::cusolverMgDestroyGrid) DestroyGrid;
decltype(&::cusolverMgCreateMatrixDesc) CreateMatrixDesc;
decltype(&::cusolverMgDestroyMatrixDesc) DestroyMatrixDesc;
decltype(&::cusolverMgSyevd_bufferSize) Syevd_bufferSize;
decltype(&::cusolverMgSyevd) Syevd;
static void load() {
  static bool done = false;
  if (done) return;
  void* h = dlopen("libcusolverMg.so.11", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("/usr/local/cuda/lib64/libcusolverMg.so.11", RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    const char* why = dlerror();
    fail(4, __FILE__, __LINE__, std::string("a connected block above 32,768 cells needs libcusolverMg, which could not be "
                                            "loaded: ") + (why ? why : "?"));
  }
#define RS_MG_SYM(name)                                                                    \
  name = reinterpret_cast<decltype(name)>(dlsym(h, "cusolverMg" #name));                   \
  if (!name) fail(4, __FILE__, __LINE__, "libcusolverMg: missing symbol cusolverMg" #name)
  RS_MG_SYM(Create); RS_MG_SYM(Destroy); RS_MG_SYM(DeviceSelect); RS_MG_SYM(CreateDeviceGrid); RS_MG_SYM(DestroyGrid);
  RS_MG_SYM(CreateMatrixDesc); RS_MG_SYM(DestroyMatrixDesc); RS_MG_SYM(Syevd_bufferSize); RS_MG_SYM(Syevd);
#undef RS_MG_SYM
  done = true;
}
}  // namespace mg
#define cusolverMgCreate mg::Create
#define cusolverMgDestroy mg::Destroy
#define cusolverMgDeviceSelect mg::DeviceSelect
#define cusolverMgCreateDeviceGrid mg::CreateDeviceGrid
#define cusolverMgDestroyGrid mg::DestroyGrid
#define cusolverMgCreateMatrixDesc mg::CreateMatrixDesc
#define cusolverMgDestroyMatrixDesc mg::DestroyMatrixDesc
#define cusolverMgSyevd_bufferSize mg::Syevd_bufferSize
#define cusolverMgSyevd mg::Syevd

static void sym_eig_mg(double* d_A, int64_t n, double* d_lam, bool vectors) {
  Context& c = ctx();
  RS_REQUIRE(n <= 2147483647LL / 2, "sym_eig: block too large");
  mg::load();
  static cusolverMgHandle_t handle = nullptr;
  static int handle_dev = -1;
  if (!handle || handle_dev != c.device) {
    if (handle) cusolverMgDestroy(handle);
    RS_CUSOLVER(cusolverMgCreate(&handle));
    int dev = c.device;
    RS_CUSOLVER(cusolverMgDeviceSelect(handle, 1, &dev));
    handle_dev = c.device;
  }
  const int64_t TB = 256;
  const int64_t cols_pad = (n + TB - 1) / TB * TB;  // the 1-D block-cyclic layout owns whole column tiles
  int32_t dev32 = c.device;
  cudaLibMgGrid_t grid;
  RS_CUSOLVER(cusolverMgCreateDeviceGrid(&grid, 1, 1, &dev32, CUDALIBMG_GRID_MAPPING_COL_MAJOR));
  cudaLibMgMatrixDesc_t desc;
  RS_CUSOLVER(cusolverMgCreateMatrixDesc(&desc, n, n, n, TB, CUDA_R_64F, grid));
  DevBuf<double> Ap((size_t)n * cols_pad);
  RS_CUDA(cudaMemcpyAsync(Ap.p, d_A, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, c.stream));
  if (cols_pad > n)
    RS_CUDA(cudaMemsetAsync(Ap.p + (size_t)n * n, 0, sizeof(double) * n * (cols_pad - n), c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  void* arr_A[1] = {Ap.p};
  std::vector<double> W(n);
  const cusolverEigMode_t mode = vectors ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR;
  int64_t lwork = 0;
  RS_CUSOLVER(cusolverMgSyevd_bufferSize(handle, mode, CUBLAS_FILL_MODE_LOWER, (int)n, arr_A, 1, 1, desc, W.data(),
                                         CUDA_R_64F, CUDA_R_64F, &lwork));
  DevBuf<double> work((size_t)std::max<int64_t>(lwork, 1));
  void* arr_W[1] = {work.p};
  int info = 0;
  cusolverStatus_t st = cusolverMgSyevd(handle, mode, CUBLAS_FILL_MODE_LOWER, (int)n, arr_A, 1, 1, desc, W.data(),
                                        CUDA_R_64F, CUDA_R_64F, arr_W, lwork, &info);
  RS_CUDA(cudaDeviceSynchronize());
  cusolverMgDestroyMatrixDesc(desc);
  cusolverMgDestroyGrid(grid);
  RS_CUSOLVER(st);
  RS_REQUIRE(info == 0, "symmetric eigensolver (Mg) did not converge (info=" + std::to_string(info) + ")");
  if (vectors) RS_CUDA(cudaMemcpyAsync(d_A, Ap.p, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, c.stream));
  RS_CUDA(cudaMemcpyAsync(d_lam, W.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
}

// ---- native solver: own tridiagonalisation (sytrd_kernel.cuh) + own tridiagonal divide and conquer
// (stedc.cu, GEMMs on tcgen05) + own blocked back-transformation (backtransform.cu, GEMMs on tcgen05) ---
void stedc_device(const double* d_d, const double* d_e, int64_t n, double* d_lam, double* d_Z, double* d_S1, double* d_S2,
                  bool collective);
// sytrd_mg.cu: the fused multi-GPU tridiagonalisation (all ranks, replicated matrix in, replicated reflectors out)
bool sytrd_mg_available(int64_t n);
void sytrd_mg_run(double* d_A, int64_t n, double* d_d, double* d_e, double* d_tau);

// RSOPTSC_EIG=cusolver forces the library path, RSOPTSC_EIG=native lowers the size threshold to 256
// (parity tests); default: native from RSOPTSC_NATIVE_MIN_N rows (4000: below that cuSOLVER's lower per-column latency
// wins, measured on B200: 4,000 rows 81 vs 84 ms, 5,000 rows 119 vs 145 ms, 7,954 rows 296 vs 437 ms) up to 65,535.
static int64_t native_min_n() {
  static int64_t t = -2;
  if (t == -2) {
    const char* mode = getenv("RSOPTSC_EIG");
    const char* e = getenv("RSOPTSC_NATIVE_MIN_N");
    if (mode && std::string(mode) == "cusolver") t = -1;
    else if (mode && std::string(mode) == "native") t = 256;
    else t = e ? std::max<int64_t>(atoll(e), 2) : 4000;
  }
  return t;
}

// Blocks below this size go through cusolverDnXsyevd and cuBLAS Grams anyway: run_admm may batch them (svt.cu)
int64_t svt_small_limit() {
  const char* off = getenv("RSOPTSC_SVT_BATCH");
  if (off && off[0] == '0') return 0;
  // the batch route forms its Grams with cuBLAS: stay below the size where the tcgen05 kernel takes them over
  const int64_t gemm_lim = gemm_backend() == 2 ? 256 : 1536;
  const int64_t lim = std::min<int64_t>(native_min_n() < 0 ? gemm_lim : native_min_n(), gemm_lim);
  return std::max<int64_t>(lim, 0);
}

// trailing update of the tridiagonalisation on the library's own GEMM (tcgen05 split-integer kernel, gemm.cu)
// user: null = update the whole block; else int[2] {nranks, rank}: this rank's 128-row groups only (sytrd_mg.cu)
void sytrd_trailing_update(const double* V, long long ldv, const double* W, long long ldw, long long nr, int k, double* C,
                           long long ldc, long long r0, void* user) {
  const int* sh = static_cast<const int*>(user);
  syr2k_lower_sub(V, ldv, W, ldw, nr, k, C, ldc, r0, sh ? sh[0] : 1, sh ? sh[1] : 0);
  if (sh) {
    // sharded update: every row of the trailing block is current on its owner only.  The next panel's kernel reads
    // its 64 columns in full on every rank, so those columns are gathered (the diagonal entry of the last row too)
    double* A = C - r0 - r0 * ldc;
    const long long n = r0 + nr;
    comm_gather_cyclic_rows(A, ldc, n, r0, (int)std::min<long long>(sytrd::NB, nr), 2 * sytrd::TILE);
  }
}

static void sym_eig_native(double* d_A, int64_t n, double* d_lam, double keep_above = -1.0e300, bool collective = false) {
  Context& c = ctx();
  collective = collective && sharded();
  DevBuf<double> d(n), e(n), tau(n), Z((size_t)n * n);
  if (collective && sytrd_mg_available(n)) {
    Phase ph("eig_sytrd");
    sytrd_mg_run(d_A, n, d.p, e.p, tau.p);
  } else {
    Phase ph("eig_sytrd");
    DevBuf<double> work(sytrd::workspace_doubles(n));
    int64_t launches = 0;
    sytrd::Hooks hooks;
    Phase* panel_phase = nullptr;
    hooks.user = &panel_phase;
    hooks.syr2k = sytrd_trailing_update;
    hooks.before = [](void* u, double bytes) {
      add_counter("sytrd_panel_bytes", bytes);
      *static_cast<Phase**>(u) = new Phase("sytrd_panel");
    };
    hooks.after = [](void* u) {
      delete *static_cast<Phase**>(u);
      *static_cast<Phase**>(u) = nullptr;
    };
    RS_CUDA(sytrd::run(c.cublas, c.stream, d_A, n, n, d.p, e.p, tau.p, work.p, &launches, nullptr, &hooks));
    g_launches.fetch_add(launches, std::memory_order_relaxed);
  }
  {
    Phase ph("eig_stedc");
    DevBuf<double> S1((size_t)n * n), S2((size_t)n * n);
    stedc_device(d.p, e.p, n, d_lam, Z.p, S1.p, S2.p, collective);
  }
  // eigenvalues ascending: the wanted eigenvectors are the trailing columns c0 .. n-1
  int64_t c0 = 0;
  if (keep_above > -1.0e299) {
    std::vector<double> lam(n);
    RS_CUDA(cudaMemcpyAsync(lam.data(), d_lam, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    RS_CUDA(cudaStreamSynchronize(c.stream));
    while (c0 < n && lam[c0] < keep_above) ++c0;
  }
  const int64_t ncols = n - c0;
  double* Zk = Z.p + (size_t)c0 * n;
  if (ncols > 0 && collective) {
    // every rank back-transforms one panel of the wanted eigenvectors (own blocked compact-WY routine, GEMMs on
    // tcgen05), the panels are all-gathered
    Phase ph("eig_backtransform");
    const std::vector<int64_t> zb = split_even(ncols, comm().nranks, 128);
    const int64_t z0 = zb[comm().rank], z1 = zb[comm().rank + 1];
    if (z1 > z0) apply_q_left(d_A, n, tau.p, n, Zk + (size_t)z0 * n, z1 - z0);
    comm_allgather_cols(Zk, n, zb);
    RS_CUDA(cudaMemcpyAsync(d_A + (size_t)c0 * n, Zk, sizeof(double) * (size_t)n * ncols, cudaMemcpyDeviceToDevice, c.stream));
    RS_CUDA(cudaStreamSynchronize(c.stream));
  } else if (ncols > 0) {
    Phase ph("eig_backtransform");
    // Own blocked compact-WY back-transformation (backtransform.cu, panel GEMMs on tcgen05): 41 ms at 7,954 rows
    // against cusolverDnDormtr's 47 ms since the GEMM epilogue keeps its loads in flight, and no 32-bit workspace
    // limit.  RSOPTSC_BACKTRANSFORM=cusolver keeps the library routine for comparison (tests).
    static int use_lib = -1;
    if (use_lib < 0) {
      const char* e = getenv("RSOPTSC_BACKTRANSFORM");
      use_lib = (e && std::string(e) == "cusolver") ? 1 : 0;
    }
    if (use_lib && n <= 32768) {
      int lwork = 0;
      RS_CUSOLVER(cusolverDnDormtr_bufferSize(c.cusolver, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)n, (int)ncols,
                                              d_A, (int)n, tau.p, Zk, (int)n, &lwork));
      DevBuf<double> work((size_t)std::max(lwork, 1));
      DevBuf<int> info(1);
      RS_CUSOLVER(cusolverDnDormtr(c.cusolver, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)n, (int)ncols, d_A,
                                   (int)n, tau.p, Zk, (int)n, work.p, lwork, info.p));
      int hinfo = 0;
      RS_CUDA(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
      RS_CUDA(cudaStreamSynchronize(c.stream));
      RS_REQUIRE(hinfo == 0, "back-transformation failed (info=" + std::to_string(hinfo) + ")");
    } else {
      apply_q_left(d_A, n, tau.p, n, Zk, ncols);
    }
    RS_CUDA(cudaMemcpyAsync(d_A + (size_t)c0 * n, Zk, sizeof(double) * (size_t)n * ncols, cudaMemcpyDeviceToDevice, c.stream));
    RS_CUDA(cudaStreamSynchronize(c.stream));
  }
}

// ---- values only: own tridiagonalisation + bisection on Sturm counts (tridiag_bisect.h), one thread per eigenvalue ---
// scal: [0] gl, [1] gu (Gershgorin interval, widened as in dstebz), [2] pivmin
__global__ void k_bisect_prep(const double* __restrict__ d, const double* __restrict__ e, int n, double* __restrict__ e2,
                              double* __restrict__ scal) {
  __shared__ double s_lo[32], s_hi[32], s_e[32];
  double lo = INFINITY, hi = -INFINITY, em = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i + 1 < n ? fabs(e[i]) : 0.0;
    lo = fmin(lo, d[i] - el - er);
    hi = fmax(hi, d[i] + el + er);
    if (i + 1 < n) { const double x = e[i] * e[i]; e2[i] = x; em = fmax(em, x); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    em = fmax(em, __shfl_xor_sync(0xffffffffu, em, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; s_e[w] = em; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned k = 1; k < blockDim.x / 32; ++k) { lo = fmin(lo, s_lo[k]); hi = fmax(hi, s_hi[k]); em = fmax(em, s_e[k]); }
    const double pivmin = DBL_MIN * fmax(1.0, em);
    const double tnorm = fmax(fabs(lo), fabs(hi));
    scal[0] = lo - 2.0 * tnorm * DBL_EPSILON * n - 2.0 * pivmin;
    scal[1] = hi + 2.0 * tnorm * DBL_EPSILON * n + 2.0 * pivmin;
    scal[2] = pivmin;
  }
}
__global__ void k_bisect(const double* __restrict__ d, const double* __restrict__ e2, int n, const double* __restrict__ scal,
                         double* __restrict__ lam) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) lam[k] = bisect::kth_eigenvalue(d, e2, n, k, scal[0], scal[1], scal[2]);
}
void tridiag_eigenvalues_device(const double* d_d, const double* d_e, int64_t n, double* d_lam) {
  DevBuf<double> e2(std::max<int64_t>(n, 1)), scal(4);
  RS_LAUNCH(k_bisect_prep, 1, 1024, 0, d_d, d_e, (int)n, e2.p, scal.p);
  RS_LAUNCH(k_bisect, (unsigned)cdiv(n, 64), 64, 0, d_d, e2.p, (int)n, scal.p, d_lam);  // 64 threads/CTA: spread over the SMs
}

static void sym_eigvals_native(double* d_A, int64_t n, double* d_lam) {
  Context& c = ctx();
  if (n == 1) {
    RS_CUDA(cudaMemcpyAsync(d_lam, d_A, sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    return;
  }
  DevBuf<double> d(n), e(n), tau(n);
  {
    Phase ph("eig_sytrd");
    DevBuf<double> work(sytrd::workspace_doubles(n));
    int64_t launches = 0;
    sytrd::Hooks hooks;
    hooks.syr2k = sytrd_trailing_update;
    RS_CUDA(sytrd::run(c.cublas, c.stream, d_A, n, n, d.p, e.p, tau.p, work.p, &launches, nullptr, &hooks));
    g_launches.fetch_add(launches, std::memory_order_relaxed);
  }
  Phase ph("eig_bisect");
  tridiag_eigenvalues_device(d.p, e.p, n, d_lam);
}

static void sym_eig_library(double* d_A, int64_t n, double* d_lam, bool vectors);

void sym_eig_select(double* d_A, int64_t n, double* d_lam, int solver) {
  if (solver == 1) return sym_eig_native(d_A, n, d_lam);
  sym_eig_library(d_A, n, d_lam, true);
}

// A (n x n, lower triangle referenced) -> eigenvectors in A (when `vectors`), eigenvalues ascending
// in d_lam.  Throws when the solver reports failure.
void sym_eig(double* d_A, int64_t n, double* d_lam, bool vectors, double keep_above, bool collective) {
  RS_REQUIRE(n >= 1, "sym_eig: empty matrix");
  // collective calls always take the native solver: its stages are the ones that are split across the ranks
  if (vectors && n <= 65535 && ((collective && sharded() && n >= 2) || (native_min_n() >= 0 && n >= native_min_n())))
    return sym_eig_native(d_A, n, d_lam, keep_above, collective);
  // spectra without vectors (PCA a2, Laplacian a14): own tridiagonalisation + bisection, any size up to 65,535
  if (!vectors && native_min_n() >= 0 && n <= 65535) return sym_eigvals_native(d_A, n, d_lam);
  sym_eig_library(d_A, n, d_lam, vectors);
}

static void sym_eig_library(double* d_A, int64_t n, double* d_lam, bool vectors) {
  Context& c = ctx();
  if (n > mg_threshold()) return sym_eig_mg(d_A, n, d_lam, vectors);
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
