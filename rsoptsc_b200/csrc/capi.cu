// extern "C" boundary of librsoptsc_b200.so (declared in include/rsoptsc_b200.h).
// Every entry point converts C++ exceptions into a status code + thread-local message.
#include <mutex>
#include <unordered_map>

#include "../../include/rsoptsc_b200.h"
#include "comm.h"
#include "gemm.h"
#include "internal.h"
#include "tridiag_bisect.h"

namespace rs {

std::atomic<int64_t> g_launches{0};
static Context g_ctx;
static thread_local std::string g_err;
void release_dense_scratch();  // dense.cu
void sytrd_mg_tables_host(int64_t n, int nranks, int rank, std::vector<int>& ownI, std::vector<long long>& pref,
                          std::vector<int>& lfirst);  // sytrd_mg.cu

void ensure_init(int dev);
Context& ctx() {
  if (!g_ctx.initialized) ensure_init(-1);  // lazy: current device; fails loudly without a GPU
  return g_ctx;
}

void ensure_init(int dev) {
  if (g_ctx.initialized && (dev < 0 || dev == g_ctx.device)) return;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail(6, __FILE__, __LINE__, std::string("no CUDA device available: the rsoptsc_b200 path has no CPU fallback (") +
                                    cudaGetErrorString(e) + ")");
  if (g_ctx.initialized) {  // switching device: cached blocks belong to the old one
    release_dense_scratch();
    pool_trim();
    cublasDestroy(g_ctx.cublas); cusolverDnDestroy(g_ctx.cusolver); cudaStreamDestroy(g_ctx.stream);
    g_ctx = Context();
  }
  if (dev < 0) RS_CUDA(cudaGetDevice(&dev));
  RS_REQUIRE(dev < count, "device index out of range");
  RS_CUDA(cudaSetDevice(dev));
  g_ctx.device = dev;
  RS_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  RS_CUBLAS(cublasCreate(&g_ctx.cublas));
  RS_CUBLAS(cublasSetStream(g_ctx.cublas, g_ctx.stream));
  RS_CUSOLVER(cusolverDnCreate(&g_ctx.cusolver));
  RS_CUSOLVER(cusolverDnSetStream(g_ctx.cusolver, g_ctx.stream));
  cudaDeviceProp prop;
  RS_CUDA(cudaGetDeviceProperties(&prop, dev));
  g_ctx.sm_count = prop.multiProcessorCount;
  g_ctx.initialized = true;
}

// ---- caching device allocator ---------------------------------------------------------
// Buckets keep 3 significant bits of the size (<= 12.5 % slack).  All library work is issued on
// one stream, so a recycled buffer is only touched after the kernels that last used it.
static std::map<size_t, std::vector<void*>>& pool() {  // leaked on purpose: static DevBufs may release at exit
  static auto* p = new std::map<size_t, std::vector<void*>>();
  return *p;
}
#define g_pool pool()
static size_t g_pooled_bytes = 0;                  // bytes currently idle in the cache
static size_t bucket_of(size_t bytes) {
  if (bytes <= 512) return 512;
  int hb = 63 - __builtin_clzll((unsigned long long)bytes);
  const size_t step = (size_t)1 << (hb - 3);
  return (bytes + step - 1) / step * step;
}
// Blocks remember the bucket they were allocated with: a request may be served by a cached block of a LARGER bucket
// (big requests only, at most 1.5x), so that a sequence of slowly shrinking multi-GB requests -- the int8 slice buffer
// of the Z panel in the back-transformation, one size per 1,024-reflector block -- reuses one block instead of going
// through cudaMalloc / cudaFree for every new size.
static std::unordered_map<void*, size_t>& block_sizes() {
  static auto* m = new std::unordered_map<void*, size_t>();
  return *m;
}
static constexpr size_t POOL_FUZZY_MIN = (size_t)64 << 20;
void* pool_alloc(size_t bytes) {
  const size_t b = bucket_of(bytes);
  auto it = g_pool.find(b);
  if ((it == g_pool.end() || it->second.empty()) && b >= POOL_FUZZY_MIN) {
    for (it = g_pool.lower_bound(b); it != g_pool.end() && it->first <= b + b / 2; ++it)
      if (!it->second.empty()) break;
    if (it != g_pool.end() && it->first > b + b / 2) it = g_pool.end();
  }
  if (it != g_pool.end() && !it->second.empty()) {
    void* p = it->second.back();
    it->second.pop_back();
    g_pooled_bytes -= std::min(g_pooled_bytes, it->first);
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, b);
  if (e != cudaSuccess) {  // out of memory: give the cached blocks back and retry once
    cudaGetLastError();
    pool_trim();
    e = cudaMalloc(&p, b);
  }
  if (e != cudaSuccess)
    fail(2, __FILE__, __LINE__, std::string("CUDA: cudaMalloc of ") + std::to_string(b) + " bytes failed: " + cudaGetErrorString(e));
  block_sizes()[p] = b;
  return p;
}
static size_t pool_limit() {                       // idle bytes above which the cache is handed back
  static size_t lim = 0;
  if (!lim) {
    const char* e = getenv("RSOPTSC_POOL_LIMIT_GB");
    lim = (size_t)((e ? atof(e) : 48.0) * (double)(1ull << 30));
  }
  return lim;
}
void pool_free(void* p, size_t bytes) {
  if (!p) return;
  auto known = block_sizes().find(p);
  const size_t b = known != block_sizes().end() ? known->second : bucket_of(bytes);  // the block's own bucket
  g_pool[b].push_back(p);
  g_pooled_bytes += b;
  if (g_pooled_bytes > pool_limit()) pool_trim();
}
void pool_trim() {
  for (auto& kv : g_pool)
    for (void* p : kv.second) {
      cudaFree(p);
      block_sizes().erase(p);
    }
  g_pool.clear();
  g_pooled_bytes = 0;
}

// ---- phase timers ---------------------------------------------------------------------
struct PhaseRec { double ms = 0; int64_t calls = 0; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending; };
static std::vector<std::string> g_phase_names;
static std::vector<PhaseRec> g_phases;
static bool g_timers_on = false;

static int phase_id(const char* name) {
  for (size_t i = 0; i < g_phase_names.size(); ++i)
    if (g_phase_names[i] == name) return (int)i;
  g_phase_names.emplace_back(name);
  g_phases.emplace_back();
  return (int)g_phase_names.size() - 1;
}
Phase::Phase(const char* name) : id(-1) {
  if (!g_timers_on || !g_ctx.initialized) return;
  id = phase_id(name);
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, g_ctx.stream);
}
Phase::~Phase() {
  if (id < 0 || !g_ctx.initialized) return;
  cudaEventRecord(e1, g_ctx.stream);
  g_phases[id].pending.emplace_back(e0, e1);
  g_phases[id].calls++;
}
static std::unordered_map<std::string, cudaEvent_t> g_open_timers;
static std::unordered_map<std::string, double> g_counters;
void add_counter(const char* name, double v) { g_counters[name] += v; }
static void resolve_phases() {
  if (!g_ctx.initialized) return;
  cudaStreamSynchronize(g_ctx.stream);
  for (auto& p : g_phases) {
    for (auto& ev : p.pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) p.ms += ms;
      cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    p.pending.clear();
  }
}

void release_dense_scratch();

static rsoptsc_progress_fn g_progress = nullptr;
static void* g_progress_user = nullptr;
void report_progress(int stage, int iter, double value) {
  if (g_progress && g_progress(stage, iter, value, g_progress_user) != 0)
    fail(RSOPTSC_STATUS_INTERRUPTED, __FILE__, __LINE__, "interrupted by the caller's progress callback");
}

// One caller at a time: the context, the caching allocator, the phase tables and the per-function statics are
// shared, unsynchronised state (an R host is single-threaded anyway; Python hosts with threads serialise here).
// The process's current device is re-asserted on every entry: a host such as torch may have switched it.
static std::recursive_mutex g_api_mutex;
template <class F>
static int guarded(F&& f) {
  std::lock_guard<std::recursive_mutex> lock(g_api_mutex);
  try {
    if (g_ctx.initialized) cudaSetDevice(g_ctx.device);
    f();
    return 0;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code ? e.code : 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  } catch (...) {
    g_err = "unknown error";
    return 1;
  }
}

}  // namespace rs

using namespace rs;

extern "C" {

int rsoptsc_abi_version(void) { return RSOPTSC_ABI_VERSION; }
const char* rsoptsc_last_error(void) { return g_err.c_str(); }
int rsoptsc_init(int device) { return guarded([&] { ensure_init(device); }); }
int rsoptsc_finalize(void) {
  return guarded([&] {
    if (!g_ctx.initialized) return;
    comm_finalize();
    resolve_phases();
    release_dense_scratch();
    pool_trim();
    cublasDestroy(g_ctx.cublas); cusolverDnDestroy(g_ctx.cusolver); cudaStreamDestroy(g_ctx.stream);
    g_ctx = Context();
  });
}
int64_t rsoptsc_kernel_launches(void) { return g_launches.load(); }
int rsoptsc_set_progress(rsoptsc_progress_fn fn, void* user) {
  g_progress = fn;
  g_progress_user = user;
  return 0;
}
int rsoptsc_timers_enable(int on) { g_timers_on = on != 0; return 0; }
int rsoptsc_timers_reset(void) {
  resolve_phases();
  for (auto& p : g_phases) { p.ms = 0; p.calls = 0; }
  g_counters.clear();
  return 0;
}
double rsoptsc_counter(const char* name) {
  auto it = g_counters.find(name);
  return it == g_counters.end() ? 0.0 : it->second;
}
double rsoptsc_phase_ms(const char* name) {
  resolve_phases();
  for (size_t i = 0; i < g_phase_names.size(); ++i)
    if (g_phase_names[i] == name) return g_phases[i].ms;
  return -1.0;
}
const char* rsoptsc_phase_name(int32_t i) {
  return (i >= 0 && (size_t)i < g_phase_names.size()) ? g_phase_names[i].c_str() : nullptr;
}
int rsoptsc_timer_begin(const char* name) {
  return guarded([&] {
    cudaEvent_t e;
    RS_CUDA(cudaEventCreate(&e));
    RS_CUDA(cudaEventRecord(e, ctx().stream));
    g_open_timers[name] = e;
  });
}
int rsoptsc_timer_end(const char* name) {
  return guarded([&] {
    auto it = g_open_timers.find(name);
    RS_REQUIRE(it != g_open_timers.end(), "timer_end without timer_begin");
    cudaEvent_t e1;
    RS_CUDA(cudaEventCreate(&e1));
    RS_CUDA(cudaEventRecord(e1, ctx().stream));
    const int id = phase_id(name);
    g_phases[id].pending.emplace_back(it->second, e1);
    g_phases[id].calls++;
    g_open_timers.erase(it);
  });
}
int64_t rsoptsc_phase_calls(const char* name) {
  for (size_t i = 0; i < g_phase_names.size(); ++i)
    if (g_phase_names[i] == name) return g_phases[i].calls;
  return -1;
}

int rsoptsc_comm_unique_id(uint8_t* id128) {
  return guarded([&] {
    RS_REQUIRE(id128 != nullptr, "comm_unique_id: null buffer");
    comm_unique_id(id128);
  });
}
int rsoptsc_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128) {
  return guarded([&] { comm_init(nranks, rank, id128); });
}
int rsoptsc_comm_finalize(void) { return guarded([&] { comm_finalize(); }); }
int rsoptsc_comm_info(int32_t* nranks, int32_t* rank, int32_t* p2p, int64_t* collectives, double* bytes) {
  const Comm& cm = comm();
  if (nranks) *nranks = cm.nranks;
  if (rank) *rank = cm.rank;
  if (p2p) *p2p = cm.p2p ? 1 : 0;
  if (collectives) *collectives = cm.collectives;
  if (bytes) *bytes = cm.collective_bytes;
  return 0;
}

int rsoptsc_dev_alloc(void** dptr, int64_t bytes) {
  return guarded([&] { ctx(); RS_CUDA(cudaMalloc(dptr, (size_t)bytes)); });
}
int rsoptsc_dev_free(void* dptr) { return guarded([&] { RS_CUDA(cudaFree(dptr)); }); }
int rsoptsc_dev_upload(void* dptr, const void* host, int64_t bytes) {
  return guarded([&] {
    RS_CUDA(cudaMemcpyAsync(dptr, host, (size_t)bytes, cudaMemcpyHostToDevice, ctx().stream));
    RS_CUDA(cudaStreamSynchronize(ctx().stream));
  });
}
int rsoptsc_dev_download(void* host, const void* dptr, int64_t bytes) {
  return guarded([&] {
    RS_CUDA(cudaMemcpyAsync(host, dptr, (size_t)bytes, cudaMemcpyDeviceToHost, ctx().stream));
    RS_CUDA(cudaStreamSynchronize(ctx().stream));
  });
}
int rsoptsc_dev_sync(void) { return guarded([&] { RS_CUDA(cudaStreamSynchronize(ctx().stream)); }); }

// ---- a1 ---------------------------------------------------------------------------------
int rsoptsc_normalize_dev(const double* d_data, int64_t m, int64_t n, double* d_X) {
  return guarded([&] {
    RS_REQUIRE(m > 0 && n > 0, "normalize: empty matrix");
    normalize_columns(d_data, m, n, d_X);
    RS_CUDA(cudaStreamSynchronize(ctx().stream));
  });
}
int rsoptsc_normalize(const double* data, int64_t m, int64_t n, double* X) {
  return guarded([&] {
    RS_REQUIRE(m > 0 && n > 0, "normalize: empty matrix");
    cudaStream_t s = ctx().stream;
    DevBuf<double> d((size_t)m * n), x((size_t)m * n);
    d.upload(data, (size_t)m * n, s);
    normalize_columns(d.p, m, n, x.p);
    x.download(X, (size_t)m * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

// ---- a2 ---------------------------------------------------------------------------------
int rsoptsc_pca_spectrum(const double* X, int64_t m, int64_t n, double* eig, int32_t* K, double* emb,
                         int32_t ncomp) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    DevBuf<double> x((size_t)m * n), e;
    x.upload(X, (size_t)m * n, s);
    if (emb && ncomp > 0) e.alloc((size_t)n * ncomp);
    std::vector<double> ev;
    pca_spectrum(x.p, m, n, ev, e.p, emb ? ncomp : 0);
    if (eig) std::memcpy(eig, ev.data(), sizeof(double) * ev.size());
    if (K) *K = choose_K(ev);
    if (e.p) { e.download(emb, (size_t)n * ncomp, s); RS_CUDA(cudaStreamSynchronize(s)); }
  });
}

// ---- a3 ---------------------------------------------------------------------------------
int rsoptsc_knn(const double* emb, int64_t n, int32_t dim, int32_t K, int32_t* idx) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(emb && idx && n > 0, "knn: empty embedding");
    RS_REQUIRE(dim >= 1 && dim <= 3, "knn: dim must be in [1, 3]");
    RS_REQUIRE(K >= 1 && K <= 32 && K <= n, "knn: K must be in [1, min(32, n)]");
    DevBuf<double> e((size_t)n * dim);
    DevBuf<int32_t> out((size_t)n * K);
    e.upload(emb, (size_t)n * dim, s);
    knn_bruteforce(e.p, n, n, dim, K, out.p);
    out.download(idx, (size_t)n * K, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

// ---- a4-a10 -----------------------------------------------------------------------------
static void computeM_common(const double* X, int64_t m, int64_t n, Support& sup, double lambda, double* Z, double* E,
                            int32_t* iters, int32_t* stop, double* err_trace) {
  cudaStream_t s = ctx().stream;
  DevBuf<double> x((size_t)m * n);
  x.upload(X, (size_t)m * n, s);
  AdmmState st;
  AdmmResult res;
  run_admm(x.p, m, n, sup, lambda, st, res);
  if (Z) {
    DevBuf<double> zd((size_t)n * n);
    scatter_support_dense(sup, st.Zs.p, zd.p);
    zd.download(Z, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  }
  if (E) *E = res.E;
  if (iters) *iters = res.iters;
  if (stop) *stop = res.stop;
  if (err_trace)
    for (int i = 0; i < 100; ++i) { err_trace[i] = res.err1[i]; err_trace[100 + i] = res.err2[i]; }
}
int rsoptsc_computeM(const double* X, int64_t m, int64_t n, const int32_t* idx, int32_t K, double lambda, double* Z,
                     double* E, int32_t* iters, int32_t* stop, double* err_trace) {
  return guarded([&] {
    ctx();
    RS_REQUIRE(m > 0 && n > 0, "computeM: empty matrix");
    Support sup;
    build_support_from_idx(sup, idx, n, K);
    computeM_common(X, m, n, sup, lambda, Z, E, iters, stop, err_trace);
  });
}
int rsoptsc_computeM_dense_mask(const double* D, const double* X, int64_t m, int64_t n, double lambda, double* Z,
                                double* E, int32_t* iters, int32_t* stop, double* err_trace) {
  return guarded([&] {
    ctx();
    RS_REQUIRE(m > 0 && n > 0, "computeM: empty matrix");
    Support sup;
    build_support_from_dense_mask(sup, D, n);
    computeM_common(X, m, n, sup, lambda, Z, E, iters, stop, err_trace);
  });
}

// ---- a11 --------------------------------------------------------------------------------
int rsoptsc_threshold_symmetrise_dev(const double* d_Z, int64_t n, double* d_W) {
  return guarded([&] {
    RS_REQUIRE(n > 0, "symmetrise: empty matrix");
    threshold_symmetrise_dense(d_Z, n, d_W);
    RS_CUDA(cudaStreamSynchronize(ctx().stream));
  });
}
int rsoptsc_threshold_symmetrise(const double* Z, int64_t n, double* W) {
  return guarded([&] {
    RS_REQUIRE(n > 0, "symmetrise: empty matrix");
    cudaStream_t s = ctx().stream;
    DevBuf<double> z((size_t)n * n), w((size_t)n * n);
    z.upload(Z, (size_t)n * n, s);
    threshold_symmetrise_dense(z.p, n, w.p);
    w.download(W, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

// ---- SimilarityM ------------------------------------------------------------------------
struct SimOut {
  Support sup;
  AdmmState st;
  AdmmResult res;
  int K = 0;
  DevBuf<double> emb;  // n x 3 when computed in-library
};
static void similarity_core(const double* d_data, int64_t m, int64_t n, const double* d_emb, int emb_dim,
                            double lambda, int K_override, SimOut& out, DevBuf<double>& X) {
  cudaStream_t s = ctx().stream;
  RS_REQUIRE(m > 0 && n > 1, "similarity: need a genes x cells matrix with at least 2 cells");
  X.alloc((size_t)m * n);
  normalize_columns(d_data, m, n, X.p);                       // a1
  std::vector<double> eig;
  const bool own_emb = d_emb == nullptr;
  const int ncomp = 3;                                        // No_Comps1 (quirk Q1)
  if (own_emb) out.emb.alloc((size_t)n * ncomp);
  pca_spectrum(X.p, m, n, eig, own_emb ? out.emb.p : nullptr, own_emb ? std::min<int64_t>(ncomp, std::min(m, n)) : 0);  // a2
  out.K = K_override > 0 ? K_override : choose_K(eig);
  out.K = (int)std::min<int64_t>(out.K, n);
  const double* e = own_emb ? out.emb.p : d_emb;
  const int dim = own_emb ? (int)std::min<int64_t>(ncomp, std::min(m, n)) : std::min(emb_dim, ncomp);
  RS_REQUIRE(dim >= 1, "similarity: embedding needs at least one column");
  DevBuf<int32_t> idx((size_t)n * out.K);
  knn_bruteforce(e, n, n, dim, out.K, idx.p);                 // a3
  std::vector<int32_t> h_idx((size_t)n * out.K);
  idx.download(h_idx.data(), h_idx.size(), s);
  RS_CUDA(cudaStreamSynchronize(s));
  build_support_from_idx(out.sup, h_idx.data(), n, out.K);
  run_admm(X.p, m, n, out.sup, lambda, out.st, out.res);      // a4-a10
}

int rsoptsc_similarity_dev(const double* d_data, int64_t m, int64_t n, const double* d_emb, int32_t emb_dim,
                           double lambda, int32_t K_override, double* d_W, double* E, int32_t* iters,
                           int32_t* K_out) {
  return guarded([&] {
    ctx();
    SimOut out;
    DevBuf<double> X;
    similarity_core(d_data, m, n, d_emb, emb_dim, lambda, K_override, out, X);
    if (d_W) threshold_symmetrise_sparse(out.sup, out.st.Zs.p, d_W);  // a11
    RS_CUDA(cudaStreamSynchronize(ctx().stream));
    if (E) *E = out.res.E;
    if (iters) *iters = out.res.iters;
    if (K_out) *K_out = out.K;
  });
}

int rsoptsc_similarity(const double* data, int64_t m, int64_t n, const double* emb, int32_t emb_dim, double lambda,
                       int32_t K_override, double* W, int64_t* nnz_inout, int32_t* rows, int32_t* cols, double* vals,
                       double* E, int32_t* iters, int32_t* K_out, double* emb_out) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(m > 0 && n > 1, "similarity: need a genes x cells matrix with at least 2 cells");
    DevBuf<double> d((size_t)m * n), e;
    d.upload(data, (size_t)m * n, s);
    if (emb) { e.alloc((size_t)n * emb_dim); e.upload(emb, (size_t)n * emb_dim, s); }
    SimOut out;
    DevBuf<double> X;
    similarity_core(d.p, m, n, emb ? e.p : nullptr, emb_dim, lambda, K_override, out, X);
    if (W) {
      DevBuf<double> w((size_t)n * n);
      threshold_symmetrise_sparse(out.sup, out.st.Zs.p, w.p);   // a11
      w.download(W, (size_t)n * n, s);
      RS_CUDA(cudaStreamSynchronize(s));
    }
    if (nnz_inout && rows && cols && vals) {
      // sparse W (upper triangle incl. diagonal): assembled from the n*K support values
      std::vector<double> zs(out.sup.nnz);
      out.st.Zs.download(zs.data(), zs.size(), s);
      RS_CUDA(cudaStreamSynchronize(s));
      const auto& rp = out.sup.h_rowptr;
      const auto& cl = out.sup.h_col;
      auto find = [&](int i, int j) -> int64_t {  // CSR position of (i,j) or -1
        const int* b = cl.data() + rp[i];
        const int* en = cl.data() + rp[i + 1];
        const int* it = std::lower_bound(b, en, j);
        return (it == en || *it != j) ? -1 : (int64_t)(it - cl.data());
      };
      auto thr = [&](int64_t pos) -> double {  // thresholded |Z| at a CSR position
        if (pos < 0) return 0.0;
        const double z = zs[pos];
        return z <= 2.220446049250313e-16 ? 0.0 : z;
      };
      int64_t cap = *nnz_inout, cnt = 0;
      for (int64_t i = 0; i < n; ++i)
        for (int p = rp[i]; p < rp[i + 1]; ++p) {
          const int j = cl[p];
          const int64_t mirror = i == j ? p : find(j, (int)i);
          if (j < i && mirror >= 0) continue;  // the (j,i) entry (upper triangle) emits this pair
          const double a = thr(p);
          const double b = thr(mirror);
          const double w = 0.5 * (a + b);
          if (w == 0.0) continue;
          if (cnt < cap) { rows[cnt] = (int32_t)std::min<int64_t>(i, j); cols[cnt] = (int32_t)std::max<int64_t>(i, j); vals[cnt] = w; }
          ++cnt;
        }
      *nnz_inout = cnt;
      RS_REQUIRE(cnt <= cap, "similarity: sparse output capacity too small (required count returned in nnz)");
    }
    if (emb_out) {
      if (out.emb.p) { out.emb.download(emb_out, (size_t)n * 3, s); RS_CUDA(cudaStreamSynchronize(s)); }
      else if (emb) std::memcpy(emb_out, emb, sizeof(double) * n * std::min(emb_dim, 3));
    }
    if (E) *E = out.res.E;
    if (iters) *iters = out.res.iters;
    if (K_out) *K_out = out.K;
  });
}

int rsoptsc_similarity_knn(const double* data, int64_t m, int64_t n, const int32_t* idx, int32_t K, double lambda,
                           double* W, double* E, int32_t* iters) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(data && idx && m > 0 && n > 1, "similarity_knn: need a genes x cells matrix with at least 2 cells and a neighbour table");
    RS_REQUIRE(K >= 1 && K <= n, "similarity_knn: K must be in [1, n]");
    for (int64_t t = 0; t < n * K; ++t)
      RS_REQUIRE(idx[t] >= 0 && idx[t] < n, "similarity_knn: neighbour index out of range (0-based, < n)");
    DevBuf<double> d((size_t)m * n), X((size_t)m * n);
    d.upload(data, (size_t)m * n, s);
    normalize_columns(d.p, m, n, X.p);                           // a1
    d.release();
    SimOut out;
    out.K = K;
    build_support_from_idx(out.sup, idx, n, K);                  // :90-92 with the caller's IDX
    run_admm(X.p, m, n, out.sup, lambda, out.st, out.res);       // a4-a10
    if (W) {
      DevBuf<double> w((size_t)n * n);
      threshold_symmetrise_sparse(out.sup, out.st.Zs.p, w.p);    // a11
      w.download(W, (size_t)n * n, s);
      RS_CUDA(cudaStreamSynchronize(s));
    }
    if (E) *E = out.res.E;
    if (iters) *iters = out.res.iters;
  });
}

// ---- a12-a13 ----------------------------------------------------------------------------
int rsoptsc_nmf_lee_dev(const double* d_V, int64_t n, int32_t k, int32_t max_iter, double* basis, double* coef,
                        int32_t* labels, int32_t* iters) {
  return guarded([&] {
    ctx();
    NmfResult r;
    nmf_lee(d_V, n, k, max_iter, basis, coef, labels, r);
    if (iters) *iters = r.iters;
  });
}
int rsoptsc_nmf_lee(const double* V, int64_t n, int32_t k, int32_t max_iter, double* basis, double* coef,
                    int32_t* labels, int32_t* iters) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "nmf: empty matrix");
    DevBuf<double> v((size_t)n * n);
    v.upload(V, (size_t)n * n, s);
    NmfResult r;
    nmf_lee(v.p, n, k, max_iter, basis, coef, labels, r);
    if (iters) *iters = r.iters;
  });
}

int rsoptsc_nmf_sweep(const double* V, int64_t n, int32_t k_lo, int32_t k_hi, int32_t max_iter, int32_t* labels,
                      int32_t* iters) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(V && n > 0, "nmf_sweep: empty matrix");
    DevBuf<double> d((size_t)n * n);
    d.upload(V, (size_t)n * n, s);
    nmf_sweep(d.p, n, k_lo, k_hi, max_iter, labels, iters);
  });
}
int rsoptsc_nmf_sweep_dev(const double* d_V, int64_t n, int32_t k_lo, int32_t k_hi, int32_t max_iter, int32_t* labels,
                          int32_t* iters) {
  return guarded([&] {
    ctx();
    nmf_sweep(d_V, n, k_lo, k_hi, max_iter, labels, iters);
  });
}

// ---- a14-a18 ----------------------------------------------------------------------------
int rsoptsc_get_components(const double* S, int64_t n, double tol, double* vals, int32_t* n_eigs) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "get_components: empty matrix");
    DevBuf<double> d((size_t)n * n);
    d.upload(S, (size_t)n * n, s);
    get_components(d.p, n, tol, vals, n_eigs);
  });
}
int rsoptsc_consensus(const int32_t* labels, int32_t R, int64_t n, double tau, double* consen) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0 && R > 0, "consensus: empty label table");
    DevBuf<int32_t> l((size_t)R * n);
    DevBuf<double> c((size_t)n * n);
    l.upload(labels, (size_t)R * n, s);
    consensus_dense(l.p, R, n, tau, c.p);
    c.download(consen, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}
int rsoptsc_pam(const double* P, int64_t n, int32_t dim, int32_t k, int32_t* labels, int32_t* medoids) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0 && dim > 0, "pam: empty input");
    DevBuf<double> p((size_t)n * dim);
    p.upload(P, (size_t)n * dim, s);
    pam(p.p, n, dim, k, labels, medoids);
  });
}
int rsoptsc_count_clusters(const double* S, int64_t n, double tol, int32_t range_lo, int32_t range_hi, int32_t n_comp,
                           int32_t* upper_bound, int32_t* lower_bound, double* vals, int32_t* n_eigs) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "count_clusters: empty matrix");
    DevBuf<double> d((size_t)n * n);
    d.upload(S, (size_t)n * n, s);
    count_clusters(d.p, n, tol, range_lo, range_hi, n_comp, upper_bound, lower_bound, vals, n_eigs);
  });
}

int rsoptsc_get_ensemble(const double* S, int64_t n, int32_t n_comp, double tau, int32_t range_lo, int32_t range_hi,
                         double* consen) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(S && consen && n > 0, "get_ensemble: empty matrix");
    DevBuf<double> d((size_t)n * n), c((size_t)n * n);
    d.upload(S, (size_t)n * n, s);
    get_ensemble(d.p, n, n_comp, tau, range_lo, range_hi, c.p);
    c.download(consen, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

// ---- RepresentationMap pieces --------------------------------------------------------------
int rsoptsc_dist_flat(const double* emb, int64_t n, int32_t dim, double* dist) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0 && dim > 0, "dist_flat: empty embedding");
    DevBuf<double> e((size_t)n * dim), d((size_t)n * n);
    e.upload(emb, (size_t)n * dim, s);
    dist_flat(e.p, n, dim, d.p);
    d.download(dist, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}
int rsoptsc_adjacency(const double* S, int64_t n, double* adj) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "adjacency: empty matrix");
    DevBuf<double> d((size_t)n * n), a((size_t)n * n);
    d.upload(S, (size_t)n * n, s);
    adjacency(d.p, n, a.p);
    a.download(adj, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}
int rsoptsc_graph_components(const double* adj, int64_t n, int32_t* membership, int32_t* n_components) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "graph_components: empty matrix");
    DevBuf<double> a((size_t)n * n);
    a.upload(adj, (size_t)n * n, s);
    graph_components(a.p, n, membership, n_components);
  });
}
int rsoptsc_graph_distances(const double* adj, int64_t n, double* dist) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "graph_distances: empty matrix");
    DevBuf<double> a((size_t)n * n), d((size_t)n * n);
    a.upload(adj, (size_t)n * n, s);
    graph_distances(a.p, n, d.p);
    d.download(dist, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

// ---- preprocessing (R/DataSelection.R) -------------------------------------------------------
int rsoptsc_log_normalize(const double* M, int64_t m, int64_t n, double scale, int32_t normalize, double* out) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(m > 0 && n > 0, "log_normalize: empty matrix");
    DevBuf<double> a((size_t)m * n), o((size_t)m * n);
    a.upload(M, (size_t)m * n, s);
    log_normalize(a.p, m, n, scale, normalize != 0, o.p);
    o.download(out, (size_t)m * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}
int rsoptsc_scale_center(const double* M, int64_t m, int64_t n, double* out) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(m > 0 && n > 0, "scale_center: empty matrix");
    DevBuf<double> a((size_t)m * n), o((size_t)m * n);
    a.upload(M, (size_t)m * n, s);
    scale_center(a.p, m, n, o.p);
    o.download(out, (size_t)m * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}
int rsoptsc_count_threshold(const double* M, int64_t m, int64_t n, double countL, double countU, double X,
                            int32_t* keep) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(m > 0 && n > 0, "count_threshold: empty matrix");
    DevBuf<double> a((size_t)m * n);
    a.upload(M, (size_t)m * n, s);
    std::vector<double> st;
    gene_stats(a.p, m, n, countL, countU, st);
    for (int64_t g = 0; g < m; ++g) {                       // R/DataSelection.R:94-102
      const bool upper = !((st[4 * m + g] / (double)n) >= (100.0 - X) * 0.01);
      const bool lower = (st[3 * m + g] / (double)n) >= X * 0.01;
      keep[g] = (upper && lower) ? 1 : 0;
    }
  });
}
int rsoptsc_select_data(const double* M, int64_t m, int64_t n, double gene_expression_threshold, int32_t n_features,
                        int32_t* gene_use, int32_t* n_selected, int32_t* max_component) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(m > 0 && n > 0, "select_data: empty matrix");
    DevBuf<double> a((size_t)m * n);
    a.upload(M, (size_t)m * n, s);
    std::vector<int> sel;
    int mc = 0;
    select_data(a.p, m, n, gene_expression_threshold, n_features, sel, &mc);
    for (size_t i = 0; i < sel.size(); ++i) gene_use[i] = sel[i];
    if (n_selected) *n_selected = (int32_t)sel.size();
    if (max_component) *max_component = mc;
  });
}

int rsoptsc_signaling_pair(const double* lig, const double* rec, const double* beta, const double* gamma, int64_t n,
                           int32_t normalize, double* P) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(n > 0, "signaling_pair: empty input");
    DevBuf<double> l(n), r(n), b, g, p((size_t)n * n);
    l.upload(lig, n, s);
    r.upload(rec, n, s);
    if (beta) { b.alloc(n); b.upload(beta, n, s); }
    if (gamma) { g.alloc(n); g.upload(gamma, n, s); }
    signaling_pair(l.p, r.p, b.p, g.p, n, normalize != 0, p.p);
    p.download(P, (size_t)n * n, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

int rsoptsc_set_gemm_backend(int32_t backend) {
  return guarded([&] {
    RS_REQUIRE(backend >= 0 && backend <= 2, "gemm backend must be 0 (cublas), 1 (ozaki, large only) or 2 (ozaki, forced)");
    set_gemm_backend(backend);
  });
}
int32_t rsoptsc_get_gemm_backend(void) { return gemm_backend(); }
int rsoptsc_debug_gemm(int32_t op, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                       int32_t backend, int32_t reps, double* ms_out) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(op >= 0 && op <= 4 && M > 0 && N > 0 && K > 0 && reps >= 1, "debug_gemm: bad arguments");
    if (op == 0 || op == 4) N = M;
    const size_t na = (size_t)M * K, nb = op == 0 ? 0 : (size_t)N * K, nc = (size_t)M * N;
    DevBuf<double> a(na), b(std::max<size_t>(nb, 1)), c(nc);
    a.upload(A, na, s);
    if (nb) b.upload(B, nb, s);
    if (op == 4) c.upload(C, nc, s);  // in/out: the update is applied 1 + reps times
    const int saved = gemm_backend();
    set_gemm_backend(backend);
    cudaEvent_t e0, e1;
    RS_CUDA(cudaEventCreate(&e0)); RS_CUDA(cudaEventCreate(&e1));
    auto call = [&] {
      switch (op) {
        case 0: gram_AtA(a.p, K, M, c.p); break;
        case 1: gemm_nn(a.p, b.p, c.p, M, N, K); break;
        case 2: gemm_nt(a.p, b.p, c.p, M, N, K); break;
        case 4: syr2k_lower_sub(a.p, M, b.p, M, M, (int)K, c.p, M); break;
        default: gemm_tn(a.p, b.p, c.p, M, N, K); break;
      }
    };
    try {
      call();  // warm-up (also sets function attributes)
      RS_CUDA(cudaEventRecord(e0, s));
      for (int r = 0; r < reps; ++r) call();
      RS_CUDA(cudaEventRecord(e1, s));
      RS_CUDA(cudaEventSynchronize(e1));
    } catch (...) {
      set_gemm_backend(saved);
      throw;
    }
    set_gemm_backend(saved);
    float ms = 0;
    RS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_out) *ms_out = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c.download(C, nc, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

int rsoptsc_debug_sym_eig(const double* A, int64_t n, int32_t solver, double* lam_out, double* V_out, int32_t reps,
                          double* ms_out) {
  return guarded([&] {
    cudaStream_t s = ctx().stream;
    RS_REQUIRE(A && lam_out && n >= 1 && n <= 65535 && reps >= 1 && solver >= 0 && solver <= 3, "debug_sym_eig: bad arguments");
    const size_t nn = (size_t)n * n;
    DevBuf<double> a0(nn), a(nn), lam(n);
    a0.upload(A, nn, s);
    cudaEvent_t e0, e1;
    RS_CUDA(cudaEventCreate(&e0)); RS_CUDA(cudaEventCreate(&e1));
    float total = 0;
    for (int r = 0; r <= reps; ++r) {  // r == 0 is the warm-up
      RS_CUDA(cudaMemcpyAsync(a.p, a0.p, sizeof(double) * nn, cudaMemcpyDeviceToDevice, s));
      RS_CUDA(cudaEventRecord(e0, s));
      if (solver == 2) sym_eig(a.p, n, lam.p, true, -1.0e300, /*collective=*/true);  // all ranks of the communicator
      else if (solver == 3) sym_eig(a.p, n, lam.p, /*vectors=*/false);              // values only (a2, a14)
      else sym_eig_select(a.p, n, lam.p, solver);
      RS_CUDA(cudaEventRecord(e1, s));
      RS_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      RS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) total += ms;
    }
    if (ms_out) *ms_out = total / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    lam.download(lam_out, n, s);
    if (V_out && solver != 3) a.download(V_out, nn, s);
    RS_CUDA(cudaStreamSynchronize(s));
  });
}

int rsoptsc_host_tridiag_bisect(int32_t k, const double* d, const double* e, double* lam) {
  return guarded([&] {
    RS_REQUIRE(k >= 1 && d && lam && (e || k == 1), "tridiag_bisect: bad arguments");
    std::vector<double> e2(std::max(1, k - 1), 0.0);
    double lo = INFINITY, hi = -INFINITY, em = 0.0;
    for (int i = 0; i < k; ++i) {
      const double el = i > 0 ? std::fabs(e[i - 1]) : 0.0, er = i + 1 < k ? std::fabs(e[i]) : 0.0;
      lo = std::min(lo, d[i] - el - er);
      hi = std::max(hi, d[i] + el + er);
      if (i + 1 < k) { e2[i] = e[i] * e[i]; em = std::max(em, e2[i]); }
    }
    const double pivmin = DBL_MIN * std::max(1.0, em), tnorm = std::max(std::fabs(lo), std::fabs(hi));
    lo -= 2.0 * tnorm * DBL_EPSILON * k + 2.0 * pivmin;
    hi += 2.0 * tnorm * DBL_EPSILON * k + 2.0 * pivmin;
    for (int i = 0; i < k; ++i) lam[i] = bisect::kth_eigenvalue(d, e2.data(), k, i, lo, hi, pivmin);
  });
}

int rsoptsc_host_tridiag_eig(int32_t k, double* d, const double* e, double* V) {
  return guarded([&] {
    RS_REQUIRE(k >= 1, "tridiag_eig: k must be positive");
    std::vector<double> dd(d, d + k), ee(e, e + std::max(0, k - 1)), vv;
    tridiag_eig(dd, ee, V ? &vv : nullptr);
    std::memcpy(d, dd.data(), sizeof(double) * k);
    if (V) std::memcpy(V, vv.data(), sizeof(double) * (size_t)k * k);
  });
}
int rsoptsc_host_mg_tables(int64_t n, int32_t nranks, int32_t rank, int32_t* ownI, int32_t* nl, int64_t* pref, int32_t* lfirst) {
  return guarded([&] {
    RS_REQUIRE(n >= 1 && n <= 65535 && nranks >= 1 && nranks <= MAX_RANKS && rank >= 0 && rank < nranks,
               "host_mg_tables: bad arguments");
    std::vector<int> o, l;
    std::vector<long long> p;
    sytrd_mg_tables_host(n, nranks, rank, o, p, l);
    if (nl) *nl = (int32_t)o.size();
    if (ownI) std::copy(o.begin(), o.end(), ownI);
    if (pref) std::copy(p.begin(), p.end(), pref);
    if (lfirst) std::copy(l.begin(), l.end(), lfirst);
  });
}
int rsoptsc_host_shard_owner(const int32_t* idx, int64_t n, int32_t K, int32_t nranks, int64_t split_min,
                             int32_t* owner, int32_t* component, int32_t* split) {
  return guarded([&] {
    RS_REQUIRE(idx && n > 0 && K > 0 && nranks >= 1 && nranks <= MAX_RANKS, "host_shard_owner: bad arguments");
    Support sup;
    build_support_from_idx(sup, idx, n, K, /*device=*/false);
    for (int r = 0; r < nranks; ++r) {
      ShardPlan plan;
      build_shard_plan(sup, nranks, r, split_min, plan, /*device=*/false);
      if (owner)
        for (int j : plan.own_cols) owner[j] = r;
      if (r == 0)
        for (int64_t j = 0; j < n; ++j) {
          if (component) component[j] = sup.h_comp[j];
          if (split) split[j] = plan.big[sup.h_comp[j]] ? 1 : 0;
        }
    }
  });
}
int32_t rsoptsc_host_choose_K(const double* eig, int64_t len) {
  std::vector<double> v(eig, eig + len);
  return choose_K(v);
}

}  // extern "C"
