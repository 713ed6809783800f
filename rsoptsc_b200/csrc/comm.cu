// NCCL communicator + peer-mapped exchange regions of the multi-GPU path (comm.h).
#include "comm.h"

#include <nccl.h>

#include <algorithm>
#include <cmath>

#include "internal.h"

namespace rs {

#define RS_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t _e = (expr);                                                                      \
    if (_e != ncclSuccess)                                                                         \
      ::rs::fail(7, __FILE__, __LINE__, std::string("NCCL: ") + ncclGetErrorString(_e));           \
  } while (0)

static Comm g_comm;
Comm& comm() { return g_comm; }
static inline ncclComm_t nc() { return static_cast<ncclComm_t>(g_comm.nccl); }

// peer mappings opened with cudaIpcOpenMemHandle (closed before the owner frees its region)
static void* g_xchg_local = nullptr;
static void* g_xchg_mapped[MAX_RANKS] = {};

static void release_xchg() {
  for (int q = 0; q < MAX_RANKS; ++q) {
    if (g_xchg_mapped[q]) cudaIpcCloseMemHandle(g_xchg_mapped[q]);
    g_xchg_mapped[q] = nullptr;
    g_comm.xchg[q] = nullptr;
  }
  if (g_comm.nccl && g_comm.nranks > 1) {  // nobody frees while a peer may still have the region mapped
    cudaStreamSynchronize(ctx().stream);
    comm_barrier();
  }
  if (g_xchg_local) cudaFree(g_xchg_local);
  g_xchg_local = nullptr;
  g_comm.xchg_doubles = 0;
  g_comm.p2p = false;
}

void comm_unique_id(void* out_128) {
  static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
  ncclUniqueId id;
  RS_NCCL(ncclGetUniqueId(&id));
  std::memcpy(out_128, &id, sizeof id);
}

void comm_init(int nranks, int rank, const void* unique_id_128) {
  RS_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS, "comm_init: 1 to 8 ranks (one node)");
  RS_REQUIRE(rank >= 0 && rank < nranks, "comm_init: rank out of range");
  Context& c = ctx();
  RS_CUDA(cudaSetDevice(c.device));
  if (g_comm.nccl) comm_finalize();
  if (nranks == 1) { g_comm = Comm(); return; }
  RS_REQUIRE(unique_id_128 != nullptr, "comm_init: the NCCL unique id of rank 0 is required");
  ncclUniqueId id;
  std::memcpy(&id, unique_id_128, sizeof id);
  ncclComm_t cm = nullptr;
  RS_NCCL(ncclCommInitRank(&cm, nranks, id, rank));
  g_comm = Comm();
  g_comm.nranks = nranks;
  g_comm.rank = rank;
  g_comm.nccl = cm;
  comm_barrier();
}

void comm_finalize() {
  if (!g_comm.nccl) { g_comm = Comm(); return; }
  release_xchg();
  cudaStreamSynchronize(ctx().stream);
  ncclCommDestroy(nc());
  g_comm = Comm();
}

void comm_allreduce_sum(double* d_buf, size_t count) {
  if (!sharded() || !count) return;
  RS_NCCL(ncclAllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, nc(), ctx().stream));
  g_comm.collectives++;
  g_comm.collective_bytes += 8.0 * (double)count;
}

static double scalar_reduce(double v, ncclRedOp_t op) {
  if (!sharded()) return v;
  static DevBuf<double>* buf = nullptr;  // leaked on purpose (static destruction order vs the pool)
  if (!buf) buf = new DevBuf<double>(8);
  cudaStream_t s = ctx().stream;
  RS_CUDA(cudaMemcpyAsync(buf->p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  RS_NCCL(ncclAllReduce(buf->p, buf->p, 1, ncclDouble, op, nc(), s));
  double out = 0;
  RS_CUDA(cudaMemcpyAsync(&out, buf->p, sizeof(double), cudaMemcpyDeviceToHost, s));
  RS_CUDA(cudaStreamSynchronize(s));
  g_comm.collectives++;
  g_comm.collective_bytes += 8.0;
  return out;
}
double comm_sum(double v) { return scalar_reduce(v, ncclSum); }
double comm_max(double v) { return scalar_reduce(v, ncclMax); }
void comm_barrier() {
  if (!sharded()) return;
  scalar_reduce(0.0, ncclSum);
}

void comm_allgather_cols(double* d_base, int64_t rows, const std::vector<int64_t>& begin) {
  if (!sharded()) return;
  RS_REQUIRE((int)begin.size() == g_comm.nranks + 1, "allgather_cols: one range per rank");
  Phase ph("allgather");
  cudaStream_t s = ctx().stream;
  RS_NCCL(ncclGroupStart());
  for (int q = 0; q < g_comm.nranks; ++q) {
    const int64_t w = begin[q + 1] - begin[q];
    if (w <= 0) continue;
    double* p = d_base + (size_t)begin[q] * rows;
    RS_NCCL(ncclBroadcast(p, p, (size_t)w * rows, ncclDouble, q, nc(), s));
    g_comm.collective_bytes += 8.0 * (double)w * (double)rows;
  }
  RS_NCCL(ncclGroupEnd());
  g_comm.collectives++;
}

// ---- gather of a column panel whose ROWS are owned cyclically in groups (sharded trailing update of sytrd_mg) ----
// rows of rank q below x: full cycles, the full groups of the last cycle, the partial group
__host__ __device__ inline long long cyc_rows_below(long long x, int group, int P, int q) {
  const long long G = x / group, cyc = G / P, rem = G % P;
  return cyc * group + (rem > q ? group : 0) + (rem == q ? x % group : 0);
}
// dir 0: pack the rows of `rank` into buf[(idx) + c * lmax]; dir 1: unpack the rows of every OTHER rank from
// buf[q * lmax * ncols + idx + c * lmax]
__global__ void k_cyc_panel(double* __restrict__ A, long long lda, int n, int r0, int ncols, int group, int P, int rank, int lmax,
                            double* __restrict__ buf, int dir) {
  const int c = blockIdx.y;
  for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int q = (i / group) % P;
    if ((dir == 0) != (q == rank)) continue;
    const long long idx = cyc_rows_below(i, group, P, q) - cyc_rows_below(r0, group, P, q);
    double* a = A + i + (long long)(r0 + c) * lda;
    double* b = buf + (dir == 0 ? 0 : (size_t)q * lmax * ncols) + idx + (size_t)c * lmax;
    if (dir == 0) *b = *a; else *a = *b;
  }
}
void comm_gather_cyclic_rows(double* d_A, int64_t lda, int64_t n, int64_t r0, int ncols, int group) {
  if (!sharded() || ncols < 1 || r0 >= n) return;
  Phase ph("panel_gather");
  cudaStream_t s = ctx().stream;
  const int P = g_comm.nranks;
  long long lmax = 0;
  for (int q = 0; q < P; ++q) lmax = std::max(lmax, cyc_rows_below(n, group, P, q) - cyc_rows_below(r0, group, P, q));
  if (lmax < 1) return;
  DevBuf<double> buf((size_t)P * lmax * ncols);
  double* mine = buf.p + (size_t)g_comm.rank * lmax * ncols;
  dim3 grid((unsigned)std::min<int64_t>(cdiv(n - r0, 256), 128), (unsigned)ncols);
  RS_LAUNCH(k_cyc_panel, grid, 256, 0, d_A, (long long)lda, (int)n, (int)r0, ncols, group, P, g_comm.rank, (int)lmax, mine, 0);
  RS_NCCL(ncclAllGather(mine, buf.p, (size_t)lmax * ncols, ncclDouble, nc(), s));
  RS_LAUNCH(k_cyc_panel, grid, 256, 0, d_A, (long long)lda, (int)n, (int)r0, ncols, group, P, g_comm.rank, (int)lmax, buf.p, 1);
  g_comm.collectives++;
  g_comm.collective_bytes += 8.0 * (double)lmax * ncols * (P - 1);
}

bool comm_ensure_xchg(size_t doubles) {
  if (!sharded()) return false;
  if (g_comm.xchg_doubles >= doubles) return g_comm.p2p;
  static bool refused = false;  // CUDA IPC unavailable: do not retry on every solve
  if (refused) return false;
  release_xchg();
  Context& c = ctx();
  cudaStream_t s = c.stream;
  const int P = g_comm.nranks;
  const size_t bytes = ((doubles * sizeof(double) + (1u << 21) - 1) >> 21) << 21;
  double ok = 1.0;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof mine);
  if (cudaMalloc(&g_xchg_local, bytes) != cudaSuccess) { cudaGetLastError(); g_xchg_local = nullptr; ok = 0.0; }
  if (ok > 0) {
    cudaMemsetAsync(g_xchg_local, 0, bytes, s);
    if (cudaIpcGetMemHandle(&mine, g_xchg_local) != cudaSuccess) { cudaGetLastError(); ok = 0.0; }
  }
  // all-gather the handles through NCCL (byte payload staged on the device)
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  DevBuf<char> stage((size_t)64 * P);
  RS_CUDA(cudaMemcpyAsync(stage.p + 64 * g_comm.rank, &mine, 64, cudaMemcpyHostToDevice, s));
  RS_NCCL(ncclAllGather(stage.p + 64 * g_comm.rank, stage.p, 64, ncclChar, nc(), s));
  std::vector<cudaIpcMemHandle_t> all(P);
  RS_CUDA(cudaMemcpyAsync(all.data(), stage.p, (size_t)64 * P, cudaMemcpyDeviceToHost, s));
  RS_CUDA(cudaStreamSynchronize(s));
  ok = -comm_max(-ok);  // min over ranks: every rank allocated and exported
  if (ok > 0) {
    for (int q = 0; q < P; ++q) {
      if (q == g_comm.rank) { g_comm.xchg[q] = static_cast<double*>(g_xchg_local); continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0.0;
        break;
      }
      g_xchg_mapped[q] = p;
      g_comm.xchg[q] = static_cast<double*>(p);
    }
  }
  ok = -comm_max(-ok);
  if (!(ok > 0)) {
    release_xchg();
    refused = true;
    return false;
  }
  g_comm.xchg_doubles = bytes / sizeof(double);
  g_comm.p2p = true;
  return true;
}

std::vector<int64_t> split_even(int64_t count, int nranks, int64_t align) {
  std::vector<int64_t> b(nranks + 1, 0);
  for (int q = 1; q < nranks; ++q) {
    int64_t v = count * q / nranks;
    v = (v + align / 2) / align * align;
    b[q] = std::min(count, std::max(b[q - 1], v));
  }
  b[nranks] = count;
  return b;
}

std::vector<int64_t> split_lower_triangle(int64_t count, int nranks, int64_t align) {
  // area left of column c: c count - c (c - 1) / 2 ~ count^2/2 (1 - (1 - c/count)^2)
  std::vector<int64_t> b(nranks + 1, 0);
  for (int q = 1; q < nranks; ++q) {
    const double f = 1.0 - std::sqrt(1.0 - (double)q / nranks);
    int64_t v = (int64_t)std::llround(f * (double)count);
    v = (v + align / 2) / align * align;
    b[q] = std::min(count, std::max(b[q - 1], v));
  }
  b[nranks] = count;
  return b;
}

}  // namespace rs
