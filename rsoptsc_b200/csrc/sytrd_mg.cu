// Host side of the fused multi-GPU tridiagonalisation (sytrd_mg.cuh): peer-mapped exchange regions from
// comm.cu, one cooperative launch per 64-column panel on every rank, the ranks meeting inside the kernel.
#include "sytrd_mg.cuh"

#include "comm.h"
#include "gemm.h"
#include "internal.h"

namespace rs {

void sytrd_trailing_update(const double* V, long long ldv, const double* W, long long ldw, long long nr, int k, double* C,
                           long long ldc, long long r0, void* user);  // eig.cu

static int64_t mg_min_n() {
  // the exchange adds ~3 us per column; the product phase of a column takes 14.5 us (n / 7,954)^2 on one GPU and
  // 1/P of that on P: below these sizes the exchange costs more than it saves
  const char* e = getenv("RSOPTSC_SYTRD_MG_MIN_N");
  int64_t t = e ? atoll(e) : (comm().nranks >= 4 ? 4000 : 6000);
  if (t < 2 * sytrd::NB) t = 2 * sytrd::NB;
  return t;
}

void sytrd_mg_tables_host(int64_t n, int nranks, int rank, std::vector<int>& ownI, std::vector<long long>& pref,
                          std::vector<int>& lfirst) {
  sytrd::mg_tables_host(n, nranks, rank, ownI, pref, lfirst);
}

bool sytrd_mg_available(int64_t n) {
  if (!sharded() || n < mg_min_n() || comm().nranks > sytrd::MG_MAX_RANKS) return false;
  const char* off = getenv("RSOPTSC_SYTRD_MG");
  if (off && off[0] == '0') return false;
  return comm_ensure_xchg(sytrd::mg_region_doubles(n, comm().nranks));
}

void sytrd_mg_run(double* d_A, int64_t n, double* d_d, double* d_e, double* d_tau) {
  Context& c = ctx();
  Comm& cm = comm();
  RS_REQUIRE(cm.p2p && cm.xchg_doubles >= sytrd::mg_region_doubles(n, cm.nranks), "sytrd_mg: exchange regions not mapped");
  RS_REQUIRE(n <= 65535, "sytrd_mg: more than 65,535 rows");
  // packet flags = epoch << 16 | column + 1: a fresh epoch per factorisation (identical on every rank: the library is
  // SPMD), the regions cleared whenever the 16-bit counter wraps or the regions were re-created
  static unsigned epoch = 0;
  static const double* epoch_region = nullptr;
  if (epoch_region != cm.xchg[cm.rank]) { epoch = 0; epoch_region = cm.xchg[cm.rank]; }
  if (++epoch > 65535u) epoch = 1;
  if (epoch == 1) {
    RS_CUDA(cudaStreamSynchronize(c.stream));
    comm_barrier();  // no peer is still writing packets of an earlier factorisation
    RS_CUDA(cudaMemsetAsync(cm.xchg[cm.rank], 0, sizeof(double) * cm.xchg_doubles, c.stream));
    RS_CUDA(cudaStreamSynchronize(c.stream));
    comm_barrier();  // nobody writes into a region before its owner has cleared it
  }
  DevBuf<double> work(sytrd::workspace_doubles(n));
  DevBuf<char> tabbuf(sytrd::mg_tab_bytes(n));
  sytrd::Hooks hooks;
  Phase* panel_phase = nullptr;
  hooks.user = &panel_phase;
  hooks.syr2k = sytrd_trailing_update;
  // RSOPTSC_SYR2K_SHARD=0 keeps the trailing update replicated (every rank updates everything; tests compare)
  static const bool shard_update = !(getenv("RSOPTSC_SYR2K_SHARD") && getenv("RSOPTSC_SYR2K_SHARD")[0] == '0');
  int shard_info[2] = {cm.nranks, cm.rank};
  hooks.syr2k_user = (shard_update && gemm_backend() != 0) ? shard_info : nullptr;
  hooks.before = [](void* u, double bytes) {
    add_counter("sytrd_panel_bytes", bytes);
    *static_cast<Phase**>(u) = new Phase("sytrd_panel");
  };
  hooks.after = [](void* u) {
    delete *static_cast<Phase**>(u);
    *static_cast<Phase**>(u) = nullptr;
  };
  int64_t launches = 0;
  const int grid = std::min(c.sm_count, sytrd::MAX_GRID);  // identical on every rank (arrival targets count CTAs)
  RS_CUDA(sytrd::run_mg(c.cublas, c.stream, d_A, n, n, d_d, d_e, d_tau, work.p, tabbuf.p, cm.nranks, cm.rank, cm.xchg, epoch, grid,
                        &launches, &hooks));
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  cm.collectives += (n + sytrd::NB - 1) / sytrd::NB;  // one fused compute + exchange kernel per panel
  cm.collective_bytes += 16.0 * (double)n * (double)n / 2.0 * (double)(cm.nranks - 1);  // 16-byte y packets pushed to the peers
}

}  // namespace rs
