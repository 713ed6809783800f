// Host side of the fused multi-GPU tridiagonalisation (sytrd_mg.cuh): peer-mapped exchange regions from
// comm.cu, one cooperative launch per 64-column panel on every rank, the ranks meeting inside the kernel.
#include "sytrd_mg.cuh"

#include "comm.h"
#include "internal.h"

namespace rs {

static int64_t mg_min_n() {
  static int64_t t = -1;
  if (t < 0) {
    // below ~3,000 rows a column streams in under 2 us per rank: the NVLink exchange costs more than it saves
    const char* e = getenv("RSOPTSC_SYTRD_MG_MIN_N");
    t = e ? atoll(e) : 3000;
    if (t < 2 * sytrd::NB) t = 2 * sytrd::NB;
  }
  return t;
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
  // arrival counters restart at zero; nobody may signal a region before its owner has cleared it
  RS_CUDA(cudaMemsetAsync(cm.xchg[cm.rank], 0, sizeof(double) * sytrd::XCH_HDR, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  comm_barrier();
  DevBuf<double> work(sytrd::workspace_doubles(n));
  sytrd::Hooks hooks;
  Phase* panel_phase = nullptr;
  hooks.user = &panel_phase;
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
  RS_CUDA(sytrd::run_mg(c.cublas, c.stream, d_A, n, n, d_d, d_e, d_tau, work.p, cm.nranks, cm.rank, cm.xchg, grid,
                        &launches, &hooks));
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  cm.collectives += (n + sytrd::NB - 1) / sytrd::NB;  // one fused compute + exchange kernel per panel
  cm.collective_bytes += 8.0 * (double)n * (double)n / 2.0 * (double)(cm.nranks - 1);  // y vectors pushed to the peers
}

}  // namespace rs
