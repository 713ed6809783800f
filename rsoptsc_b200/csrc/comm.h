// Multi-GPU plumbing of librsoptsc_b200.so: one process per GPU, one NCCL communicator per process
// (SURVEY.md section 8e).  The library is SPMD once a communicator with more than one rank exists:
// every rank calls the same entry point with the same (replicated) inputs and receives the same
// outputs; inside, the column-local ADMM state, the SVT contractions, the eigensolver stages and the
// NMF row blocks are split across the ranks.
//
//   NCCL over NVLink      allreduce of the small Gram matrices and scalars of the stop rule,
//                         all-gather of column panels of the SVT operands (grouped broadcasts: the
//                         panels have unequal widths)
//   peer-mapped memory    the per-column exchange of the fused multi-GPU tridiagonalisation
//                         (sytrd_mg.cuh): every rank maps every other rank's exchange region with
//                         CUDA IPC and writes its partial products there with plain NVLink stores
#pragma once
#include <vector>

#include "common.h"

namespace rs {

constexpr int MAX_RANKS = 8;

struct Comm {
  int nranks = 1, rank = 0;
  void* nccl = nullptr;         // ncclComm_t
  bool p2p = false;             // exchange regions of all ranks are mapped into this process
  // exchange region of the fused tridiagonalisation; xchg[q] = rank q's region as seen from here
  double* xchg[MAX_RANKS] = {};
  size_t xchg_doubles = 0;
  int64_t collectives = 0;      // data-path collectives issued (bench.py reports them)
  double collective_bytes = 0;  // payload this rank contributed / received
};

Comm& comm();
inline bool sharded() { return comm().nranks > 1; }

void comm_init(int nranks, int rank, const void* unique_id_128);
void comm_finalize();
void comm_unique_id(void* out_128);

// in-place sum / max over ranks on the library stream (device buffers); bitwise identical on every rank
void comm_allreduce_sum(double* d_buf, size_t count);
double comm_sum(double v);   // host scalars (one small device round trip)
double comm_max(double v);
void comm_barrier();
// Column-major matrix with `rows` rows replicated on every rank; rank q has just produced columns
// [begin[q], begin[q+1]).  On return every rank holds every column.
void comm_allgather_cols(double* d_base, int64_t rows, const std::vector<int64_t>& begin);
// Column panel [r0, r0 + ncols) x rows [r0, n) of a replicated column-major matrix whose rows are CURRENT only on
// their owner -- row i belongs to rank (i / group) mod nranks -- made current on every rank (pack, all-gather, unpack).
void comm_gather_cyclic_rows(double* d_A, int64_t lda, int64_t n, int64_t r0, int ncols, int group);
// (re)allocate the peer-mapped exchange regions (collective).  Returns false when CUDA IPC is not
// available between the ranks: the callers then keep the tridiagonalisation replicated.
bool comm_ensure_xchg(size_t doubles);

// equal split of [0, count) into nranks contiguous ranges, aligned to `align`
std::vector<int64_t> split_even(int64_t count, int nranks, int64_t align = 1);
// split of the columns of a lower triangle (column j has count - j entries) into ranges of equal area
std::vector<int64_t> split_lower_triangle(int64_t count, int nranks, int64_t align = 1);

}  // namespace rs
