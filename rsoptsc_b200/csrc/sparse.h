// Compressed sparse views of a dense column-major n x n matrix (nmf.cu) shared by the NMF, the
// Laplacian spectrum and the sparse PCA of CountClusters.
#pragma once
#include "common.h"

namespace rs {

struct Compressed {  // CSR (by row) or CSC (by column) of an n x n matrix
  int64_t n = 0, nnz = 0;
  DevBuf<int> ptr, idx;
  DevBuf<double> val;
};

// dense -> compressed (entries != 0), indices ascending inside each line
void compress(const double* d_V, int64_t n, bool by_row, Compressed& out);
// y[r] = sum_e val[e] x[idx[e]] over the entries of compressed line r
void spmv(const Compressed& A, const double* x, double* y);

}  // namespace rs
